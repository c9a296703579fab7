// table.cuh -- the open-addressing site table every kernel probes (m_counts / m_kmerToHash of the reference:
// src/FingerPrint.hpp:466, src/MultiCount.hpp:207), built on the device by build_tables_kernel (kernels.cuh).
#pragma once
#include <stdint.h>

namespace ntsm {

struct __align__(16) TableSlot {
	uint64_t key;    // reference hash64 value; kEmptyKey = unused
	uint32_t idx;    // dense k-mer index into counts[]
	uint32_t pad;
};
constexpr uint64_t kEmptyKey = ~0ULL;

}  // namespace ntsm
