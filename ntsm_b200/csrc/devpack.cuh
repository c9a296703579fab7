// devpack.cuh -- decode + pack ON THE DEVICE: ASCII reads (as they lie in the caller's pinned host
// buffer, DMA'd over PCIe untouched by any host core) -> the packed "2-bit + N-mask" position stream
// the count kernels scan (layout: include/ntsm_b200.h; host twin: pack.cpp).
//
// This is the table lookup of KseqHashIterator::step (vendor/KseqHashIterator.hpp:96-97,114-127)
// done once per base by a GPU thread instead of a host packer thread: the host then spends zero
// memory traffic per base (the DMA engine reads 1 byte/base), where the host packer costs
// ~2.1 bytes/base of DRAM traffic (read ASCII, write-allocate + write the packed words, DMA read).
// On a box whose host feeds N GPUs from a fixed number of cores that is what lets the in-memory
// insertCount path scale with the GPUs.
//
// Output layout is exactly the host packer's: read r occupies read_span(len) = (len + 8) & ~7
// positions (bases, one separator, padding to a multiple of 8), so a group of 8 positions never
// straddles two reads; everything from the separator on is invalid with base code 0; the tail up to
// padded_positions() is invalid.  A thread produces one 32-position chunk: 8 bytes of bases + 4 bytes
// of mask, stored coalesced.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "kmer_math.h"

namespace ntsm {

struct DevPackParams {
	const uint8_t *ascii;      // device copy of the reads' bytes
	// fixed-length form: read r = ascii[r * stride, r * stride + read_len)
	uint32_t read_len, stride, span;     // span = read_span(read_len)
	uint32_t groups_per_read;            // span / 8
	// variable-length form (in_off != nullptr): read r = ascii[in_off[r], in_off[r + 1]), its first
	// position is out_pos[r] (prefix sums of read_span, computed by the feeder thread on the host)
	const uint32_t *in_off;
	const uint32_t *out_pos;
	uint32_t n_reads;
	uint64_t n_pos;            // data positions = out_pos[n_reads] (fixed: n_reads * span)
	uint64_t n_chunks_out;     // chunks to write: padded_positions(n_pos) / 32
	uint2 *bases;
	uint32_t *mask;
};

// One group of 8 positions of read bytes s[j .. j + 8) clipped to the read's length -> 16 bits of base
// codes + 8 invalid flags.
__device__ __forceinline__ void pack_group8(const uint8_t *__restrict__ s, uint32_t j, uint32_t len, const uint8_t *lut,
                                            uint32_t &bb, uint32_t &mm)
{
	bb = 0;
	mm = 0;
#pragma unroll
	for (uint32_t t = 0; t < 8; ++t) {
		uint32_t c = 4;
		if (j + t < len) c = lut[s[j + t]];
		bb |= (c & 3u) << (2 * t);
		mm |= (c >> 2) << t;
	}
}

template <bool FIXED>
__global__ void __launch_bounds__(256) pack_ascii_kernel(const DevPackParams P)
{
	__shared__ uint8_t lut[256];
	lut[threadIdx.x] = (uint8_t)nt4((unsigned char)threadIdx.x);     // vendor/KseqHashIterator.hpp:114-127
	__syncthreads();
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < P.n_chunks_out; c += stride) {
		uint32_t b[2] = { 0, 0 }, m = 0xFFFFFFFFu;
		const uint64_t p0 = c * 32;
		if (p0 < P.n_pos) {
			m = 0;
			uint32_t r = 0;
			if (!FIXED) {
				// the read that owns position p0: last r with out_pos[r] <= p0
				uint32_t lo = 0, hi = P.n_reads;               // invariant: out_pos[lo] <= p0 < out_pos[hi]
				while (hi - lo > 1) {
					const uint32_t mid = (lo + hi) >> 1;
					if ((uint64_t)__ldg(P.out_pos + mid) <= p0) lo = mid;
					else hi = mid;
				}
				r = lo;
			}
#pragma unroll
			for (uint32_t q = 0; q < 4; ++q) {
				const uint64_t p = p0 + 8 * q;
				uint32_t bb = 0, mm = 0xFFu;
				if (p < P.n_pos) {
					if (FIXED) {
						const uint64_t g = p >> 3;
						const uint32_t rr = (uint32_t)(g / P.groups_per_read);
						const uint32_t j = (uint32_t)(g - (uint64_t)rr * P.groups_per_read) * 8;
						pack_group8(P.ascii + (uint64_t)rr * P.stride, j, P.read_len, lut, bb, mm);
					} else {
						while ((uint64_t)__ldg(P.out_pos + r + 1) <= p) ++r;
						const uint32_t a = __ldg(P.in_off + r), len = __ldg(P.in_off + r + 1) - a;
						pack_group8(P.ascii + a, (uint32_t)(p - __ldg(P.out_pos + r)), len, lut, bb, mm);
					}
				}
				b[q >> 1] |= bb << (16 * (q & 1));
				m |= mm << (8 * q);
			}
		}
		P.bases[c] = make_uint2(b[0], b[1]);
		P.mask[c] = m;
	}
}

}  // namespace ntsm
