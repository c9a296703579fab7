// procpipe.h -- the shared-memory protocol between ntsm_count_files (the process that owns the CUDA contexts)
// and its parser WORKER PROCESSES (bin/ntsm_parse_worker).
//
// Why processes: a plain FASTQ file is parsed fastest in place through a mapping of the page cache (no read()
// copy: 4.6 against 2.8 Gbases/s per parser), but setting page tables up for 4 KiB pages does not scale across
// the threads of ONE address space -- 16 mapped parser threads reach 27 Gbases/s, 16 `gzread` threads 40, 16
// mapped parser PROCESSES 52 (profiles/r02f_hostpath.txt, r02r_hostpath_procs.txt).  So beyond a few parser
// threads the file pipeline (FingerPrint::computeCounts, src/FingerPrint.hpp:46-87) runs its parsers as processes:
// each maps its files, parses and packs exactly as the parser threads do, into batch slots that live in one
// shared mapping which the owner has page-locked (cudaHostRegister) -- the owner only issues the H2D copies and
// the kernels.  No CUDA in the workers; nothing but packed reads crosses the process boundary.
#pragma once
#include <stdint.h>

#include <atomic>
#include <new>

namespace ntsm {

constexpr uint32_t kProcMagic = 0x4E54534Du;   // "NTSM"
constexpr uint32_t kProcVersion = 1;

enum ProcSlotState : uint32_t { kSlotFree = 0, kSlotFilling = 1, kSlotReady = 2, kSlotInflight = 3 };

struct alignas(64) ProcSlot {
	std::atomic<uint32_t> state;
	uint32_t worker;
	uint64_t n_pos, n_bases, n_reads;     // valid when state == kSlotReady
};

struct alignas(64) ProcHeader {
	uint32_t magic, version;
	uint32_t k, verbose;
	uint32_t n_slots, n_files, n_workers, pad0;
	uint64_t cap_pos;                      // positions per batch slot (a multiple of 8)
	uint64_t slots_off;                    // ProcSlot[n_slots]
	uint64_t paths_off;                    // uint64 offset[n_files] (from the start of the mapping), then the strings
	uint64_t data_off, slot_stride;        // slot i: bases at data_off + i * slot_stride, mask at + bases_bytes
	uint64_t bases_bytes, mask_bytes;
	uint64_t total_bytes;
	alignas(64) std::atomic<uint32_t> next_file;
	std::atomic<uint32_t> stop;            // the owner gave up (error elsewhere): workers leave
	std::atomic<int32_t> error;            // first NTSM_ERR_* a worker met; text below
	std::atomic<uint32_t> error_lock;
	char error_text[480];
};

// ---- owner-side helpers (pipeline.cpp; tools/procpipe_selftest.cpp drives the same code without a GPU) ----
struct ProcLayout {
	uint64_t bases_bytes, mask_bytes, slot_stride, slots_off, paths_off, data_off, total;
};
constexpr uint64_t kProcPathBytes = 1u << 20;      // room for the paths of one call

inline ProcLayout proc_layout(uint64_t cap_pos, uint32_t n_slots)
{
	ProcLayout L;
	const uint64_t padded = (cap_pos + 8191) / 8192 * 8192 + 64;          // padded_positions (pack.h)
	L.bases_bytes = (padded / 32 * 8 + 63) & ~63ull;
	L.mask_bytes = (padded / 32 * 4 + 63) & ~63ull;
	L.slot_stride = L.bases_bytes + L.mask_bytes;
	L.slots_off = (sizeof(ProcHeader) + 63) & ~63ull;
	L.paths_off = L.slots_off + (uint64_t)n_slots * sizeof(ProcSlot);
	L.data_off = (L.paths_off + kProcPathBytes + 4095) & ~4095ull;
	L.total = L.data_off + L.slot_stride * n_slots;
	return L;
}

// fresh header + free slots + the call's paths; false if the paths do not fit
inline bool proc_init(uint8_t *base, const ProcLayout &L, uint64_t cap_pos, uint32_t n_slots, uint32_t k, uint32_t verbose,
                      const char *const *paths, uint32_t n_paths, uint32_t n_workers)
{
	ProcHeader *h = new (base) ProcHeader();
	h->magic = kProcMagic; h->version = kProcVersion;
	h->k = k; h->verbose = verbose;
	h->n_slots = n_slots; h->n_files = n_paths; h->n_workers = n_workers; h->pad0 = 0;
	h->cap_pos = cap_pos;
	h->slots_off = L.slots_off; h->paths_off = L.paths_off; h->data_off = L.data_off; h->slot_stride = L.slot_stride;
	h->bases_bytes = L.bases_bytes; h->mask_bytes = L.mask_bytes; h->total_bytes = L.total;
	h->next_file.store(0); h->stop.store(0); h->error.store(0); h->error_lock.store(0);
	h->error_text[0] = 0;
	ProcSlot *slots = reinterpret_cast<ProcSlot *>(base + L.slots_off);
	for (uint32_t i = 0; i < n_slots; ++i) {
		slots[i].state.store(kSlotFree);
		slots[i].worker = 0;
		slots[i].n_pos = slots[i].n_bases = slots[i].n_reads = 0;
	}
	uint64_t *off = reinterpret_cast<uint64_t *>(base + L.paths_off);
	uint64_t at = L.paths_off + (uint64_t)n_paths * 8;
	for (uint32_t i = 0; i < n_paths; ++i) {
		uint64_t len = 0;
		while (paths[i][len]) ++len;
		++len;
		if (at + len > L.data_off) return false;
		off[i] = at;
		for (uint64_t j = 0; j < len; ++j) base[at + j] = (uint8_t)paths[i][j];
		at += len;
	}
	std::atomic_thread_fence(std::memory_order_seq_cst);
	return true;
}

}  // namespace ntsm
