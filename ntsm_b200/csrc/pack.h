// pack.h -- ASCII reads -> the packed "2-bit + N-mask" position stream (layout: include/ntsm_b200.h).
//
// Replaces the per-byte table lookup inside KseqHashIterator::step
// (vendor/KseqHashIterator.hpp:96-97,114-127): the decode happens once, on the host parse
// thread, and the GPU sees 3 bits per base.
#pragma once
#include <stdint.h>
#include <string.h>

namespace ntsm {

constexpr uint64_t kTilePositions = 8192;   // positions one CTA pass covers (256 threads x 32)
constexpr uint64_t kHaloPositions = 64;     // readable slack after the last tile

inline uint64_t padded_positions(uint64_t n_pos)
{
	return (n_pos + kTilePositions - 1) / kTilePositions * kTilePositions + kHaloPositions;
}

const uint8_t *code_table();
const char *pack_isa();                      // "avx512vbmi" | "avx2" | "scalar": the packer picked at start-up
void pack_reselect();                        // re-read NTSM_PACK_ISA (tests)                 // byte -> 0..3, or 4 for "not a base" (256 entries)

// Every read starts at a multiple of kReadAlign positions: after its bases come the separator and
// 0-7 more invalid positions.  Both planes are then byte-granular at every read start (2 bits x 8
// = two bytes of bases, one byte of mask), so the packers write whole vector results with plain
// unaligned stores -- no bit offsets, no read-modify-write -- at the price of ~0.7 % more
// positions for 150-base reads.
constexpr uint64_t kReadAlign = 8;
inline uint64_t read_span(uint64_t n_bases) { return (n_bases + kReadAlign) & ~(kReadAlign - 1); }   // bases + separator + padding

// Streaming packer over caller-owned word arrays.  Stores may run up to 63 positions past `pos`
// (always invalid positions); the arrays must be sized with padded_positions().
//
// Two output modes.  reset(): the packers' vector stores go straight to the arrays (any alignment).
// reset_streaming(): for the PINNED batch buffers, which no host core reads again -- the packers write
// into a small staging area that stays in L1/L2, and completed 512-position blocks leave it as whole
// 64-byte lines with non-temporal stores.  Ordinary stores to the pinned buffer cost a read-for-
// ownership of every line before it is overwritten (0.375 B/base of DRAM reads for nothing, a sixth
// of the packer's memory traffic, and the packers are DRAM-bound: DESIGN 5); streaming stores do not.
struct Packer {
	uint64_t *bases = nullptr;   // 32 positions per word (little endian == two uint32 of 16)
	uint32_t *mask = nullptr;    // 32 positions per word
	uint64_t pos = 0;            // next stream position, always a multiple of kReadAlign

	static constexpr uint64_t kStagePos = 1u << 16;       // staging capacity in positions (16 KiB of bases + 8 KiB of mask)
	static constexpr uint64_t kBlockPos = 512;            // flush granularity: 128 B of bases, 64 B of mask
	static constexpr uint64_t kSlackPos = 1024;           // room kept free: the packers' overrun + the padding of the last block
	bool streaming = false;
	uint64_t origin = 0;         // stream position of the staging area's first byte (a multiple of kBlockPos)
	uint8_t *wb = nullptr, *wm = nullptr;    // where position 0 WOULD be written: the packers store at wb + pos / 4, wm + pos / 8
	alignas(64) uint8_t stage_b[kStagePos / 4 + 64];
	alignas(64) uint8_t stage_m[kStagePos / 8 + 64];

	void reset(uint64_t *b, uint32_t *m)
	{
		bases = b; mask = m; pos = 0; streaming = false;
		wb = reinterpret_cast<uint8_t *>(b); wm = reinterpret_cast<uint8_t *>(m);
	}
	void reset_streaming(uint64_t *b, uint32_t *m)        // b and m 64-byte aligned
	{
		bases = b; mask = m; pos = 0; streaming = true; origin = 0;
		wb = stage_b; wm = stage_m;
	}

	void put_read(const char *s, uint64_t n);           // n bases, the separator, padding: read_span(n) positions

	// pad with invalid positions up to padded_positions(pos); returns the data length (pos before padding)
	uint64_t finish();

private:
	void flush_blocks(bool all);
};

}  // namespace ntsm
