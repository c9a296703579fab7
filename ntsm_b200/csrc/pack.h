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

// Streaming bit packer over caller-owned word arrays.
struct Packer {
	uint64_t *bases = nullptr;   // 32 positions per word (little endian == two uint32 of 16)
	uint32_t *mask = nullptr;    // 32 positions per word
	uint64_t pos = 0;            // next stream position
	uint64_t bacc = 0;           // partial words for positions [pos & ~31, pos)
	uint32_t macc = 0;

	void reset(uint64_t *b, uint32_t *m) { bases = b; mask = m; pos = 0; bacc = 0; macc = 0; }

	inline void put_code(unsigned code)
	{
		const unsigned sh = (unsigned)pos & 31;
		bacc |= (uint64_t)(code & 3) << (2 * sh);
		macc |= (uint32_t)(code >> 2) << sh;
		if (sh == 31) {
			bases[pos >> 5] = bacc;
			mask[pos >> 5] = macc;
			bacc = 0;
			macc = 0;
		}
		++pos;
	}

	// n positions (1..32) at once: b = their 2-bit codes (low 2n bits, rest zero), m = invalid flags
	inline void put_group(uint64_t b, uint32_t m, unsigned n)
	{
		const unsigned sh = (unsigned)pos & 31;
		bacc |= b << (2 * sh);
		macc |= m << sh;
		if (sh + n >= 32) {
			bases[pos >> 5] = bacc;
			mask[pos >> 5] = macc;
			const unsigned used = 32 - sh;               // positions of this group that went into the flushed word
			bacc = used < 32 ? b >> (2 * used) : 0;
			macc = used < 32 ? m >> used : 0;
		}
		pos += n;
	}

	void put_bases(const char *s, uint64_t n);          // decode + pack n bytes
	void put_read(const char *s, uint64_t n);           // n bytes + the separator position
	inline void put_separator() { put_code(4); }

	// pad with invalid positions up to padded_positions(pos); returns the data length (pos before padding)
	uint64_t finish();
};

}  // namespace ntsm
