// inflate.h -- DEFLATE (RFC 1951) decoder for the ingest path, written for the one case that matters
// here: gigabytes of gzip'd FASTQ flowing through a parser thread that is otherwise ~13x faster
// than zlib's inflate (2.1 Gbases/s parse + pack against 0.155 Gbases/s gzread per thread).
//
// The reference reads every input through zlib's gzread (src/FingerPrint.hpp:50,
// vendor/kseq.h:68-79 via KSEQ_INIT(gzFile, gzread)); what must be reproduced is the BYTE STREAM
// gzread delivers, nothing else.  This decoder therefore only has to be right on valid streams and
// to notice, without ever touching memory it does not own, when a stream is not valid: every
// irregularity (bad header, over-subscribed code, distance too far back, truncated input, CRC or
// length mismatch) makes it stop with an error, and the caller (gzsource.cpp) hands the rest of the
// file -- from the start of the gzip member in question -- to zlib itself, so error behaviour is
// zlib's own by construction.
//
// Shape: 64-bit bit buffer refilled with one unaligned 8-byte load; 11-bit primary literal/length
// table and 9-bit primary distance table with second-level tables for longer codes; entries carry
// "bits to drop" and the base value so a symbol costs one lookup; literals decoded in runs per refill;
// matches copied 8 bytes at a time (distance 1 as a fill).  The decoder runs from a memory-mapped
// input straight into a caller-owned window and can stop at any symbol boundary when the window is
// full, so a multi-gigabyte member streams through a 1 MiB window.
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace ntsm {

class Inflater {
public:
	enum Status { kNeedOutput, kStreamEnd, kError, kBlockBoundary };

	// start a new raw DEFLATE stream whose first byte is at `in`
	void begin(const uint8_t *in, const uint8_t *in_end);

	// Decode until `out` reaches `out_limit`, the stream ends, or something is wrong.
	//   hist:      oldest byte a match may refer to (start of this stream's output still in memory)
	//   out:       in/out, next byte to write
	//   out_limit: stop once *out >= out_limit; the buffer must extend kSlack bytes past out_limit
	Status run(const uint8_t *hist, uint8_t **out, uint8_t *out_limit);

	// ---- for decoding one stream on several threads (pargz.h) ----
	// start in the middle of a byte: the stream continues at bit `bitpos` counted from `base`
	void begin_bits(const uint8_t *base, uint64_t bitpos, const uint8_t *in_end);
	// make run()/run16() return kBlockBoundary in front of the first block header at or after `bitpos`
	void stop_at_block_boundary(const uint8_t *base, uint64_t bitpos)
	{
		stop_base_ = base;
		stop_bit_ = bitpos;
	}
	// next unread bit, counted from `base`
	uint64_t bit_position(const uint8_t *base) const { return (uint64_t)(in_ - base) * 8 - (uint64_t)bitsleft_; }
	// the same decoder writing 16-bit symbols: bytes as 0-255; a caller that does not know the 32 KiB
	// before its starting point pre-fills them with markers 256 + i and resolves what got copied later
	Status run16(const uint16_t *hist, uint16_t **out, uint16_t *out_limit);
	// cheap test whether a dynamic-Huffman block header can start at this bit (used to look for block starts)
	static bool plausible_dynamic_header(const uint8_t *base, uint64_t bitpos, const uint8_t *end);
	// first such bit in [from_bit, to_bit), or ~0 if there is none
	static uint64_t find_plausible_dynamic_header(const uint8_t *base, uint64_t from_bit, uint64_t to_bit, const uint8_t *end);

	// first input byte not consumed (valid after kStreamEnd: the stream is byte-aligned there)
	const uint8_t *in_pos() const { return in_; }
	const char *error() const { return err_; }

	static constexpr size_t kSlack = 512;      // one iteration writes at most 56 literals + 258 + 7 bytes past the limit

private:
	static constexpr int kLitBits = 11, kDistBits = 9, kPreBits = 7;
	static constexpr int kLitSize = (1 << kLitBits) + 288 * 16, kDistSize = (1 << kDistBits) + 32 * 64;

	template <bool SAFE, typename OutT> Status huffman_loop(const OutT *hist, OutT **out, OutT *out_limit);
	template <typename OutT> Status run_t(const OutT *hist, OutT **out, OutT *out_limit);
	bool read_block_header();
	bool read_dynamic_tables();
	void refill();
	bool need(int n);                          // at least n bits in the buffer (after a refill)?
	void byte_align_and_rewind();
	Status fail(const char *why) { err_ = why; return kError; }

	const uint8_t *in_ = nullptr, *in_end_ = nullptr;
	uint64_t bitbuf_ = 0;
	int bitsleft_ = 0;
	enum { kBlockHeader, kStored, kHuffman, kDone } state_ = kBlockHeader;
	bool final_ = false;
	uint32_t stored_left_ = 0;
	const uint32_t *lit_ = nullptr, *dist_ = nullptr;   // tables of the current block
	const char *err_ = "";
	const uint8_t *stop_base_ = nullptr;
	uint64_t stop_bit_ = ~0ull;
	uint32_t lit_dyn_[kLitSize], dist_dyn_[kDistSize];
};

// Canonical Huffman decode table over code lengths lens[0..n): primary index = the first `bits` bits
// of the code as transmitted.  Returns false for an over-subscribed set, or an incomplete one that
// zlib would also refuse (anything but a single 1-bit code, or -- for distances -- no code at all).
// kind: 0 = literal/length alphabet, 1 = distance alphabet, 2 = code-length alphabet.
bool build_decode_table(uint32_t *table, int bits, const uint8_t *lens, int n, int kind);

}  // namespace ntsm
