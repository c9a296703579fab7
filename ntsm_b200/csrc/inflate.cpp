// inflate.cpp -- see inflate.h.  Format per RFC 1951; which malformed inputs are refused follows
// zlib's inflate (the reference's reader, vendor/kseq.h:68-79 over gzread), because on refusal the
// caller replays the gzip member through zlib and zlib's verdict becomes ours.
#include "inflate.h"

#include <string.h>

namespace ntsm {

namespace {

// table entry: bits 0-7 bits to drop at this stage (code + extra bits), 8-11 code bits at this stage,
// 12-27 payload (literal, base value, or second-level table start), 28-31 kind.  0 = no such code.
constexpr uint32_t kLen = 1u << 28, kSub = 1u << 29, kEob = 1u << 30, kLit = 1u << 31;

const uint16_t kLenBase[29] = { 3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258 };
const uint8_t kLenExtra[29] = { 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0 };
const uint16_t kDistBase[30] = { 1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577 };
const uint8_t kDistExtra[30] = { 0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13 };
const uint8_t kPrecodeOrder[19] = { 16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15 };

inline uint32_t make_entry(int kind, int sym, int cw)
{
	const uint32_t c = (uint32_t)cw;
	if (kind == 0) {
		if (sym < 256) return kLit | ((uint32_t)sym << 12) | (c << 8) | c;
		if (sym == 256) return kEob | (c << 8) | c;
		if (sym <= 285) return kLen | ((uint32_t)kLenBase[sym - 257] << 12) | (c << 8) | (c + kLenExtra[sym - 257]);
		return 0;                                     // 286, 287: take code space, never valid in data
	}
	if (kind == 1) return sym < 30 ? kLen | ((uint32_t)kDistBase[sym] << 12) | (c << 8) | (c + kDistExtra[sym]) : 0;
	return kLit | ((uint32_t)sym << 12) | (c << 8) | c;
}

inline uint32_t bit_reverse(uint32_t v, int n)
{
	uint32_t r = 0;
	for (int i = 0; i < n; ++i) {
		r = (r << 1) | (v & 1);
		v >>= 1;
	}
	return r;
}

inline uint64_t load64(const uint8_t *p)
{
	uint64_t v;
	memcpy(&v, p, 8);
	return v;                                             // x86-64: little endian
}

struct FixedTables {
	uint32_t lit[(1 << 11)], dist[(1 << 9)];
	FixedTables()
	{
		uint8_t l[288];
		for (int i = 0; i < 288; ++i) l[i] = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8;
		build_decode_table(lit, 11, l, 288, 0);
		uint8_t d[32];
		for (int i = 0; i < 32; ++i) d[i] = 5;              // 30 and 31 complete the code but are not valid symbols
		build_decode_table(dist, 9, d, 32, 1);
	}
};

}  // namespace

bool build_decode_table(uint32_t *table, int bits, const uint8_t *lens, int n, int kind)
{
	int count[16] = { 0 };
	for (int i = 0; i < n; ++i) count[lens[i] & 15]++;
	int max_len = 15;
	while (max_len > 0 && !count[max_len]) --max_len;
	memset(table, 0, sizeof(uint32_t) << bits);
	if (max_len == 0) return kind == 1;                     // a block of literals only may define no distance code
	int left = 1;
	for (int l = 1; l <= 15; ++l) {
		left = (left << 1) - count[l];
		if (left < 0) return false;                         // over-subscribed
	}
	if (left > 0 && (kind == 2 || max_len != 1)) return false;   // incomplete (zlib: inftrees.c, same rule)

	uint32_t next[16];
	uint32_t code = 0;
	count[0] = 0;
	for (int l = 1; l <= 15; ++l) {
		code = (code + (uint32_t)count[l - 1]) << 1;
		next[l] = code;
	}
	uint16_t rev[288];
	uint8_t submax[1 << 11];
	const uint32_t pmask = (1u << bits) - 1;
	if (max_len > bits) memset(submax, 0, (size_t)1 << bits);
	for (int s = 0; s < n; ++s) {
		const int l = lens[s];
		if (!l) continue;
		const uint32_t r = bit_reverse(next[l]++, l);
		rev[s] = (uint16_t)r;
		if (l > bits && submax[r & pmask] < l) submax[r & pmask] = (uint8_t)l;
	}
	uint32_t next_free = 1u << bits;
	for (int s = 0; s < n; ++s) {
		const int l = lens[s];
		if (!l) continue;
		const uint32_t r = rev[s];
		if (l <= bits) {
			const uint32_t e = make_entry(kind, s, l);
			for (uint32_t i = r; i <= pmask; i += 1u << l) table[i] = e;
		} else {
			const uint32_t p = r & pmask;
			if (!(table[p] & kSub)) {
				const uint32_t sb = (uint32_t)submax[p] - (uint32_t)bits;
				table[p] = kSub | (next_free << 12) | (sb << 8) | (uint32_t)bits;
				memset(table + next_free, 0, sizeof(uint32_t) << sb);
				next_free += 1u << sb;
			}
			const uint32_t sb = (table[p] >> 8) & 15, start = (table[p] >> 12) & 0xFFFF;
			const uint32_t e = make_entry(kind, s, l - bits);
			for (uint32_t i = r >> bits; i < (1u << sb); i += 1u << (l - bits)) table[start + i] = e;
		}
	}
	return true;
}

void Inflater::begin(const uint8_t *in, const uint8_t *in_end)
{
	in_ = in;
	in_end_ = in_end;
	bitbuf_ = 0;
	bitsleft_ = 0;
	state_ = kBlockHeader;
	final_ = false;
	stored_left_ = 0;
	err_ = "";
	stop_base_ = nullptr;
	stop_bit_ = ~0ull;
}

void Inflater::begin_bits(const uint8_t *base, uint64_t bitpos, const uint8_t *in_end)
{
	begin(base + bitpos / 8, in_end);
	const unsigned skip = (unsigned)(bitpos & 7);
	if (skip && in_ < in_end_) {                            // the rest of the byte the stream continues in
		bitbuf_ = (uint64_t)*in_++ >> skip;
		bitsleft_ = 8 - (int)skip;
	}
}

// BTYPE = 2, HLIT <= 29, HDIST <= 29, and the code-length code complete (Kraft sum exactly 1): a few
// shifts and four table lookups per bit position, passed by roughly one random position in a thousand.
namespace {
struct KraftLut {
	uint16_t sum[1 << 15];                                  // five 3-bit code lengths -> their Kraft sum in 1/128ths
	KraftLut()
	{
		for (uint32_t v = 0; v < (1u << 15); ++v) {
			unsigned k = 0;
			for (int i = 0; i < 5; ++i) {
				const unsigned l = (v >> (3 * i)) & 7;
				if (l) k += 128u >> l;
			}
			sum[v] = (uint16_t)k;
		}
	}
};
}  // namespace

uint64_t Inflater::find_plausible_dynamic_header(const uint8_t *base, uint64_t from_bit, uint64_t to_bit, const uint8_t *end)
{
	static const KraftLut lut;
	for (uint64_t bit = from_bit; bit < to_bit; ++bit) {
		const uint8_t *p = base + bit / 8;
		if (end - p < 24) return ~0ull;
		const unsigned sh = (unsigned)(bit & 7);
		const uint64_t lo = load64(p) >> sh;                // >= 56 bits
		if (((lo >> 1) & 3) != 2) continue;
		if (((lo >> 3) & 31) > 29 || ((lo >> 8) & 31) > 29) continue;
		const unsigned hclen = (unsigned)((lo >> 13) & 15) + 4;
		// the code-length code's lengths follow at bit 17: up to 57 bits
		uint64_t v = (load64(p + 2) >> (sh + 1));           // bits from 17 on: >= 55 of them
		v |= (uint64_t)p[10] << (63 - sh);                  // top them up to 64 - ... >= 63 bits
		if (hclen < 19) v &= (1ull << (3 * hclen)) - 1;
		else v &= (1ull << 57) - 1;
		const unsigned kraft = lut.sum[v & 0x7FFF] + lut.sum[(v >> 15) & 0x7FFF] + lut.sum[(v >> 30) & 0x7FFF] + lut.sum[(v >> 45) & 0x7FFF];
		if (kraft == 128) return bit;
	}
	return ~0ull;
}

bool Inflater::plausible_dynamic_header(const uint8_t *base, uint64_t bitpos, const uint8_t *end)
{
	return find_plausible_dynamic_header(base, bitpos, bitpos + 1, end) == bitpos;
}

// Invariant of the bit buffer: bits [0, bitsleft_) of bitbuf_ are the next unread bits of the stream
// and in_ points at the byte right after them.  Bits above bitsleft_ may hold bytes that were loaded
// early; they are the true next bytes of the stream, so ORing them in again later changes nothing.
inline void Inflater::refill()
{
	if (in_end_ - in_ >= 8) {
		bitbuf_ |= load64(in_) << bitsleft_;
		in_ += (63 - bitsleft_) >> 3;
		bitsleft_ |= 56;
	} else {
		while (bitsleft_ <= 56 && in_ < in_end_) {
			bitbuf_ |= (uint64_t)*in_++ << bitsleft_;
			bitsleft_ += 8;
		}
	}
}

inline bool Inflater::need(int n)
{
	if (bitsleft_ < n) refill();
	return bitsleft_ >= n;
}

void Inflater::byte_align_and_rewind()
{
	bitsleft_ -= bitsleft_ & 7;
	in_ -= bitsleft_ >> 3;
	bitbuf_ = 0;
	bitsleft_ = 0;
}

bool Inflater::read_dynamic_tables()
{
	if (!need(14)) return fail("truncated block header"), false;
	const int hlit = (int)(bitbuf_ & 31) + 257, hdist = (int)((bitbuf_ >> 5) & 31) + 1, hclen = (int)((bitbuf_ >> 10) & 15) + 4;
	bitbuf_ >>= 14;
	bitsleft_ -= 14;
	if (hlit > 286 || hdist > 30) return fail("too many length or distance symbols"), false;
	uint8_t pre[19] = { 0 };
	for (int i = 0; i < hclen; ++i) {
		if (!need(3)) return fail("truncated block header"), false;
		pre[kPrecodeOrder[i]] = (uint8_t)(bitbuf_ & 7);
		bitbuf_ >>= 3;
		bitsleft_ -= 3;
	}
	uint32_t ptab[1 << kPreBits];
	if (!build_decode_table(ptab, kPreBits, pre, 19, 2)) return fail("invalid code lengths set"), false;
	uint8_t lens[286 + 30 + 138];
	const int total = hlit + hdist;
	int i = 0;
	while (i < total) {
		refill();
		const uint32_t e = ptab[bitbuf_ & ((1u << kPreBits) - 1)];
		if (!e) return fail("invalid code lengths set"), false;
		bitbuf_ >>= e & 0xFF;
		bitsleft_ -= (int)(e & 0xFF);
		const int sym = (int)((e >> 12) & 0xFFFF);
		int rep;
		uint8_t val = 0;
		if (sym < 16) {
			lens[i++] = (uint8_t)sym;
			if (bitsleft_ < 0) return fail("truncated block header"), false;
			continue;
		}
		if (sym == 16) {
			if (i == 0) return fail("invalid bit length repeat"), false;
			val = lens[i - 1];
			rep = 3 + (int)(bitbuf_ & 3);
			bitbuf_ >>= 2;
			bitsleft_ -= 2;
		} else if (sym == 17) {
			rep = 3 + (int)(bitbuf_ & 7);
			bitbuf_ >>= 3;
			bitsleft_ -= 3;
		} else {
			rep = 11 + (int)(bitbuf_ & 127);
			bitbuf_ >>= 7;
			bitsleft_ -= 7;
		}
		if (bitsleft_ < 0) return fail("truncated block header"), false;
		if (i + rep > total) return fail("invalid bit length repeat"), false;
		memset(lens + i, val, (size_t)rep);
		i += rep;
	}
	if (lens[256] == 0) return fail("invalid code -- missing end-of-block"), false;
	if (!build_decode_table(lit_dyn_, kLitBits, lens, hlit, 0)) return fail("invalid literal/lengths set"), false;
	if (!build_decode_table(dist_dyn_, kDistBits, lens + hlit, hdist, 1)) return fail("invalid distances set"), false;
	lit_ = lit_dyn_;
	dist_ = dist_dyn_;
	return true;
}

bool Inflater::read_block_header()
{
	if (!need(3)) return fail("truncated stream"), false;
	final_ = bitbuf_ & 1;
	const int type = (int)((bitbuf_ >> 1) & 3);
	bitbuf_ >>= 3;
	bitsleft_ -= 3;
	if (type == 0) {
		byte_align_and_rewind();
		if (in_end_ - in_ < 4) return fail("truncated stored block"), false;
		const uint32_t len = in_[0] | ((uint32_t)in_[1] << 8), nlen = in_[2] | ((uint32_t)in_[3] << 8);
		if ((len ^ 0xFFFFu) != nlen) return fail("invalid stored block lengths"), false;
		in_ += 4;
		stored_left_ = len;
		state_ = kStored;
		return true;
	}
	if (type == 1) {
		static const FixedTables fixed;
		lit_ = fixed.lit;
		dist_ = fixed.dist;
		state_ = kHuffman;
		return true;
	}
	if (type == 2) {
		if (!read_dynamic_tables()) return false;
		state_ = kHuffman;
		return true;
	}
	return fail("invalid block type"), false;
}

// The symbol loop.  SAFE = false: at least 16 input bytes remain at every refill, so the buffer never
// runs dry and no per-symbol check is needed; SAFE = true: the last bytes of the input, same code plus
// the check that no more bits were consumed than the input had.
template <bool SAFE, typename OutT> Inflater::Status Inflater::huffman_loop(const OutT *hist, OutT **outp, OutT *out_limit)
{
	OutT *out = *outp;
	const uint8_t *in = in_;
	uint64_t bb = bitbuf_;
	int bl = bitsleft_;
	const uint32_t *const lit = lit_, *const dist = dist_;
	const uint8_t *const in_end = in_end_;
	Status st = kNeedOutput;
	const char *why = nullptr;

#define NTSM_REFILL()                                                      \
	do {                                                                   \
		if (!SAFE || in_end - in >= 8) {                                   \
			bb |= load64(in) << bl;                                        \
			in += (63 - bl) >> 3;                                          \
			bl |= 56;                                                      \
		} else {                                                           \
			while (bl <= 56 && in < in_end) {                              \
				bb |= (uint64_t)*in++ << bl;                               \
				bl += 8;                                                   \
			}                                                              \
		}                                                                  \
	} while (0)
#define NTSM_LOOKUP(e, table, tbits)                                       \
	do {                                                                   \
		e = table[bb & ((1u << (tbits)) - 1)];                             \
		if (e & kSub) {                                                    \
			bb >>= (tbits);                                                \
			bl -= (tbits);                                                 \
			e = table[((e >> 12) & 0xFFFF) + (bb & ((1u << ((e >> 8) & 15)) - 1))]; \
		}                                                                  \
	} while (0)

	for (;;) {
		if (out >= out_limit) break;
		if (!SAFE && in_end - in < 16) break;                 // the caller continues in the careful loop
		NTSM_REFILL();
		uint32_t e;
		NTSM_LOOKUP(e, lit, kLitBits);
		if (e & kLit) {
			// a run of literals: a lookup needs at most 15 bits, so keep going while that many are left
			do {
				bb >>= e & 0xFF;
				bl -= (int)(e & 0xFF);
				*out++ = (OutT)((e >> 12) & 0xFF);
				if (bl < 15) break;
				NTSM_LOOKUP(e, lit, kLitBits);
			} while (e & kLit);
			if (SAFE && bl < 0) { why = "truncated stream"; st = kError; break; }
			if (e & kLit) continue;                            // out of bits after a literal: refill at the top
			NTSM_REFILL();                                     // e is looked up but not dropped yet; a match needs up to 48 bits
		}
		if (e & kEob) {
			bb >>= e & 0xFF;
			bl -= (int)(e & 0xFF);
			if (SAFE && bl < 0) { why = "truncated stream"; st = kError; break; }
			st = kStreamEnd;                                   // end of BLOCK; run() decides what follows
			break;
		}
		if (!(e & kLen)) { why = "invalid literal/length code"; st = kError; break; }
		uint64_t saved = bb;
		bb >>= e & 0xFF;
		bl -= (int)(e & 0xFF);
		const uint32_t len = ((e >> 12) & 0xFFFF) + (uint32_t)((saved >> ((e >> 8) & 15)) & ((1u << ((e & 0xFF) - ((e >> 8) & 15))) - 1));
		uint32_t d;
		NTSM_LOOKUP(d, dist, kDistBits);
		if (!(d & kLen)) { why = "invalid distance code"; st = kError; break; }
		saved = bb;
		bb >>= d & 0xFF;
		bl -= (int)(d & 0xFF);
		if (SAFE && bl < 0) { why = "truncated stream"; st = kError; break; }
		const uint32_t off = ((d >> 12) & 0xFFFF) + (uint32_t)((saved >> ((d >> 8) & 15)) & ((1u << ((d & 0xFF) - ((d >> 8) & 15))) - 1));
		if (off > (size_t)(out - hist)) { why = "invalid distance too far back"; st = kError; break; }
		const OutT *src = out - off;
		OutT *const end = out + len;
		if (sizeof(OutT) == 1) {
			if (off >= 8) {
				do {
					memcpy(out, src, 8);
					out += 8;
					src += 8;
				} while (out < end);
			} else if (off == 1) {
				memset(out, (int)*src, len);
			} else {
				do { *out++ = *src++; } while (out < end);
			}
		} else {
			if (off >= 4) {                                    // four 16-bit symbols per move
				do {
					memcpy(out, src, 8);
					out += 4;
					src += 4;
				} while (out < end);
			} else {
				do { *out++ = *src++; } while (out < end);
			}
		}
		out = end;
	}
#undef NTSM_REFILL
#undef NTSM_LOOKUP
	*outp = out;
	in_ = in;
	bitbuf_ = bb;
	bitsleft_ = bl;
	if (st == kError) err_ = why;
	return st;
}

Inflater::Status Inflater::run(const uint8_t *hist, uint8_t **out, uint8_t *out_limit) { return run_t<uint8_t>(hist, out, out_limit); }
Inflater::Status Inflater::run16(const uint16_t *hist, uint16_t **out, uint16_t *out_limit) { return run_t<uint16_t>(hist, out, out_limit); }

template <typename OutT> Inflater::Status Inflater::run_t(const OutT *hist, OutT **out, OutT *out_limit)
{
	for (;;) {
		if (state_ == kDone) return kStreamEnd;
		if (*out >= out_limit) return kNeedOutput;
		if (state_ == kBlockHeader) {
			if (stop_bit_ != ~0ull && bit_position(stop_base_) >= stop_bit_) return kBlockBoundary;
			if (!read_block_header()) return kError;
			continue;
		}
		if (state_ == kStored) {
			size_t n = stored_left_;
			if (n > (size_t)(out_limit - *out)) n = (size_t)(out_limit - *out);
			if (n > (size_t)(in_end_ - in_)) return fail("truncated stored block");
			if (sizeof(OutT) == 1) memcpy(*out, in_, n);
			else
				for (size_t i = 0; i < n; ++i) (*out)[i] = (OutT)in_[i];
			*out += n;
			in_ += n;
			stored_left_ -= (uint32_t)n;
			if (stored_left_) return kNeedOutput;
			state_ = final_ ? kDone : kBlockHeader;
			continue;
		}
		// Huffman block: the fast loop while plenty of input remains, the careful one for the tail
		Status st = in_end_ - in_ >= 16 ? huffman_loop<false, OutT>(hist, out, out_limit) : huffman_loop<true, OutT>(hist, out, out_limit);
		if (st == kError) return kError;
		if (st == kStreamEnd) {                                 // end-of-block symbol
			state_ = final_ ? kDone : kBlockHeader;
			if (final_) byte_align_and_rewind();
			continue;
		}
		// kNeedOutput from the fast loop can also mean "input is getting short": loop and re-dispatch
	}
}

}  // namespace ntsm
