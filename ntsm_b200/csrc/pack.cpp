// pack.cpp -- host packer (see pack.h).
#include "pack.h"

#include "../../include/ntsm_b200.h"
#include "kmer_math.h"

#include <stdlib.h>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace ntsm {

// vendor/KseqHashIterator.hpp:114-127 as a flat table (built once from nt4()).
struct CodeTableInit {
	uint8_t t[256];
	CodeTableInit() { for (int i = 0; i < 256; ++i) t[i] = (uint8_t)nt4((unsigned char)i); }
};
static const CodeTableInit g_code_init;
const uint8_t *code_table() { return g_code_init.t; }

// One read of n bytes appended at dst.pos (a multiple of 8): its bases, then invalid positions up to
// read_span(n).  Three implementations of the same function, picked once at start-up: AVX-512 VBMI
// (64 bases per step: one byte permute through the 128-entry decode table, two multiply-adds to
// squeeze four codes into a byte), AVX2+BMI2 (32 per step: compares, pdep to interleave the
// bit-planes), scalar table walk.  All write whole groups; whatever a group covers beyond the
// read's span is invalid and is overwritten by the next read or by finish().
static void run_scalar(Packer &dst, const char *s, uint64_t n)
{
	const uint8_t *t = g_code_init.t;
	uint8_t *ob = dst.wb + dst.pos / 4;
	uint8_t *om = dst.wm + dst.pos / 8;
	const uint64_t span = read_span(n);
	for (uint64_t i = 0; i < span; i += 8) {
		unsigned bb = 0, mm = 0;
		for (unsigned j = 0; j < 8; ++j) {
			const unsigned c = i + j < n ? t[(unsigned char)s[i + j]] : 4u;
			bb |= (c & 3) << (2 * j);
			mm |= (c >> 2) << j;
		}
		ob[0] = (uint8_t)bb;
		ob[1] = (uint8_t)(bb >> 8);
		om[0] = (uint8_t)mm;
		ob += 2;
		om += 1;
	}
	dst.pos += span;
}

#if defined(__x86_64__)
#define NTSM_TGT_AVX2 __attribute__((target("avx2,bmi2")))
// Reads arrive back to back (a bulk of reads in host memory, or the reader's window), and a packer
// thread walks them once: the hardware prefetcher alone leaves it waiting on DRAM (measured: 27 ->
// 42 Gbases/s on 8 threads with a software prefetch 4 KiB ahead).  A prefetch never faults.
static constexpr size_t kPrefetchAhead = 4096;

#define NTSM_TGT_AVX512 __attribute__((target("avx512f,avx512bw,avx512vl,avx512vbmi,bmi2")))

// 32 ASCII bytes -> (bit-plane 0, bit-plane 1, valid flags).  Valid letters: A C G T U in either
// case (clearing bit 5 folds the case) and raw bytes 0..3 (table rows 0-3).
NTSM_TGT_AVX2 static inline void planes32_avx2(__m256i v, uint32_t *bit0, uint32_t *bit1, uint32_t *valid)
{
	const __m256i up = _mm256_and_si256(v, _mm256_set1_epi8((char)0xDF));
	const __m256i isA = _mm256_cmpeq_epi8(up, _mm256_set1_epi8('A'));
	const __m256i isC = _mm256_cmpeq_epi8(up, _mm256_set1_epi8('C'));
	const __m256i isG = _mm256_cmpeq_epi8(up, _mm256_set1_epi8('G'));
	const __m256i isT = _mm256_or_si256(_mm256_cmpeq_epi8(up, _mm256_set1_epi8('T')),
	                                    _mm256_cmpeq_epi8(up, _mm256_set1_epi8('U')));
	const __m256i isRaw = _mm256_cmpeq_epi8(_mm256_and_si256(v, _mm256_set1_epi8((char)0xFC)), _mm256_setzero_si256());
	const __m256i r1 = _mm256_and_si256(isRaw, _mm256_cmpeq_epi8(_mm256_and_si256(v, _mm256_set1_epi8(1)), _mm256_set1_epi8(1)));
	const __m256i r2 = _mm256_and_si256(isRaw, _mm256_cmpeq_epi8(_mm256_and_si256(v, _mm256_set1_epi8(2)), _mm256_set1_epi8(2)));
	*bit0 = (uint32_t)_mm256_movemask_epi8(_mm256_or_si256(_mm256_or_si256(isC, isT), r1));     // set for C, T
	*bit1 = (uint32_t)_mm256_movemask_epi8(_mm256_or_si256(_mm256_or_si256(isG, isT), r2));     // set for G, T
	*valid = (uint32_t)_mm256_movemask_epi8(
	    _mm256_or_si256(_mm256_or_si256(_mm256_or_si256(isA, isC), _mm256_or_si256(isG, isT)), isRaw));
}

NTSM_TGT_AVX2 static void run_avx2(Packer &dst, const char *s, uint64_t n)
{
	const uint64_t kEven = 0x5555555555555555ULL, kOdd = 0xAAAAAAAAAAAAAAAAULL;
	uint8_t *ob = dst.wb + dst.pos / 4;
	uint8_t *om = dst.wm + dst.pos / 8;
	uint64_t i = 0;
	uint32_t b0, b1, va;
	for (; i + 32 <= n; i += 32, ob += 8, om += 4) {
		if ((i & 32) == 0) _mm_prefetch(s + i + kPrefetchAhead, _MM_HINT_T0);
		planes32_avx2(_mm256_loadu_si256((const __m256i *)(s + i)), &b0, &b1, &va);
		const uint64_t bb = _pdep_u64(b0, kEven) | _pdep_u64(b1, kOdd);
		const uint32_t mm = ~va;
		memcpy(ob, &bb, 8);
		memcpy(om, &mm, 4);
	}
	const unsigned r = (unsigned)(n - i);           // 0..31 bases left; the separator and the padding ride in the same group
	// A 32-byte load that stays inside the 4 KiB page of its first byte cannot fault, so the bytes
	// after the read are loaded and masked off; only a load that would cross a page is bounced.
	__m256i v = _mm256_setzero_si256();
	if (r) {
		if (((uintptr_t)(s + i) & 4095u) <= 4096u - 32u) {
			v = _mm256_loadu_si256((const __m256i *)(s + i));
		} else {
			char tmp[32] = { 0 };
			memcpy(tmp, s + i, r);
			v = _mm256_loadu_si256((const __m256i *)tmp);
		}
	}
	planes32_avx2(v, &b0, &b1, &va);
	const uint32_t keep = (uint32_t)((1ull << r) - 1);
	const uint64_t bb = _pdep_u64(b0 & keep, kEven) | _pdep_u64(b1 & keep, kOdd);
	const uint32_t mm = ~(va & keep);               // everything from the separator on is invalid
	memcpy(ob, &bb, 8);
	memcpy(om, &mm, 4);
	dst.pos += read_span(n);
}

struct alignas(64) Vbmi128 {
	uint8_t t[128];
	Vbmi128() { for (int i = 0; i < 128; ++i) { const unsigned c = nt4((unsigned char)i); t[i] = c < 4 ? (uint8_t)c : 0x80; } }
};
static const Vbmi128 g_vbmi_tab;

// 64 ASCII bytes -> 16 bytes of 2-bit codes (four positions per byte) + 64 invalid flags
NTSM_TGT_AVX512 static inline __m128i pack64_avx512(__m512i v, __m512i tab_lo, __m512i tab_hi, uint64_t *inv)
{
	// index bits 0-6 pick one of 128 table bytes (bit 7 is ignored by the permute); bytes >= 0x80
	// are never bases, their own sign bit marks them invalid
	const __m512i code = _mm512_permutex2var_epi8(tab_lo, v, tab_hi);
	const __mmask64 bad = _mm512_movepi8_mask(_mm512_or_si512(code, v));
	const __m512i c = _mm512_maskz_mov_epi8(~bad, code);                                  // invalid positions carry code 0
	const __m512i t16 = _mm512_maddubs_epi16(c, _mm512_set1_epi16(0x0401));               // c0 + 4 c1 per 16-bit lane
	const __m512i t32 = _mm512_madd_epi16(t16, _mm512_set1_epi32(0x00100001));            // + 16 (c2 + 4 c3) per 32-bit lane
	*inv = (uint64_t)bad;
	return _mm512_cvtepi32_epi8(t32);
}

NTSM_TGT_AVX512 static void run_avx512(Packer &dst, const char *s, uint64_t n)
{
	const __m512i tab_lo = _mm512_load_si512(g_vbmi_tab.t), tab_hi = _mm512_load_si512(g_vbmi_tab.t + 64);
	uint8_t *ob = dst.wb + dst.pos / 4;
	uint8_t *om = dst.wm + dst.pos / 8;
	uint64_t i = 0, iv;
	for (; i + 64 <= n; i += 64, ob += 16, om += 8) {
		_mm_prefetch(s + i + kPrefetchAhead, _MM_HINT_T0);
		const __m128i bb = pack64_avx512(_mm512_loadu_si512(s + i), tab_lo, tab_hi, &iv);
		_mm_storeu_si128((__m128i *)ob, bb);
		memcpy(om, &iv, 8);
	}
	_mm_prefetch(s + i + kPrefetchAhead, _MM_HINT_T0);
	const unsigned r = (unsigned)(n - i);           // 0..63 bases left
	const uint64_t keep = (1ull << r) - 1;          // a masked load never touches (or faults on) the bytes it skips
	const __m128i bb = pack64_avx512(_mm512_maskz_loadu_epi8((__mmask64)keep, s + i), tab_lo, tab_hi, &iv);
	iv |= ~keep;                                    // byte 0 decodes as 'A': everything from the separator on is invalid
	_mm_storeu_si128((__m128i *)ob, bb);          // skipped bytes loaded as 0 = code 0
	memcpy(om, &iv, 8);
	dst.pos += read_span(n);
}
#endif

typedef void (*RunFn)(Packer &, const char *, uint64_t);
static RunFn pick_run()
{
	const char *force = getenv("NTSM_PACK_ISA");        // "scalar" | "avx2" | "avx512": tests walk all three
#if defined(__x86_64__)
	const bool avx2 = __builtin_cpu_supports("avx2") && __builtin_cpu_supports("bmi2");
	const bool avx512 = avx2 && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vl") &&
	                    __builtin_cpu_supports("avx512vbmi");
	if (force) {
		if (!strcmp(force, "scalar")) return run_scalar;
		if (!strcmp(force, "avx2") && avx2) return run_avx2;
		if (!strcmp(force, "avx512") && avx512) return run_avx512;
	}
	if (avx512) return run_avx512;
	if (avx2) return run_avx2;
#else
	(void)force;
#endif
	return run_scalar;
}
static RunFn g_run = pick_run();

const char *pack_isa()
{
#if defined(__x86_64__)
	if (g_run == run_avx512) return "avx512vbmi";
	if (g_run == run_avx2) return "avx2";
#endif
	return "scalar";
}
void pack_reselect() { g_run = pick_run(); }

// Completed blocks of the staging area -> the pinned arrays, as whole 64-byte lines that bypass the cache
// (no read-for-ownership); what is left (less than a block, plus the packers' overrun) slides to the front.
#if defined(__x86_64__)
__attribute__((target("avx2"))) static void stream_out(uint8_t *dst, const uint8_t *src, size_t bytes)      // all 64-byte multiples / aligned
{
	for (size_t i = 0; i < bytes; i += 64) {
		_mm256_stream_si256((__m256i *)(dst + i), _mm256_load_si256((const __m256i *)(src + i)));
		_mm256_stream_si256((__m256i *)(dst + i + 32), _mm256_load_si256((const __m256i *)(src + i + 32)));
	}
}
static const bool g_have_avx2 = __builtin_cpu_supports("avx2");
#endif

void Packer::flush_blocks(bool all)
{
	const uint64_t upto = all ? (pos + kBlockPos - 1) / kBlockPos * kBlockPos : pos / kBlockPos * kBlockPos;   // `all`: the ragged tail too (finish() pads it first)
	const uint64_t n = upto - origin;
	if (n) {
		uint8_t *db = reinterpret_cast<uint8_t *>(bases) + origin / 4, *dm = reinterpret_cast<uint8_t *>(mask) + origin / 8;
#if defined(__x86_64__)
		if (g_have_avx2) {
			stream_out(db, stage_b, n / 4);
			stream_out(dm, stage_m, n / 8);
		} else
#endif
		{
			memcpy(db, stage_b, n / 4);
			memcpy(dm, stage_m, n / 8);
		}
	}
	// keep the unfinished block and the 64 positions the packers may already have written past pos
	const uint64_t keep = pos + 64 > upto ? pos + 64 - upto : 0;
	if (keep && n) {
		memmove(stage_b, stage_b + n / 4, (keep + 3) / 4);
		memmove(stage_m, stage_m + n / 8, (keep + 7) / 8);
	}
	origin = upto;
	wb = stage_b - origin / 4;
	wm = stage_m - origin / 8;
}

void Packer::put_read(const char *s, uint64_t n)
{
	if (streaming) {
		const uint64_t span = read_span(n);
		if (pos - origin + span + kSlackPos > kStagePos) {
			flush_blocks(false);
			if (span + kBlockPos + kSlackPos > kStagePos) {
				// a read longer than the staging area (long-read data): the finished blocks are out, the open
				// one goes out as it is, and this read is written in place with ordinary stores
				const uint64_t tail = pos - origin;
				memcpy(reinterpret_cast<uint8_t *>(bases) + origin / 4, stage_b, (tail + 3) / 4);
				memcpy(reinterpret_cast<uint8_t *>(mask) + origin / 8, stage_m, (tail + 7) / 8);
				wb = reinterpret_cast<uint8_t *>(bases);
				wm = reinterpret_cast<uint8_t *>(mask);
				g_run(*this, s, n);
				// staging resumes at the block that holds the new pos: bring that block's finished part back in
				origin = pos / kBlockPos * kBlockPos;
				const uint64_t back = pos - origin;
				memcpy(stage_b, reinterpret_cast<uint8_t *>(bases) + origin / 4, (back + 3) / 4);
				memcpy(stage_m, reinterpret_cast<uint8_t *>(mask) + origin / 8, (back + 7) / 8);
				wb = stage_b - origin / 4;
				wm = stage_m - origin / 8;
				return;
			}
		}
	}
	g_run(*this, s, n);
}

uint64_t Packer::finish()
{
	const uint64_t n = pos;                       // a multiple of 8: both planes end on a byte
	const uint64_t end = padded_positions(n);
	if (streaming) {
		// pad the open block inside the staging area, send everything out, then pad the rest in place
		const uint64_t blk_end = (n + kBlockPos - 1) / kBlockPos * kBlockPos;
		memset(wb + n / 4, 0, (blk_end - n) / 4);
		memset(wm + n / 8, 0xFF, (blk_end - n) / 8);
		flush_blocks(true);
		memset(reinterpret_cast<uint8_t *>(bases) + blk_end / 4, 0, (end - blk_end) / 4);
		memset(reinterpret_cast<uint8_t *>(mask) + blk_end / 8, 0xFF, (end - blk_end) / 8);
#if defined(__x86_64__)
		_mm_sfence();                             // the streamed lines are globally visible before the batch is handed to the DMA engine
#endif
		streaming = false;
		wb = reinterpret_cast<uint8_t *>(bases);
		wm = reinterpret_cast<uint8_t *>(mask);
		pos = end;
		return n;
	}
	memset(reinterpret_cast<uint8_t *>(bases) + n / 4, 0, (end - n) / 4);
	memset(reinterpret_cast<uint8_t *>(mask) + n / 8, 0xFF, (end - n) / 8);
	pos = end;
	return n;
}

}  // namespace ntsm

extern "C" uint64_t ntsm_padded_positions(uint64_t n_pos) { return ntsm::padded_positions(n_pos); }

extern "C" uint32_t ntsm_nt4(uint8_t c) { return ntsm::nt4(c); }

extern "C" const char *ntsm_pack_isa(const char *force)
{
	if (force) {
		if (*force) setenv("NTSM_PACK_ISA", force, 1);
		else unsetenv("NTSM_PACK_ISA");
		ntsm::pack_reselect();
	}
	return ntsm::pack_isa();
}

extern "C" uint64_t ntsm_pack_reads(const char *buf, const uint64_t *off, uint64_t n_reads, uint32_t *bases2,
                                    uint32_t *nmask, uint64_t *read_off)
{
	return ntsm_pack_reads2(buf, off, n_reads, bases2, nmask, read_off, 0);
}

extern "C" uint64_t ntsm_pack_reads2(const char *buf, const uint64_t *off, uint64_t n_reads, uint32_t *bases2,
                                     uint32_t *nmask, uint64_t *read_off, int streaming)
{
	ntsm::Packer p;
	if (streaming && ((uintptr_t)bases2 % 64 == 0) && ((uintptr_t)nmask % 64 == 0)) p.reset_streaming(reinterpret_cast<uint64_t *>(bases2), nmask);
	else p.reset(reinterpret_cast<uint64_t *>(bases2), nmask);
	for (uint64_t r = 0; r < n_reads; ++r) {
		if (read_off) read_off[r] = p.pos;
		p.put_read(buf + off[r], off[r + 1] - off[r]);
	}
	if (read_off) read_off[n_reads] = p.pos;
	return p.finish();
}
