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

static inline void pack32_scalar(const char *s, unsigned n, uint64_t *b, uint32_t *m)
{
	const uint8_t *t = g_code_init.t;
	uint64_t bb = 0;
	uint32_t mm = 0;
	for (unsigned j = 0; j < n; ++j) {
		const unsigned c = t[(unsigned char)s[j]];
		bb |= (uint64_t)(c & 3) << (2 * j);
		mm |= (uint32_t)(c >> 2) << j;
	}
	*b = bb;
	*m = mm;
}

// One run of n bytes (+ the separator position when sep) appended to the stream.  Three
// implementations of the same function, picked once at start-up: AVX-512 VBMI (64 bases per
// step: one byte permute through the 128-entry decode table), AVX2+BMI2 (32 per step: compares),
// scalar table walk.  The interleave of the two code bit-planes into 2-bit fields is pdep.
static void run_scalar(Packer &dst, const char *s, uint64_t n, bool sep)
{
	Packer p = dst;          // local copy: stores into the stream cannot alias the accumulators
	uint64_t i = 0, b;
	uint32_t m;
	for (; i + 32 <= n; i += 32) {
		pack32_scalar(s + i, 32, &b, &m);
		p.put_group(b, m, 32);
	}
	const unsigned r = (unsigned)(n - i);
	pack32_scalar(s + i, r, &b, &m);
	if (sep) m |= 1u << r;
	if (r + sep) p.put_group(b, m, r + sep);
	dst = p;
}

#if defined(__x86_64__)
#define NTSM_TGT_AVX2 __attribute__((target("avx2,bmi2")))
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

NTSM_TGT_AVX2 static void run_avx2(Packer &dst, const char *s, uint64_t n, bool sep)
{
	Packer p = dst;          // local copy: stores into the stream cannot alias the accumulators
	const uint64_t kEven = 0x5555555555555555ULL, kOdd = 0xAAAAAAAAAAAAAAAAULL;
	uint64_t i = 0;
	uint32_t b0, b1, va;
	for (; i + 32 <= n; i += 32) {
		planes32_avx2(_mm256_loadu_si256((const __m256i *)(s + i)), &b0, &b1, &va);
		p.put_group(_pdep_u64(b0, kEven) | _pdep_u64(b1, kOdd), ~va, 32);
	}
	const unsigned r = (unsigned)(n - i);           // 0..31 bytes left; the separator rides in the same group
	if (r + sep == 0) { dst = p; return; }
	// A 32-byte load that stays inside the 4 KiB page of its first byte cannot fault, so the bytes
	// after the run are read and masked off; only a load that would cross a page is bounced.
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
	const uint32_t keep = (1u << r) - 1;
	p.put_group(_pdep_u64(b0 & keep, kEven) | _pdep_u64(b1 & keep, kOdd), (~va & keep) | ((uint32_t)sep << r), r + sep);
	dst = p;
}

struct alignas(64) Vbmi128 {
	uint8_t t[128];
	Vbmi128() { for (int i = 0; i < 128; ++i) { const unsigned c = nt4((unsigned char)i); t[i] = c < 4 ? (uint8_t)c : 0x80; } }
};
static const Vbmi128 g_vbmi_tab;

NTSM_TGT_AVX512 static inline void planes64_avx512(__m512i v, __m512i tab_lo, __m512i tab_hi, uint64_t *bit0, uint64_t *bit1,
                                                  uint64_t *inv)
{
	// index bits 0-6 pick one of 128 table bytes (bit 7 is ignored by the permute); bytes >= 0x80
	// are never bases, their own sign bit marks them invalid
	const __m512i code = _mm512_permutex2var_epi8(tab_lo, v, tab_hi);
	const uint64_t bad = (uint64_t)_mm512_movepi8_mask(_mm512_or_si512(code, v));
	*bit0 = _mm512_test_epi8_mask(code, _mm512_set1_epi8(1)) & ~bad;     // 0x81 would otherwise decode like 0x01
	*bit1 = _mm512_test_epi8_mask(code, _mm512_set1_epi8(2)) & ~bad;
	*inv = bad;
}

NTSM_TGT_AVX512 static void run_avx512(Packer &dst, const char *s, uint64_t n, bool sep)
{
	Packer p = dst;          // local copy: stores into the stream cannot alias the accumulators
	const uint64_t kEven = 0x5555555555555555ULL, kOdd = 0xAAAAAAAAAAAAAAAAULL;
	const __m512i tab_lo = _mm512_load_si512(g_vbmi_tab.t), tab_hi = _mm512_load_si512(g_vbmi_tab.t + 64);
	uint64_t i = 0, b0, b1, iv;
	for (; i + 64 <= n; i += 64) {
		planes64_avx512(_mm512_loadu_si512(s + i), tab_lo, tab_hi, &b0, &b1, &iv);
		p.put_group(_pdep_u64(b0, kEven) | _pdep_u64(b1, kOdd), (uint32_t)iv, 32);
		p.put_group(_pdep_u64(b0 >> 32, kEven) | _pdep_u64(b1 >> 32, kOdd), (uint32_t)(iv >> 32), 32);
	}
	const unsigned r = (unsigned)(n - i);           // 0..63 bytes left
	const unsigned total = r + sep;
	if (total == 0) { dst = p; return; }
	const uint64_t keep = (1ull << r) - 1;          // a masked load never touches (or faults on) the bytes it skips
	planes64_avx512(_mm512_maskz_loadu_epi8((__mmask64)keep, s + i), tab_lo, tab_hi, &b0, &b1, &iv);
	b0 &= keep;
	b1 &= keep;
	iv = (iv & keep) | ((uint64_t)sep << r);
	if (total <= 32) {
		p.put_group(_pdep_u64(b0, kEven) | _pdep_u64(b1, kOdd), (uint32_t)iv, total);
	} else {
		p.put_group(_pdep_u64(b0, kEven) | _pdep_u64(b1, kOdd), (uint32_t)iv, 32);
		p.put_group(_pdep_u64(b0 >> 32, kEven) | _pdep_u64(b1 >> 32, kOdd), (uint32_t)(iv >> 32), total - 32);
	}
	dst = p;
}
#endif

typedef void (*RunFn)(Packer &, const char *, uint64_t, bool);
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

void Packer::put_bases(const char *s, uint64_t n) { g_run(*this, s, n, false); }
void Packer::put_read(const char *s, uint64_t n) { g_run(*this, s, n, true); }

uint64_t Packer::finish()
{
	const uint64_t n = pos;
	const uint64_t end = padded_positions(n);
	while (pos & 31) put_code(4);
	for (uint64_t w = pos >> 5; w < end >> 5; ++w) {
		bases[w] = 0;
		mask[w] = 0xFFFFFFFFu;
	}
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
	ntsm::Packer p;
	p.reset(reinterpret_cast<uint64_t *>(bases2), nmask);
	for (uint64_t r = 0; r < n_reads; ++r) {
		if (read_off) read_off[r] = p.pos;
		p.put_read(buf + off[r], off[r + 1] - off[r]);
	}
	if (read_off) read_off[n_reads] = p.pos;
	return p.finish();
}
