// pack.cpp -- host packer (see pack.h).
#include "pack.h"

#include "../../include/ntsm_b200.h"
#include "kmer_math.h"

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

#if defined(__x86_64__)
// 32 ASCII bytes -> (64 bits of 2-bit codes, 32 invalid flags).
// Valid letters: A C G T U in either case (bit 5 cleared folds the case) and raw bytes 0..3.
__attribute__((target("avx2,bmi2"))) static inline void pack32_avx2(const char *s, uint64_t *b, uint32_t *m)
{
	const __m256i v = _mm256_loadu_si256((const __m256i *)s);
	const __m256i up = _mm256_and_si256(v, _mm256_set1_epi8((char)0xDF));        // fold case
	const __m256i isA = _mm256_cmpeq_epi8(up, _mm256_set1_epi8('A'));
	const __m256i isC = _mm256_cmpeq_epi8(up, _mm256_set1_epi8('C'));
	const __m256i isG = _mm256_cmpeq_epi8(up, _mm256_set1_epi8('G'));
	const __m256i isT = _mm256_or_si256(_mm256_cmpeq_epi8(up, _mm256_set1_epi8('T')),
	                                    _mm256_cmpeq_epi8(up, _mm256_set1_epi8('U')));
	// raw bytes 0..3 decode to themselves (table rows 0-3)
	const __m256i isRaw = _mm256_cmpeq_epi8(_mm256_and_si256(v, _mm256_set1_epi8((char)0xFC)), _mm256_setzero_si256());
	const __m256i r1 = _mm256_and_si256(isRaw, _mm256_cmpeq_epi8(_mm256_and_si256(v, _mm256_set1_epi8(1)), _mm256_set1_epi8(1)));
	const __m256i r2 = _mm256_and_si256(isRaw, _mm256_cmpeq_epi8(_mm256_and_si256(v, _mm256_set1_epi8(2)), _mm256_set1_epi8(2)));
	// code bit0 set for C,T ; bit1 set for G,T
	const uint32_t bit0 = (uint32_t)_mm256_movemask_epi8(_mm256_or_si256(_mm256_or_si256(isC, isT), r1));
	const uint32_t bit1 = (uint32_t)_mm256_movemask_epi8(_mm256_or_si256(_mm256_or_si256(isG, isT), r2));
	const uint32_t valid = (uint32_t)_mm256_movemask_epi8(
	    _mm256_or_si256(_mm256_or_si256(_mm256_or_si256(isA, isC), _mm256_or_si256(isG, isT)), isRaw));
	*b = _pdep_u64(bit0, 0x5555555555555555ULL) | _pdep_u64(bit1, 0xAAAAAAAAAAAAAAAAULL);
	*m = ~valid;
}
static const bool g_have_avx2 = __builtin_cpu_supports("avx2") && __builtin_cpu_supports("bmi2");
#else
static const bool g_have_avx2 = false;
#endif

static inline void pack32_scalar(const char *s, uint64_t *b, uint32_t *m)
{
	const uint8_t *t = g_code_init.t;
	uint64_t bb = 0;
	uint32_t mm = 0;
	for (int j = 0; j < 32; ++j) {
		const unsigned c = t[(unsigned char)s[j]];
		bb |= (uint64_t)(c & 3) << (2 * j);
		mm |= (uint32_t)(c >> 2) << j;
	}
	*b = bb;
	*m = mm;
}

void Packer::put_bases(const char *s, uint64_t n)
{
	const uint8_t *t = g_code_init.t;
	uint64_t i = 0;
	if (n >= 64) {
		const unsigned sh = (unsigned)pos & 31;   // positions already in the partial word
		// whole 32-byte groups, merged into the stream at bit offset sh
		for (; i + 32 <= n; i += 32) {
			uint64_t b;
			uint32_t m;
#if defined(__x86_64__)
			if (g_have_avx2) pack32_avx2(s + i, &b, &m);
			else
#endif
				pack32_scalar(s + i, &b, &m);
			if (sh == 0) {
				bases[pos >> 5] = b;
				mask[pos >> 5] = m;
			} else {
				bases[pos >> 5] = bacc | (b << (2 * sh));
				mask[pos >> 5] = macc | (m << sh);
				bacc = b >> (64 - 2 * sh);
				macc = m >> (32 - sh);
			}
			pos += 32;
		}
	}
	for (; i < n; ++i) put_code(t[(unsigned char)s[i]]);
}

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

extern "C" uint64_t ntsm_pack_reads(const char *buf, const uint64_t *off, uint64_t n_reads, uint32_t *bases2,
                                    uint32_t *nmask, uint64_t *read_off)
{
	ntsm::Packer p;
	p.reset(reinterpret_cast<uint64_t *>(bases2), nmask);
	for (uint64_t r = 0; r < n_reads; ++r) {
		if (read_off) read_off[r] = p.pos;
		p.put_bases(buf + off[r], off[r + 1] - off[r]);
		p.put_separator();
	}
	if (read_off) read_off[n_reads] = p.pos;
	return p.finish();
}
