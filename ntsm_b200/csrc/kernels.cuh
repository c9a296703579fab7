// kernels.cuh -- device side of the counting path (sm_100a).
//
// What the reference does per base (FingerPrint::insertCount, src/FingerPrint.hpp:89-103, over
// KseqHashIterator::step, vendor/KseqHashIterator.hpp:95-112): decode a byte, roll fw and rv,
// take min, run the 7-stage hash64, look the hash up in a 100 MB robin_map, and bump three
// shared counters with locked RMWs.
//
// What this kernel does per position of the packed stream:
//   1. the k-mer starting at position p is just 2k contiguous bits of the 2-bit stream, so it is
//      cut out with two funnel shifts -- nothing is rolled;
//   2. a multiplicative mix of those bits indexes a bitmap pre-filter that holds BOTH orientations
//      of every site k-mer (so no reverse complement and no min() is needed to decide "cannot be a
//      site k-mer"); >98% of positions end here;
//   3. survivors are turned into the reference's canonical value (fw = 2-bit groups reversed,
//      rv = ~s & mask, see kmer_math.h), hashed with the reference hash64, and probed in an
//      open-addressing table keyed by that hash; a hit is one atomicAdd into counts[dense index];
//   4. window validity (no N / separator among the k positions) is computed for 32 positions at a
//      time with shifts of the N-mask words; #@TK is the popcount of valid windows.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "kmer_math.h"

namespace ntsm {

struct __align__(16) TableSlot {
	uint64_t key;    // reference hash64 value; kEmptyKey = unused
	uint32_t idx;    // dense k-mer index into counts[]
	uint32_t pad;
};
constexpr uint64_t kEmptyKey = ~0ULL;

struct CountParams {
	const uint2 *bases;        // 32 positions per element (two little-endian uint32 words)
	const uint32_t *nmask;     // 32 positions per element
	uint64_t n_chunks;         // number of 32-position chunks to scan (arrays hold n_chunks + 2 at least)
	const uint32_t *minimizer; // level-1 bitmap over m-mers, 4^M bits (k = 19 kernel only)
	const uint32_t *filter;    // bitmap, 2^filter_bits bits
	uint32_t filter_shift;     // 32 - filter_bits
	const TableSlot *table;
	uint32_t table_mask;       // capacity - 1
	uint32_t k;
	uint32_t *counts;
	unsigned long long *totals;   // [0] valid windows (TK), [1] hits
};

constexpr int kCountThreads = 256;

// stream-order k-mer starting `i` positions into the 128-bit window w[0..3]
template <int K>
__device__ __forceinline__ void cut_kmer(const uint32_t (&w)[4], int i, uint32_t k_rt, uint32_t &lo, uint32_t &hi)
{
	const int a = i >> 4;
	const int sh = (2 * i) & 31;
	const uint32_t w0 = a ? w[1] : w[0], w1 = a ? w[2] : w[1], w2 = a ? w[3] : w[2];
	lo = __funnelshift_r(w0, w1, sh);
	hi = __funnelshift_r(w1, w2, sh);
	const uint64_t m = kmer_mask(K ? (unsigned)K : k_rt);
	lo &= (uint32_t)m;
	hi &= (uint32_t)(m >> 32);
}

// exact path for the few positions that survive the pre-filters: canonical value, reference hash64,
// open-addressing probe, atomicAdd.  Returns the number of table hits.
template <int K>
__device__ __forceinline__ uint32_t resolve_survivors(const CountParams &P, const uint32_t (&w)[4], uint32_t pass,
                                                      uint32_t k, uint64_t kmask)
{
	uint32_t hits = 0;
	while (pass) {
		const int i = __ffs(pass) - 1;
		pass &= pass - 1;
		uint32_t lo, hi;
		cut_kmer<K>(w, i, k, lo, hi);
		const uint64_t s = ((uint64_t)hi << 32) | lo;
		const uint64_t fw = stream_to_fw(s, k), rv = stream_to_rv(s, kmask);
		const uint64_t h = hash64(fw < rv ? fw : rv, kmask);     // KseqHashIterator.hpp:102
		uint32_t slot = (uint32_t)(h ^ (h >> 29)) & P.table_mask;
		for (;;) {
			const TableSlot e = P.table[slot];
			if (e.key == h) {
				atomicAdd(P.counts + e.idx, 1u);                 // FingerPrint.hpp:93-94
				++hits;
				break;
			}
			if (e.key == kEmptyKey) break;
			slot = (slot + 1) & P.table_mask;
		}
	}
	return hits;
}

// 32 valid-window flags for the chunk whose N-mask words are m0 (own) and m1 (next)
__device__ __forceinline__ uint32_t valid_windows(uint32_t m0, uint32_t m1, uint32_t k)
{
	uint64_t bad = (uint64_t)m0 | ((uint64_t)m1 << 32);     // bit i = some position in [i, i+k) is invalid
	uint32_t r = 1;
	while (2 * r <= k) { bad |= bad >> r; r *= 2; }
	bad |= bad >> (k - r);
	return ~(uint32_t)bad;
}

// per-CTA tallies -> two atomics per CTA (FingerPrint.hpp:95-99 did one locked RMW per k-mer)
__device__ __forceinline__ void flush_tallies(uint32_t tk, uint32_t hits, unsigned long long *totals)
{
	__shared__ unsigned long long s_tk, s_hits;
	if (threadIdx.x == 0) { s_tk = 0; s_hits = 0; }
	__syncthreads();
#pragma unroll
	for (int o = 16; o; o >>= 1) {
		tk += __shfl_xor_sync(0xffffffffu, tk, o);
		hits += __shfl_xor_sync(0xffffffffu, hits, o);
	}
	if ((threadIdx.x & 31) == 0) {
		atomicAdd(&s_tk, (unsigned long long)tk);
		if (hits) atomicAdd(&s_hits, (unsigned long long)hits);
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		if (s_tk) atomicAdd(totals + 0, s_tk);
		if (s_hits) atomicAdd(totals + 1, s_hits);
	}
}

template <int K>
__global__ void __launch_bounds__(kCountThreads) count_kernel(const CountParams P)
{
	const uint32_t k = K ? (uint32_t)K : P.k;
	const uint64_t kmask = kmer_mask(k);
	uint32_t tk = 0, hits = 0;

	const uint64_t stride = (uint64_t)gridDim.x * kCountThreads;
	for (uint64_t c = (uint64_t)blockIdx.x * kCountThreads + threadIdx.x; c < P.n_chunks; c += stride) {
		// 32 own positions + the next 32 (a window may reach k-1 <= 30 positions past its start)
		const uint2 own = __ldg(P.bases + c), nxt = __ldg(P.bases + c + 1);
		const uint32_t m0 = __ldg(P.nmask + c), m1 = __ldg(P.nmask + c + 1);
		const uint32_t w[4] = { own.x, own.y, nxt.x, nxt.y };

		const uint32_t valid = valid_windows(m0, m1, k);
		if (valid == 0) continue;
		tk += __popc(valid);

		// pre-filter: one bitmap probe per position
		uint32_t pass = 0;
#pragma unroll
		for (int half = 0; half < 2; ++half) {
			uint32_t word[16], bit[16];
#pragma unroll
			for (int j = 0; j < 16; ++j) {
				uint32_t lo, hi;
				cut_kmer<K>(w, half * 16 + j, k, lo, hi);
				const uint32_t ix = filter_mix(lo, hi) >> P.filter_shift;
				bit[j] = ix & 31;
				word[j] = __ldg(P.filter + (ix >> 5));
			}
#pragma unroll
			for (int j = 0; j < 16; ++j) pass |= ((word[j] >> bit[j]) & 1u) << (half * 16 + j);
		}
		pass &= valid;

		hits += resolve_survivors<K>(P, w, pass, k, kmask);
	}

	flush_tallies(tk, hits, P.totals);
}

// ---------------------------------------------------------------------------------------------
// Minimizer-gated kernel (the production path for the reference's default k = 19).
//
// The plain kernel above is bound by one L1->L2 request per position (ncu: l1tex2xbar 97 %, L2 tag
// 80 %, issue slots 21 %).  Adjacent k-mers overlap in k-1 bases, so they share their minimizer
// (the M-mer with the smallest multiplicative hash among the K-M+1 M-mers of the k-mer) for ~4
// positions in a row.  A direct-addressed bitmap over all 4^M M-mers marks the minimizers of every
// site k-mer (both read orientations).  A thread walks its 32 positions and probes that bitmap only
// where the minimizer CHANGES; a position whose minimizer is not a site minimizer cannot be a site
// k-mer.  That cuts the memory requests per position ~3x and leaves the k-mer bitmap and the exact
// table for the ~5 % of positions that pass.  Everything still needed for the reference's result
// (validity, canonical value, hash64, table) is unchanged.
constexpr uint32_t kMinHashMul = 0x9E3779B1u;      // odd: m -> m * C mod 2^32 is a bijection ...
constexpr uint32_t kMinHashInv = 0x0E8B2F51u;      // ... and this is its inverse (C * Cinv == 1 mod 2^32)

constexpr int kMinimizerM = 13;                    // M-mer length for k = 19 (4^13 bits = 8 MiB bitmap)
static_assert((uint32_t)(kMinHashMul * kMinHashInv) == 1u, "kMinHashInv must invert kMinHashMul mod 2^32");

// minimizer M-mer (stream order) of a stream-order k-mer, as the kernel selects it
NTSM_HD uint32_t minimizer_of(uint64_t s, int k, int m)
{
	const uint32_t mm_mask = (uint32_t)((1ull << (2 * m)) - 1);
	uint32_t best = 0xFFFFFFFFu;
	for (int j = 0; j + m <= k; ++j) {
		const uint32_t hj = ((uint32_t)(s >> (2 * j)) & mm_mask) * kMinHashMul;
		best = hj < best ? hj : best;
	}
	return best * kMinHashInv;
}

template <int K, int M>
__global__ void __launch_bounds__(kCountThreads) count_kernel_min(const CountParams P)
{
	constexpr int W = K - M + 1;                  // M-mers per k-mer
	constexpr int NH = 32 + W - 1;                // M-mer hashes a 32-position chunk needs
	constexpr uint32_t MM = (uint32_t)((1ull << (2 * M)) - 1);
	static_assert(2 * M <= 30 && W >= 2 && 2 * (NH - 1) + 2 * M <= 128, "window does not fit the 128-bit register view");
	const uint64_t kmask = kmer_mask(K);
	uint32_t tk = 0, hits = 0;

	const uint64_t stride = (uint64_t)gridDim.x * kCountThreads;
	for (uint64_t c = (uint64_t)blockIdx.x * kCountThreads + threadIdx.x; c < P.n_chunks; c += stride) {
		const uint2 own = __ldcs(P.bases + c), nxt = __ldcs(P.bases + c + 1);      // streamed once: evict-first
		const uint32_t m0 = __ldcs(P.nmask + c), m1 = __ldcs(P.nmask + c + 1);
		const uint32_t w[4] = { own.x, own.y, nxt.x, nxt.y };
		const uint32_t valid = valid_windows(m0, m1, K);
		if (valid == 0) continue;
		tk += __popc(valid);

		// hash of the M-mer starting at each of the NH positions
		uint32_t h[NH];
#pragma unroll
		for (int j = 0; j < NH; ++j) {
			const int a = j >> 4, sh = (2 * j) & 31;
			h[j] = (__funnelshift_r(w[a], w[a + 1 < 4 ? a + 1 : 3], sh) & MM) * kMinHashMul;
		}
		// sliding minimum over W consecutive hashes (van Herk / Gil-Werman: ~3 min per position)
		uint32_t win[32];
#pragma unroll
		for (int i = 0; i < 32; ++i) {
			const int b = i / W * W;               // block of W that position i falls in
			uint32_t sfx = h[b + W - 1];           // suffix minimum of that block from i
#pragma unroll
			for (int t = b + W - 2; t >= i; --t) sfx = min(sfx, h[t]);
			uint32_t v = sfx;
			if (i != b) {                          // prefix minimum of the next block up to i+W-1
				uint32_t pfx = h[b + W];
#pragma unroll
				for (int t = b + W + 1; t <= i + W - 1; ++t) pfx = min(pfx, h[t]);
				v = min(sfx, pfx);
			}
			win[i] = v;
		}
		// probe the minimizer bitmap where the minimizer changes, carry the answer along otherwise
		uint32_t pass = 0;
#pragma unroll
		for (int half = 0; half < 2; ++half) {
			uint32_t word[16];
#pragma unroll
			for (int j = 0; j < 16; ++j) {
				const int i = half * 16 + j;
				const bool changed = (i == 0) || (win[i] != win[i - 1]);
				const uint32_t mm = win[i] * kMinHashInv;          // the M-mer itself (hash is invertible)
				word[j] = changed ? __ldg(P.minimizer + (mm >> 5)) >> (mm & 31) : 0u;
			}
			uint32_t bit = half ? (pass >> 15) & 1u : 0u;          // carried over from the first half
#pragma unroll
			for (int j = 0; j < 16; ++j) {
				const int i = half * 16 + j;
				const bool changed = (i == 0) || (win[i] != win[i - 1]);
				bit = changed ? (word[j] & 1u) : bit;
				pass |= bit << i;
			}
		}
		pass &= valid;
		// level 2: the k-mer bitmap, only for positions whose minimizer is a site minimizer
		uint32_t pass2 = 0;
		while (pass) {
			const int i = __ffs(pass) - 1;
			pass &= pass - 1;
			uint32_t lo, hi;
			cut_kmer<K>(w, i, K, lo, hi);
			const uint32_t ix = filter_mix(lo, hi) >> P.filter_shift;
			pass2 |= ((__ldg(P.filter + (ix >> 5)) >> (ix & 31)) & 1u) << i;
		}
		if (pass2) hits += resolve_survivors<K>(P, w, pass2, K, kmask);
	}
	flush_tallies(tk, hits, P.totals);
}

// printCountsMax's per-site loop (src/FingerPrint.hpp:281-294): max and sum (mod 2^32) over the
// ref list and over the var list of every site.  One thread per site; lists are 0..13 long.
__global__ void site_reduce_kernel(const uint32_t *__restrict__ counts, const uint32_t *__restrict__ allele_off,
                                   uint32_t n_sites, uint32_t *__restrict__ max_ref, uint32_t *__restrict__ max_var,
                                   uint32_t *__restrict__ sum_ref, uint32_t *__restrict__ sum_var)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_sites) return;
	const uint32_t r0 = allele_off[2 * i], r1 = allele_off[2 * i + 1], v1 = allele_off[2 * i + 2];
	uint32_t mr = 0, sr = 0, mv = 0, sv = 0;
	for (uint32_t j = r0; j < r1; ++j) { const uint32_t c = counts[j]; mr = max(mr, c); sr += c; }
	for (uint32_t j = r1; j < v1; ++j) { const uint32_t c = counts[j]; mv = max(mv, c); sv += c; }
	max_ref[i] = mr; max_var[i] = mv; sum_ref[i] = sr; sum_var[i] = sv;
}

}  // namespace ntsm
