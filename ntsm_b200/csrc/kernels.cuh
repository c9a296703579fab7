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
	const uint32_t *minimizer; // level-1 bitmap over M-mers, 4^M bits (k = 19 kernels only)
	const uint32_t *level0;    // kL0Words-word image of the shared-memory level-0 bitmap (gated kernel)
	const uint32_t *minimizer2;// level-1 bitmap of the gated kernel (hashed M-mer order, 4^kGateM bits)
	const uint32_t *filter;    // bitmap, 2^filter_bits bits
	uint32_t filter_shift;     // 32 - filter_bits
	const TableSlot *table;
	uint32_t table_mask;       // capacity - 1
	uint32_t k;
	uint32_t four;             // = 4, kept in a register so address scaling stays an IMAD (FMA pipe), not an LEA
	uint32_t pair_word_mask;   // paired-seed kernel: (words of the pair table) - 1
	uint32_t delta;            // added to counts[idx] per hit: 1 to count, 0xFFFFFFFF to take a batch back out, 0 to only tally
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
				atomicAdd(P.counts + e.idx, P.delta);            // FingerPrint.hpp:93-94 (delta = 1)
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
// Predicated loads that stay predicated: nvcc turns `p ? __ldg(a) : 0` into a divergent branch per
// position (ncu: 15 of 32 lanes active on average); one predicated LDG/LDS costs one issue slot.
__device__ __forceinline__ uint32_t ldg_u32_if(const uint32_t *p, bool pred)
{
	uint32_t v;
	asm("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %2, 0;\n\tmov.u32 %0, 0;\n\t@q ld.global.nc.u32 %0, [%1];\n\t}"
	    : "=r"(v) : "l"(p), "r"((int)pred));
	return v;
}
// mul.hi kept as a multiply (FMA pipe) even when the factor is a power of two
__device__ __forceinline__ uint32_t mulhi_pipe(uint32_t a, uint32_t b)
{
	uint32_t d;
	asm("mul.hi.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
	return d;
}
__device__ __forceinline__ uint32_t lds_u32_if(uint32_t shared_addr, bool pred)
{
	uint32_t v;
	asm("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %2, 0;\n\tmov.u32 %0, 0;\n\t@q ld.shared.u32 %0, [%1];\n\t}"
	    : "=r"(v) : "r"(shared_addr), "r"((int)pred));
	return v;
}

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
				word[j] = ldg_u32_if(P.minimizer + (mm >> 5), changed) >> (mm & 31);
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


// ---------------------------------------------------------------------------------------------
// Gated kernel: the same minimizer idea with a shared-memory level 0 in front of it.
//
// ncu on count_kernel_min: l1tex2xbar 85 %, L2 tag 70 %, ALU pipe 71 % -- still ~one L1->L2
// request per 3 positions, and that request rate (1 per clock per SM) is the wall.  A persistent
// CTA per SM now keeps a 224 KiB bitmap of the site minimizers in shared memory (hashed, ~45 %
// full): a changed minimizer is first tested there (a shared-memory probe costs ~1/10 of a global
// one) and only survivors go to the exact 4^M-bit bitmap in L2.  Other changes over
// count_kernel_min: M-mer hash = window * (C << (32-2M)) (no masking: the shift kills the bits
// above the M-mer) and the bitmap is stored in that hashed order (no inverse multiply); the
// per-position carry of the probe result is a 4-instruction bit trick over the whole chunk;
// level 2 tests two bits per word (false positives 0.1 % instead of 2 %); the halo words come
// from the neighbouring lane by shuffle instead of a second load.
constexpr int kGateM = 14;
constexpr int kGateThreads = 1024;
constexpr uint32_t kL0Words = 57344;                       // 224 KiB
constexpr uint32_t kL0Bits = kL0Words * 32;
constexpr uint32_t kL0Mul = 0x85EBCA6Bu;
constexpr uint32_t kGateHashMul = kMinHashMul << (32 - 2 * kGateM);

// hashed id (0 .. 4^M-1) of the minimizer of stream-order k-mer s, as the gated kernel selects it
NTSM_HD uint32_t gate_minimizer_id(uint64_t s, int k)
{
	uint32_t best = 0xFFFFFFFFu;
	for (int j = 0; j + kGateM <= k; ++j) {
		const uint32_t hj = (uint32_t)(s >> (2 * j)) * kGateHashMul;
		best = hj < best ? hj : best;
	}
	return best >> (32 - 2 * kGateM);
}
// where minimizer id lives: level 0 (shared memory) word/bit and level 1 (global) word/bit.  Both
// use the low 5 bits of t = id * kL0Mul as the bit index, so the kernel extracts with one wrapping
// shift by t and never masks (t & 31 is a permutation of id & 31, so level 1 stays exact).
NTSM_HD uint32_t mulhi_u32(uint32_t a, uint32_t b)
{
#if defined(__CUDA_ARCH__)
	return __umulhi(a, b);
#else
	return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}
NTSM_HD void gate_slots(uint32_t id, uint32_t &l0_word, uint32_t &l1_word, uint32_t &bit)
{
	const uint32_t t = id * kL0Mul;
	l0_word = mulhi_u32(t, kL0Words);
	l1_word = id >> 5;
	bit = 31 - (t & 31);        // the kernel shifts LEFT by t so the wanted bit lands in the sign position
}
// level 2: word from the top bits of the k-mer mix, two bit positions from a second multiply
NTSM_HD void filter2_slots(uint32_t mix, uint32_t filter_shift, uint32_t &word, uint32_t &mask)
{
	word = mix >> (filter_shift + 5);
	const uint32_t t = mix * 0xC2B2AE35u;
	mask = (1u << (t >> 27)) | (1u << ((t >> 22) & 31));
}

// One position of the gated probe, written as PTX so that (a) both loads stay predicated instead
// of becoming divergent branches, (b) constant right shifts are mul.hi (FMA pipe) and (c) no
// predicate ever round-trips through a register.  `bit` carries the last probe result, `pass`
// collects one bit per position (the sign bit of `bit` is shifted in at the bottom, so after 32
// steps position i sits at bit 31-i).  Slot layout: gate_slots().
#define NTSM_STR2(x) #x
#define NTSM_STR(x) NTSM_STR2(x)
template <int SH>
__device__ __forceinline__ void gate_probe_step(uint32_t cur, uint32_t prev, uint32_t s_l0_addr, const uint32_t *level1,
                                                uint32_t four, uint32_t &bit, uint32_t &pass)
{
	asm("{\n\t"
	    ".reg .pred pc, pm;\n\t"
	    ".reg .u32 id, t, a0, w0, x0, i1, w1, nb;\n\t"
	    ".reg .u64 a1;\n\t"
	    "setp.ne.u32 pc, %2, %3;\n\t"                 // minimizer changed?
	    "mul.hi.u32 id, %2, %6;\n\t"                  // id = cur >> SH
	    "mul.lo.u32 t, id, %7;\n\t"
	    "mul.hi.u32 a0, t, %8;\n\t"                   // level-0 word
	    "mad.lo.u32 a0, a0, %10, %4;\n\t"
	    "mov.u32 w0, 0;\n\t"
	    "@pc ld.shared.u32 w0, [a0];\n\t"
	    "shf.l.wrap.b32 x0, w0, w0, t;\n\t"           // wanted bit -> sign position
	    "setp.lt.s32 pm, x0, 0;\n\t"                  // level 0 says maybe (w0 == 0 unless changed)
	    "mul.hi.u32 i1, %2, %9;\n\t"                  // level-1 word = id >> 5
	    "mad.wide.u32 a1, i1, %10, %5;\n\t"
	    "mov.u32 w1, 0;\n\t"
	    "@pm ld.global.nc.u32 w1, [a1];\n\t"
	    "shf.l.wrap.b32 nb, w1, w1, t;\n\t"
	    "@pc mov.u32 %0, nb;\n\t"
	    "shf.l.wrap.b32 %1, %0, %1, 1;\n\t"           // pass = pass << 1 | bit >> 31
	    "}"
	    : "+r"(bit), "+r"(pass)
	    : "r"(cur), "r"(prev), "r"(s_l0_addr), "l"(level1), "r"(1u << (32 - SH)), "r"(kL0Mul), "r"(kL0Words),
	      "r"(1u << (32 - SH - 5)), "r"(four));
}

template <int K>
__global__ void __launch_bounds__(kGateThreads, 1) count_kernel_gate(const CountParams P)
{
	constexpr int M = kGateM;
	constexpr int W = K - M + 1;
	constexpr int NH = 32 + W - 1;
	constexpr int SH = 32 - 2 * M;
	static_assert(W >= 2 && 2 * (NH - 1) + 32 <= 128, "window does not fit the 128-bit register view");
	extern __shared__ uint32_t s_l0[];
	{
		const uint4 *src = reinterpret_cast<const uint4 *>(P.level0);
		uint4 *dst = reinterpret_cast<uint4 *>(s_l0);
		for (uint32_t i = threadIdx.x; i < kL0Words / 4; i += kGateThreads) dst[i] = __ldg(src + i);
	}
	__syncthreads();

	const uint32_t s_l0_addr = (uint32_t)__cvta_generic_to_shared(s_l0);
	// the level-1 base in an ordinary register pair: with a uniform-register base the 64-bit address
	// costs IMAD.WIDE + IADD3 + IADD3.X per probe, with a vector base it is one IMAD.WIDE
	const uint32_t *level1;
	asm volatile("mov.u64 %0, %1;" : "=l"(level1) : "l"(P.minimizer2));
	const uint64_t kmask = kmer_mask(K);
	const int lane = threadIdx.x & 31;
	uint32_t tk = 0, hits = 0;
	const uint64_t stride = (uint64_t)gridDim.x * kGateThreads;
	for (uint64_t base = (uint64_t)blockIdx.x * kGateThreads + (threadIdx.x & ~31u); base < P.n_chunks; base += stride) {
		const uint64_t c = base + lane;
		uint2 own = make_uint2(0, 0);
		uint32_t m0 = 0xFFFFFFFFu;
		if (c <= P.n_chunks) {                                   // chunk n_chunks is padding, always readable
			own = __ldcs(P.bases + c);
			m0 = __ldcs(P.nmask + c);
		}
		uint2 nxt;
		nxt.x = __shfl_down_sync(0xffffffffu, own.x, 1);
		nxt.y = __shfl_down_sync(0xffffffffu, own.y, 1);
		uint32_t m1 = __shfl_down_sync(0xffffffffu, m0, 1);
		if (lane == 31) {
			nxt = make_uint2(0, 0);
			m1 = 0xFFFFFFFFu;
			if (c + 1 <= P.n_chunks) {
				nxt = __ldcs(P.bases + c + 1);
				m1 = __ldcs(P.nmask + c + 1);
			}
		}
		if (c >= P.n_chunks) m0 = 0xFFFFFFFFu;                  // nothing starts in the padding chunk
		const uint32_t w[4] = { own.x, own.y, nxt.x, nxt.y };
		const uint32_t valid = valid_windows(m0, m1, K);
		if (valid == 0) continue;
		tk += __popc(valid);

		uint32_t h[NH];
#pragma unroll
		for (int j = 0; j < NH; ++j) {
			const int a = j >> 4, sh = (2 * j) & 31;
			h[j] = __funnelshift_r(w[a], w[a + 1 < 4 ? a + 1 : 3], sh) * kGateHashMul;
		}
		uint32_t win[32];
#pragma unroll
		for (int i = 0; i < 32; ++i) {
			const int b = i / W * W;
			uint32_t sfx = h[b + W - 1];
#pragma unroll
			for (int t = b + W - 2; t >= i; --t) sfx = min(sfx, h[t]);
			uint32_t v = sfx;
			if (i != b) {
				uint32_t pfx = h[b + W];
#pragma unroll
				for (int t = b + W + 1; t <= i + W - 1; ++t) pfx = min(pfx, h[t]);
				v = min(sfx, pfx);
			}
			win[i] = v;
		}
		// Probe where the minimizer changes (level 0 in shared memory, then the exact bitmap), carry the
		// answer along otherwise.  Right shifts by constants are written as mul.hi so they run on the
		// FMA pipe: the ALU pipe (shifts, logic, min, compares) is what this kernel saturates.
		uint32_t pass = 0, bit = 0;
#pragma unroll
		for (int i = 0; i < 32; ++i)
			gate_probe_step<SH>(win[i], i ? win[i - 1] : ~win[0], s_l0_addr, level1, P.four, bit, pass);
		pass = __brev(pass);                 // steps pushed position 0 first, so it ended up at bit 31
		pass &= valid;

		uint32_t pass2 = 0;
		while (pass) {
			const int i = __ffs(pass) - 1;
			pass &= pass - 1;
			uint32_t lo, hi, fw_, fm_;
			cut_kmer<K>(w, i, K, lo, hi);
			filter2_slots(filter_mix(lo, hi), P.filter_shift, fw_, fm_);
			pass2 |= (uint32_t)((__ldg(P.filter + fw_) & fm_) == fm_) << i;
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
