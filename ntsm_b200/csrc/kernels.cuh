// kernels.cuh -- device side of the counting path (sm_100a): shared pieces + the generic-k kernel.
//
// What the reference does per base (FingerPrint::insertCount, src/FingerPrint.hpp:89-103, over
// KseqHashIterator::step, vendor/KseqHashIterator.hpp:95-112): decode a byte, roll fw and rv,
// take min, run the 7-stage hash64, look the hash up in a 100 MB robin_map, and bump three
// shared counters with locked RMWs.
//
// What every kernel here does per position of the packed stream:
//   1. the k-mer starting at position p is just 2k contiguous bits of the 2-bit stream, so it is
//      cut out with two funnel shifts -- nothing is rolled;
//   2. pre-filters that hold BOTH orientations of every site k-mer decide "cannot be a site k-mer"
//      for > 98 % of positions without a reverse complement or a min(): the paired-seed table
//      (pair.cuh, k >= 17) and/or the k-mer bitmap (two bits of one word per key);
//   3. survivors are turned into the reference's canonical value (fw = 2-bit groups reversed,
//      rv = ~s & mask, see kmer_math.h), hashed with the reference hash64, and probed in an
//      open-addressing table keyed by that hash; a hit is one atomicAdd into counts[dense index];
//   4. window validity (no N / separator among the k positions) is computed for 32 positions at a
//      time with shifts of the N-mask words; #@TK is the popcount of valid windows.
// The filters may only let too MANY windows through, never too few: the counts come from step 3.
//
// Earlier kernel generations (minimizer-gated, shared-memory level 0, strided seeds) and what ncu
// said about each are in DESIGN.md 4.1 and profiles/r01*; they were removed from the build in round 2.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "kmer_math.h"
#include "table.cuh"

namespace ntsm {


struct CountParams {
	const uint2 *bases;        // 32 positions per element (two little-endian uint32 words)
	const uint32_t *nmask;     // 32 positions per element
	uint64_t n_chunks;         // number of 32-position chunks to scan (arrays hold n_chunks + 2 at least)
	const uint32_t *pair;      // paired-seed table (pair.cuh); nullptr for the generic kernel
	uint32_t pair_off_mask;    // ((words of the pair table) - 1) << 2
	uint32_t pair_bshift;      // 2 * seed length: where role B's two trailing bases start in the 16-base probe word
	const uint32_t *filter;    // k-mer bitmap, 2^filter_bits bits
	uint32_t filter_shift;     // 32 - filter_bits
	const TableSlot *table;
	uint32_t table_mask;       // capacity - 1
	uint32_t k;
	uint32_t delta;            // added to counts[idx] per hit: 1 to count, 0xFFFFFFFF to take a batch back out, 0 to only tally
	uint32_t *counts;
	unsigned long long *totals;   // [0] valid windows (TK), [1] hits
};

constexpr int kCountThreads = 256;
constexpr uint32_t kMixA = 0x9E3779B1u, kMixB = 0x85EBCA6Bu, kMixC = 0xC2B2AE35u;

// k-mer bitmap: word index and the two bit numbers of stream-order k-mer (lo = bases 0-15, hi = the
// rest), shared by the builder and every probe.  The multiplier of hi has its low 64-2k bits clear,
// which pushes whatever lies above the k-mer in hi out of the product, so callers with 2k > 32 need
// not mask; for 2k <= 32 lo must be masked and hi is ignored.
NTSM_HD uint32_t filter_mix(uint32_t lo, uint32_t hi, uint32_t k)
{
	return 2 * k > 32 ? lo * kMixA + hi * (kMixB << (64 - 2 * k)) : lo * kMixA;
}
NTSM_HD void filter_slots(uint32_t mix, uint32_t filter_shift, uint32_t &word, uint32_t &ra, uint32_t &rb)
{
	word = mix >> (filter_shift + 5);
	const uint32_t t = mix * kMixC;
	ra = t & 31;
	rb = (t >> 5) & 31;
}
__device__ __forceinline__ bool filter_test(uint32_t v, uint32_t mix)
{
	const uint32_t t = mix * kMixC;
	return __funnelshift_r(v, v, t) & __funnelshift_r(v, v, t >> 5) & 1u;     // bits (t & 31) and (t >> 5 & 31) both set
}

// exact path for one candidate k-mer (stream order, lo = bases 0-15, hi = the rest): canonical value,
// the reference's hash64, open-addressing probe, atomicAdd.  Returns 1 on a table hit.
__device__ __forceinline__ uint32_t resolve_one(const CountParams &P, uint32_t lo, uint32_t hi, uint32_t k)
{
	const uint64_t kmask = kmer_mask(k);
	const uint64_t s = (((uint64_t)hi << 32) | lo) & kmask;
	const uint64_t fw = stream_to_fw(s, k), rv = stream_to_rv(s, kmask);
	const uint64_t h = hash64(fw < rv ? fw : rv, kmask);         // KseqHashIterator.hpp:102
	uint32_t slot = (uint32_t)(h ^ (h >> 29)) & P.table_mask;
	for (;;) {
		const TableSlot e = P.table[slot];
		if (e.key == h) {
			atomicAdd(P.counts + e.idx, P.delta);                // FingerPrint.hpp:93-94 (delta = 1)
			return 1;
		}
		if (e.key == kEmptyKey) return 0;
		slot = (slot + 1) & P.table_mask;
	}
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

// Generic kernel, any 1 <= k <= 31 taken at run time: one k-mer-bitmap probe per position, survivors
// through the exact path.  Request-bound (one L2 request per position, ~280 Gpos/s); used for k < 17,
// where the paired-seed geometry does not apply, and as the plain cross-check of the pair kernel.
__global__ void __launch_bounds__(kCountThreads) count_kernel_generic(const CountParams P)
{
	const uint32_t k = P.k;
	const uint64_t kmask = kmer_mask(k);
	const uint32_t lo_mask = (uint32_t)kmask;
	const uint32_t wshift = P.filter_shift + 5;
	uint32_t tk = 0, hits = 0;

	const uint64_t stride = (uint64_t)gridDim.x * kCountThreads;
	for (uint64_t c = (uint64_t)blockIdx.x * kCountThreads + threadIdx.x; c < P.n_chunks; c += stride) {
		// 32 own positions + the next 32 (a window may reach k-1 <= 30 positions past its start)
		const uint2 own = __ldg(P.bases + c), nxt = __ldg(P.bases + c + 1);
		const uint32_t m0 = __ldg(P.nmask + c), m1 = __ldg(P.nmask + c + 1);
		const uint32_t valid = valid_windows(m0, m1, k);
		if (valid == 0) continue;
		tk += __popc(valid);
#pragma unroll
		for (int half = 0; half < 2; ++half) {
			const uint32_t x0 = half ? own.y : own.x, x1 = half ? nxt.x : own.y, x2 = half ? nxt.y : nxt.x;
			uint32_t lo[16], hi[16], word[16];
#pragma unroll
			for (int j = 0; j < 16; ++j) {
				lo[j] = __funnelshift_r(x0, x1, 2 * j) & lo_mask;
				hi[j] = __funnelshift_r(x1, x2, 2 * j);
				word[j] = __ldg(P.filter + (filter_mix(lo[j], hi[j], k) >> wshift));
			}
#pragma unroll
			for (int j = 0; j < 16; ++j)
				if (((valid >> (half * 16 + j)) & 1u) && filter_test(word[j], filter_mix(lo[j], hi[j], k)))
					hits += resolve_one(P, lo[j], hi[j], k);
		}
	}
	flush_tallies(tk, hits, P.totals);
}

// ---------------------------------------------------------------------------------------------
// Device-side build of the lookup structures (the insert loop of FingerPrint::initCountsHash,
// src/FingerPrint.hpp:506-552; the host has already applied first-wins / dupes / -d).  One thread
// per listed k-mer: claim a slot of the open-addressing table with a 64-bit CAS on the key, then
// set the k-mer's bits (both read orientations) in the bitmaps the count kernel probes.
struct BuildParams {
	const uint64_t *hash;       // [n_kmers] reference hash64 values, dense (site-list) order
	const uint8_t *erased;      // [n_kmers] 1 = listed but not in the table
	uint32_t n_kmers, k;
	TableSlot *table;           // pre-set to 0xFF bytes (key = kEmptyKey)
	uint32_t table_mask;
	uint32_t *filter;
	uint32_t filter_shift;
	uint32_t *pair;             // zeroed; nullptr when the generic kernel will run
	int wide;                   // 1: `pair` is the wide table (64-byte entries, 16-mers 4 apart: pair.cuh)
	uint32_t pair_m;            // seed length of the pair table
	uint32_t pair_word_mask;    // words of the (folded) pair table - 1
	int *err;                   // [0] 0 ok, 1 hash out of range, 2 duplicate; [1] the offending index
};

__global__ void build_tables_kernel(const BuildParams B)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= B.n_kmers || B.erased[i]) return;
	const uint64_t m = kmer_mask(B.k);
	const uint64_t h = B.hash[i];
	if (h > m) {
		if (atomicCAS(B.err, 0, 1) == 0) B.err[1] = (int)i;
		return;
	}
	uint32_t slot = (uint32_t)(h ^ (h >> 29)) & B.table_mask;
	for (;;) {
		const unsigned long long old = atomicCAS(reinterpret_cast<unsigned long long *>(&B.table[slot].key), kEmptyKey, h);
		if (old == kEmptyKey) {
			B.table[slot].idx = i;
			break;
		}
		if (old == h) {                                            // the host hands over distinct hashes
			if (atomicCAS(B.err, 0, 2) == 0) B.err[1] = (int)i;
			return;
		}
		slot = (slot + 1) & B.table_mask;
	}
	// the canonical k-mer (reference orientation) and the two stream-order spellings a read can show
	const uint64_t canon = hash64_inv(h, m);
	const uint64_t ss[2] = { fw_to_stream(canon, B.k), ~canon & m };
#pragma unroll
	for (int o = 0; o < 2; ++o) {
		const uint64_t s = ss[o];
		uint32_t word, ra, rb;
		filter_slots(filter_mix((uint32_t)s, (uint32_t)(s >> 32), B.k), B.filter_shift, word, ra, rb);
		atomicOr(B.filter + word, (1u << ra) | (1u << rb));
		if (B.pair && B.wide) {
			// wide table (pair.cuh): every 16-mer of the k-mer, once per role, into the 32-byte entry of its 12 shared bases
			for (uint32_t j = 0; j + 16 <= B.k; ++j) {
				const uint32_t v = (uint32_t)(s >> (2 * j));
				const uint32_t a = v & 0x7Fu, b = (v >> 24) & 0x7Fu;
				atomicOr(B.pair + (size_t)(v >> 8) * 8 + (a >> 5), 1u << (a & 31u));                    // role A: v's last 12 bases are shared
				atomicOr(B.pair + (size_t)(v & 0xFFFFFFu) * 8 + 4 + (b >> 5), 1u << (b & 31u));        // role B: v's first 12 bases
			}
		} else if (B.pair) {
			// paired-seed table (pair.cuh): every M-mer of the k-mer entered once per role
			const uint32_t M = B.pair_m, vm = (uint32_t)((1ull << (2 * M)) - 1), core = (uint32_t)((1ull << (2 * (M - 2))) - 1);
			for (uint32_t j = 0; j + M <= B.k; ++j) {
				const uint32_t v = (uint32_t)(s >> (2 * j)) & vm;
				atomicOr(B.pair + ((v >> 4) & B.pair_word_mask), 1u << (v & 15));                      // role A: v's last M-2 bases are the shared ones
				atomicOr(B.pair + (v & core & B.pair_word_mask), 1u << (16 + (v >> (2 * M - 4))));    // role B: v's first M-2 bases
			}
		}
	}
}

// printCountsMax's per-site loop (src/FingerPrint.hpp:281-294): max and sum (mod 2^32) over the
// ref list and over the var list of every site.  One thread per site; lists are 0..13 long.
//
// FUSED COMBINE.  counts[r] is rank r's private count array (n_ranks <= kMaxPeers); ranks other than
// 0 are read straight out of the peer GPU's memory over NVLink (peer access enabled by
// ntsm_group_create) -- the all-reduce and the per-site reduce are one kernel: every k-mer's counts
// are summed across the GPUs first (mod 2^32, like the reference's unsigned), the max is taken
// afterwards (sum of per-GPU maxima would be wrong; that is what `ntsmEval --merge` does,
// src/CompareCounts.hpp:648-657).  With n_ranks == 1 this is the plain per-site reduce.  When
// `summed` is non-null the k-mer-level sums are also written there (ntsm_get_counts after a combine).
constexpr int kMaxPeers = 16;
struct PeerCounts {
	const uint32_t *counts[kMaxPeers];
	int n_ranks;
};

__global__ void site_reduce_kernel(const PeerCounts pc, const uint32_t *__restrict__ allele_off, uint32_t n_sites,
                                   uint32_t *__restrict__ max_ref, uint32_t *__restrict__ max_var,
                                   uint32_t *__restrict__ sum_ref, uint32_t *__restrict__ sum_var, uint32_t *__restrict__ summed)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_sites) return;
	const uint32_t r0 = allele_off[2 * i], r1 = allele_off[2 * i + 1], v1 = allele_off[2 * i + 2];
	uint32_t mr = 0, sr = 0, mv = 0, sv = 0;
	for (uint32_t j = r0; j < v1; ++j) {
		uint32_t c = pc.counts[0][j];
		for (int r = 1; r < pc.n_ranks; ++r) c += pc.counts[r][j];
		if (summed) summed[j] = c;
		if (j < r1) { mr = max(mr, c); sr += c; }
		else { mv = max(mv, c); sv += c; }
	}
	max_ref[i] = mr; max_var[i] = mv; sum_ref[i] = sr; sum_var[i] = sv;
}

struct PeerTotals {
	const unsigned long long *totals[kMaxPeers];
	int n_ranks;
};
// {TK, hits, bases} summed over the group's GPUs into out[0..2] (out may alias totals[0])
__global__ void sum_totals_kernel(const PeerTotals pt, unsigned long long *out)
{
	if (threadIdx.x < 3) {
		unsigned long long s = 0;
		for (int r = 0; r < pt.n_ranks; ++r) s += pt.totals[r][threadIdx.x];
		out[threadIdx.x] = s;
	}
}

__global__ void set_u64_kernel(unsigned long long *p, unsigned long long v) { *p = v; }

// counts[i] += add[i] (mod 2^32): a shard's k-mer-level counts joining this context's (ntsm_add_counts)
__global__ void add_counts_kernel(uint32_t *__restrict__ counts, const uint32_t *__restrict__ add, uint32_t n)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) counts[i] += add[i];
}
__global__ void add_totals_kernel(unsigned long long *totals, unsigned long long tk, unsigned long long hits)
{
	totals[0] += tk;
	totals[1] += hits;
}

}  // namespace ntsm
