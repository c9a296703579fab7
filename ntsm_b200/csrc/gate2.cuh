// gate2.cuh -- second generation of the minimizer-gated count kernel (sm_100a).
//
// Same algorithm as count_kernel_gate (kernels.cuh): per position of the packed stream, the
// minimizer (smallest multiplicative hash among the W = K-M+1 M-mers of the k-mer) is looked up
// only where it changes -- first in a 224 KiB hashed bitmap in shared memory (level 0), survivors
// in the exact 4^M-bit bitmap in L2 (level 1) -- and the answer is carried along the positions
// that share the minimizer.  Positions that pass go to the k-mer bitmap (level 2) and then to the
// exact path (reference hash64 + open-addressing table + atomicAdd), which is what produces the
// reference's result (src/FingerPrint.hpp:89-103 over vendor/KseqHashIterator.hpp:87-139).
//
// What changed, all from the SASS / ncu of the first version (profiles/r01v4_*): it issued
// 23.3 instructions per position in the straight-line part, 17.8 of them in the probe step, and
// the ALU pipe (shift/logic/compare) was the saturated unit (67 %).  Here the probe step is
// 15 instructions, 5 of them ALU:
//   * the level-1 address is a 32-bit multiply-add on the low word of the base pointer (the bitmap
//     never crosses a 4 GiB line) -- it was IMAD.WIDE + IADD3 + IADD3.X, two of them ALU;
//   * the level-0 word comes straight from the hash (no second multiply), both levels shift by the
//     low five bits of the minimizer id, and level 0 needs no zeroed register (pc masks the test);
//   * level 2 cuts the k-mer without masks (the mix multiplies the high word by B << (64-2K), which
//     discards the bits above the k-mer).
// M is a template parameter (12..14) so the L2-request / selectivity trade-off can be measured.
#pragma once
#include "kernels.cuh"

namespace ntsm {

constexpr uint32_t kG2MixA = 0x9E3779B1u, kG2MixB = 0x85EBCA6Bu, kG2MixC = 0xC2B2AE35u;

// hashed id (0 .. 4^m - 1) of the minimizer of stream-order k-mer s
NTSM_HD uint32_t gate2_minimizer_id(uint64_t s, int k, int m)
{
	const int sh = 32 - 2 * m;
	const uint32_t mul = kMinHashMul << sh;
	uint32_t best = 0xFFFFFFFFu;
	for (int j = 0; j + m <= k; ++j) {
		const uint32_t hj = (uint32_t)(s >> (2 * j)) * mul;
		best = hj < best ? hj : best;
	}
	return best >> sh;
}
// where id lives: level-0 word (shared memory), level-1 word (global), and the bit inside either word.
// The kernel shifts the loaded word LEFT by (id & 31), which brings bit 31 - (id & 31) to the sign.
NTSM_HD void gate2_slots(uint32_t id, int m, uint32_t &l0_word, uint32_t &l1_word, uint32_t &bit)
{
	const uint32_t cur = id << (32 - 2 * m);
	l0_word = mulhi_u32(cur * kL0Mul, kL0Words);
	l1_word = id >> 5;
	bit = 31 - (id & 31);
}
// level 2: word and two-bit mask of stream-order k-mer s (2k > 32) in a filter of 2^filter_bits bits
NTSM_HD void gate2_filter_slots(uint32_t lo, uint32_t hi, int k, uint32_t filter_shift, uint32_t &word, uint32_t &ra, uint32_t &rb)
{
	const uint32_t mix = lo * kG2MixA + hi * (kG2MixB << (64 - 2 * k));
	word = mix >> (filter_shift + 5);
	const uint32_t t = mix * kG2MixC;
	ra = t & 31;                 // the kernel tests (rot_r(v, ra) & rot_r(v, rb) & 1): bits ra and rb of the word
	rb = (t >> 5) & 31;
}

// One position of the gated probe as PTX (so that both loads stay predicated and the multiplies
// stay multiplies).  `bit` carries the last probe result in its sign bit, `pass` collects one bit
// per position (the sign of `bit` is shifted in at the bottom, so after 32 steps position i sits
// at bit 31-i).  The level-1 address is {base_lo + 4*word, base_hi}: the allocation never crosses
// a 4 GiB line (checked at load), so no carry and no 64-bit add.  w0 is written only by its
// predicated load; whatever it holds otherwise is masked by pc in the setp.
template <int SH>
__device__ __forceinline__ void gate2_step(uint32_t cur, uint32_t prev, uint32_t s_l0_addr, uint32_t base_lo, uint32_t base_hi,
                                           uint32_t four, uint32_t &bit, uint32_t &pass)
{
	asm("{\n\t"
	    ".reg .pred pc, pm;\n\t"
	    ".reg .u32 id, u, a0, w0, x0, i1, alo, w1;\n\t"
	    ".reg .u64 a1;\n\t"
	    "setp.ne.u32 pc, %2, %3;\n\t"                 // minimizer changed?                     ALU
	    "mul.hi.u32 id, %2, %6;\n\t"                  // id = cur >> SH                         FMA
	    "mul.lo.u32 u, %2, %7;\n\t"                   // level-0 hash                           FMA
	    "mul.hi.u32 a0, u, %8;\n\t"                   // level-0 word                           FMA
	    "mad.lo.u32 a0, a0, %10, %4;\n\t"             //                                        FMA
	    "@pc ld.shared.u32 w0, [a0];\n\t"
	    "shf.l.wrap.b32 x0, w0, w0, id;\n\t"          // wanted bit -> sign position            ALU
	    "setp.lt.and.s32 pm, x0, 0, pc;\n\t"          // level 0 says maybe                     ALU
	    "mul.hi.u32 i1, %2, %9;\n\t"                  // level-1 word = id >> 5                 FMA
	    "mad.lo.u32 alo, i1, %10, %5;\n\t"            //                                        FMA
	    "mov.b64 a1, {alo, %11};\n\t"
	    "mov.u32 w1, 0;\n\t"
	    "@pm ld.global.nc.u32 w1, [a1];\n\t"
	    "@pc shf.l.wrap.b32 %0, w1, w1, id;\n\t"      // new answer where the minimizer changed ALU
	    "shf.l.wrap.b32 %1, %0, %1, 1;\n\t"           // pass = pass << 1 | bit >> 31           ALU
	    "}"
	    : "+r"(bit), "+r"(pass)
	    : "r"(cur), "r"(prev), "r"(s_l0_addr), "r"(base_lo), "r"(1u << (32 - SH)), "r"(kL0Mul), "r"(kL0Words),
	      "r"(1u << (32 - SH - 5)), "r"(four), "r"(base_hi));
}

// exact path for one candidate k-mer (stream order, lo = bases 0-15, hi = the rest): canonical value,
// the reference's hash64, open-addressing probe, atomicAdd.  Returns 1 on a table hit.
template <int K>
__device__ __forceinline__ uint32_t resolve_one(const CountParams &P, uint32_t lo, uint32_t hi)
{
	const uint64_t kmask = kmer_mask(K);
	const uint64_t s = (((uint64_t)hi << 32) | lo) & kmask;
	const uint64_t fw = stream_to_fw(s, K), rv = stream_to_rv(s, kmask);
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

// ---------------------------------------------------------------------------------------------
// Device-side build of the lookup structures (the insert loop of FingerPrint::initCountsHash,
// src/FingerPrint.hpp:506-552; the host has already applied first-wins / dupes / -d).  One thread
// per listed k-mer: claim a slot of the open-addressing table with a 64-bit CAS on the key, then
// set the k-mer's bits (both read orientations) in the bitmaps the chosen count kernel probes.
struct BuildParams {
	const uint64_t *hash;       // [n_kmers] reference hash64 values, dense (site-list) order
	const uint8_t *erased;      // [n_kmers] 1 = listed but not in the table
	uint32_t n_kmers, k;
	int variant, gate_m;
	TableSlot *table;           // pre-set to 0xFF bytes (key = kEmptyKey)
	uint32_t table_mask;
	uint32_t *filter;
	uint32_t filter_shift;
	uint32_t *level1, *level0;  // zeroed; nullptr when the variant has none
	uint32_t pair_word_mask;    // variant 5: words of the (folded) pair table - 1
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
		const uint32_t lo = (uint32_t)s, hi = (uint32_t)(s >> 32);
		uint32_t l0w = 0, l1w = 0, gbit = 0, fw2 = 0, fm2 = 0;
		if (B.variant <= 1) {
			const uint32_t ix = filter_mix(lo, hi) >> B.filter_shift;
			atomicOr(B.filter + (ix >> 5), 1u << (ix & 31));
			if (B.variant == 1) {
				const uint32_t mm = minimizer_of(s, (int)B.k, kMinimizerM);
				atomicOr(B.level1 + (mm >> 5), 1u << (mm & 31));
			}
		} else if (B.variant == 2) {
			gate_slots(gate_minimizer_id(s, (int)B.k), l0w, l1w, gbit);
			atomicOr(B.level1 + l1w, 1u << gbit);
			atomicOr(B.level0 + l0w, 1u << gbit);
			filter2_slots(filter_mix(lo, hi), B.filter_shift, fw2, fm2);
			atomicOr(B.filter + fw2, fm2);
		} else if (B.variant == 4) {
			// seed kernel (seed.cuh): every M-mer of the k-mer, direct index; word = v >> 5, bit = 31 - (v & 31)
			const uint32_t mm = (uint32_t)((1ull << (2 * B.gate_m)) - 1);
			for (int j = 0; j + B.gate_m <= (int)B.k; ++j) {
				const uint32_t v = (uint32_t)(s >> (2 * j)) & mm;
				atomicOr(B.level1 + (v >> 5), 1u << (31 - (v & 31)));
			}
			uint32_t ra, rb;
			gate2_filter_slots(lo, hi, (int)B.k, B.filter_shift, fw2, ra, rb);
			atomicOr(B.filter + fw2, (1u << ra) | (1u << rb));
		} else if (B.variant == 5) {
			// paired-seed kernel (pair.cuh): every 14-mer of the k-mer entered once per role (pair_slots)
			for (int j = 0; j + 14 <= (int)B.k; ++j) {
				const uint32_t v = (uint32_t)(s >> (2 * j)) & 0x0FFFFFFFu;
				atomicOr(B.level1 + ((v >> 4) & B.pair_word_mask), 1u << (v & 15));
				atomicOr(B.level1 + (v & 0xFFFFFFu & B.pair_word_mask), 1u << (16 + (v >> 24)));
			}
			uint32_t ra, rb;
			gate2_filter_slots(lo, hi, (int)B.k, B.filter_shift, fw2, ra, rb);
			atomicOr(B.filter + fw2, (1u << ra) | (1u << rb));
		} else {
			gate2_slots(gate2_minimizer_id(s, (int)B.k, B.gate_m), B.gate_m, l0w, l1w, gbit);
			atomicOr(B.level1 + l1w, 1u << gbit);
			atomicOr(B.level0 + l0w, 1u << gbit);
			uint32_t ra, rb;
			gate2_filter_slots(lo, hi, (int)B.k, B.filter_shift, fw2, ra, rb);
			atomicOr(B.filter + fw2, (1u << ra) | (1u << rb));
		}
	}
}

constexpr int kCandSlots = 32;      // candidates one warp hands round per pass of its tail

// Tail shared by the gated and the seed kernels: the warp pools its candidate positions (bits of
// `pass`, one word per lane over the lane's window w[0..3]), lane j takes candidate j, probes the
// k-mer bitmap (level 2) and resolves survivors through the exact table.
template <int K>
__device__ __forceinline__ void pooled_tail(const CountParams &P, const uint32_t (&w)[4], uint32_t pass, uint32_t lane,
                                            uint16_t *cand, uint32_t wshift, uint32_t &hits)
{
	const uint32_t cnt = __popc(pass);
	uint32_t incl = cnt;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
		if (lane >= (uint32_t)d) incl += t;
	}
	const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
	uint32_t idx = incl - cnt;                                // slot of this lane's next candidate
	for (uint32_t r0 = 0; r0 < total; r0 += kCandSlots) {     // warp-uniform trip count, 1 almost always
		while (pass && idx < r0 + kCandSlots) {
			uint32_t i;
			asm("bfind.u32 %0, %1;" : "=r"(i) : "r"(pass));  // highest set bit (one FLO)
			pass ^= 1u << i;
			cand[idx - r0] = (uint16_t)((lane << 5) | i);
			++idx;
		}
		__syncwarp();
		const bool mine = r0 + lane < total;
		const uint32_t e = mine ? cand[lane] : (lane << 5);
		const uint32_t src = e >> 5, i = e & 31;
		const uint32_t y0 = __shfl_sync(0xffffffffu, w[0], src), y1 = __shfl_sync(0xffffffffu, w[1], src);
		const uint32_t y2 = __shfl_sync(0xffffffffu, w[2], src), y3 = __shfl_sync(0xffffffffu, w[3], src);
		if (mine) {
			const bool up = i >= 16;
			const uint32_t x0 = up ? y1 : y0, x1 = up ? y2 : y1, x2 = up ? y3 : y2;
			const uint32_t lo = __funnelshift_r(x0, x1, 2 * i), hi = __funnelshift_r(x1, x2, 2 * i);
			const uint32_t mix = lo * kG2MixA + hi * (kG2MixB << (64 - 2 * K));
			const uint32_t v = __ldg(P.filter + (mix >> wshift));
			const uint32_t t = mix * kG2MixC;
			// level 2: bits (t & 31) and (t >> 5 & 31) of the word both set?
			if (__funnelshift_r(v, v, t) & __funnelshift_r(v, v, t >> 5) & 1u) hits += resolve_one<K>(P, lo, hi);
		}
		__syncwarp();
	}
}

// Work layout.  A group is 31 chunks (992 positions) handled by lanes 0-30; lane 31 holds the chunk after them,
// which is only the halo of lane 30 (a window reaches K-1 <= 30 positions past its start).  That
// costs one idle lane (no extra issue slots: the warp executes the same instructions) and buys a
// loop with no dependence between a group and the next one's data, so every lane loads its words
// a whole iteration before they are used (ncu on the first version: the streaming loads, used
// right after they were issued, were 32 % of all long-scoreboard stalls).
//
// Tail (level 2 + exact path).  About 1.5 % of positions pass the minimizer levels, in runs of 3-4
// inside few lanes; a per-lane loop walks them with ~2 lanes active and one L2 round trip per
// step (ncu: 36 % of the long-scoreboard stalls on that one load).  With POOL the warp pools its
// candidates instead: an inclusive scan gives every lane its slots, the owners write
// (lane, position) into 32 shared-memory slots, and lane j takes candidate j -- fetching the
// owner's four words by shuffle -- so all level-2 probes of a group are in flight together.
constexpr uint32_t kGroupChunks = 31;

template <int K, int M, bool POOL, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) count_kernel_gate2(const CountParams P)
{
	constexpr int W = K - M + 1;
	constexpr int NH = 32 + W - 1;
	constexpr int SH = 32 - 2 * M;
	constexpr uint32_t kHashMul = kMinHashMul << SH;
	static_assert(2 * K > 32 && K <= 31, "level 2 cuts the k-mer as one full word plus 2K-32 bits");
	static_assert(M >= 11 && M <= 15 && W >= 2 && W <= 17 && 2 * (NH - 1) + 32 <= 128, "window does not fit the 128-bit register view");
	extern __shared__ uint32_t s_l0[];
	__shared__ uint16_t s_cand[THREADS / 32][kCandSlots];
	{
		const uint4 *src = reinterpret_cast<const uint4 *>(P.level0);
		uint4 *dst = reinterpret_cast<uint4 *>(s_l0);
		for (uint32_t i = threadIdx.x; i < kL0Words / 4; i += THREADS) dst[i] = __ldg(src + i);
	}
	__syncthreads();
	const uint32_t base_lo = (uint32_t)(uintptr_t)P.minimizer2, base_hi = (uint32_t)((uintptr_t)P.minimizer2 >> 32);
	const uint32_t s_l0_addr = (uint32_t)__cvta_generic_to_shared(s_l0);
	const uint32_t wshift = P.filter_shift + 5;
	const uint32_t lane = threadIdx.x & 31;
	uint16_t *cand = s_cand[threadIdx.x >> 5];
	uint32_t tk = 0, hits = 0;

	// groups are dealt round-robin over all warps of the grid: at any moment the whole grid reads one
	// narrow band of the stream (contiguous per-warp runs were 15 % slower -- thousands of separate
	// streams cost TLB reach and DRAM page locality)
	const uint64_t n_groups = (P.n_chunks + kGroupChunks - 1) / kGroupChunks;
	const uint64_t n_warps = (uint64_t)gridDim.x * (THREADS / 32);
	const uint64_t gw = (uint64_t)blockIdx.x * (THREADS / 32) + (threadIdx.x >> 5);

	// chunk n_chunks is padding and always readable (ntsm_padded_positions); anything later reads as invalid
	uint2 own_n = make_uint2(0, 0);
	uint32_t m0_n = 0xFFFFFFFFu;
	if (gw < n_groups && gw * kGroupChunks + lane <= P.n_chunks) {
		own_n = __ldcs(P.bases + gw * kGroupChunks + lane);
		m0_n = __ldcs(P.nmask + gw * kGroupChunks + lane);
	}
	for (uint64_t g = gw; g < n_groups; g += n_warps) {
		const uint64_t c = g * kGroupChunks + lane;
		const uint64_t cn = c + n_warps * kGroupChunks;           // this lane's chunk in the warp's next group
		const uint2 own = own_n;
		uint32_t m0 = m0_n;
		own_n = make_uint2(0, 0);
		m0_n = 0xFFFFFFFFu;
		if (g + n_warps < n_groups && cn <= P.n_chunks) {         // loaded now, used one iteration from now
			own_n = __ldcs(P.bases + cn);
			m0_n = __ldcs(P.nmask + cn);
		}
		uint2 nxt;
		nxt.x = __shfl_down_sync(0xffffffffu, own.x, 1);
		nxt.y = __shfl_down_sync(0xffffffffu, own.y, 1);
		uint32_t m1 = __shfl_down_sync(0xffffffffu, m0, 1);
		if (lane == 31 || c >= P.n_chunks) m0 = 0xFFFFFFFFu;    // lane 31 is halo only; nothing starts in the padding
		const uint32_t w[4] = { own.x, own.y, nxt.x, nxt.y };
		const uint32_t valid = valid_windows(m0, m1, K);
		uint32_t pass = 0;
		if (valid) {
			tk += __popc(valid);
			// Two halves of 16 positions, as a real loop: the live set is 21 hashes + 16 minima instead
			// of 37 + 32, which is what lets 1024 threads fit their 64 registers without spilling (with
			// 224 KiB of shared memory there is next to no L1 left: every spill reload is an L2 round trip).
			uint32_t bit = 0, prevw = 0;
#pragma unroll 1
			for (int half = 0; half < 2; ++half) {
				const uint32_t x0 = half ? w[1] : w[0], x1 = half ? w[2] : w[1], x2 = half ? w[3] : w[2];
				// hash of the M-mer starting at each position: the multiplier's low SH zero bits push the
				// bases beyond the M-mer out of the word, so no mask is needed
				constexpr int NHH = 16 + W - 1;
				uint32_t h[NHH];
#pragma unroll
				for (int j = 0; j < NHH; ++j) {
					const int sh = (2 * j) & 31;
					h[j] = (j < 16 ? __funnelshift_r(x0, x1, sh) : __funnelshift_r(x1, x2, sh)) * kHashMul;
				}
				// sliding minimum over W consecutive hashes (van Herk / Gil-Werman)
				uint32_t win[16];
#pragma unroll
				for (int i = 0; i < 16; ++i) {
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
				const uint32_t before = half ? prevw : ~win[0];       // position 0 always counts as a change
#pragma unroll
				for (int i = 0; i < 16; ++i)
					gate2_step<SH>(win[i], i ? win[i - 1] : before, s_l0_addr, base_lo, base_hi, P.four, bit, pass);
				prevw = win[15];
			}
			pass = __brev(pass) & valid;         // steps pushed position 0 first, so it ended up at bit 31
		}

		if (!POOL) {
			// ---- tail, per lane: one candidate at a time ----
			while (pass) {
				uint32_t i;
				asm("bfind.u32 %0, %1;" : "=r"(i) : "r"(pass));  // highest set bit (one FLO)
				pass ^= 1u << i;
				const bool up = i >= 16;
				const uint32_t x0 = up ? w[1] : w[0], x1 = up ? w[2] : w[1], x2 = up ? w[3] : w[2];
				const uint32_t lo = __funnelshift_r(x0, x1, 2 * i), hi = __funnelshift_r(x1, x2, 2 * i);
				const uint32_t mix = lo * kG2MixA + hi * (kG2MixB << (64 - 2 * K));
				const uint32_t v = __ldg(P.filter + (mix >> wshift));
				const uint32_t t = mix * kG2MixC;
				if (__funnelshift_r(v, v, t) & __funnelshift_r(v, v, t >> 5) & 1u) hits += resolve_one<K>(P, lo, hi);
			}
			continue;
		}
		pooled_tail<K>(P, w, pass, lane, cand, wshift, hits);
	}
	flush_tallies(tk, hits, P.totals);
}

}  // namespace ntsm
