// pair.cuh -- paired-seed count kernel (sm_100a): the production path for 17 <= k <= 31.
//
// One probe per position into an L2-resident bitmap runs at the chip's random-request rate: ncu has
// l1tex__m_l1tex2xbar_req_cycles_active at 97 % (profiles/r01_*, r01v7_*), and the microbenchmark
// (tools/microbench.cu, profiles/r01v8_microbench.txt) shows that wall is one 128-byte LINE request
// per clock per SM, however few of the line's bytes are wanted -- and that thread-block-cluster
// DSMEM probes are slower still.  The only way to go faster is to ask fewer questions per position.
//
// Seeds are M-mers with M = min(14, k - 5), so that a window of k bases contains at least 6 of them.
// Two seeds that start D = 2 positions apart share M - 2 bases.  Index a table by those shared bases
// (M = 14: 4^12 words of 32 bits = 64 MiB) and let the word answer for BOTH seeds: bits 0-15 say
// "the M-mer made of <2 bases b0 b1> + <the shared bases> is a site seed" (role A, bit = b0 | b1 << 2),
// bits 16-31 say "<the shared bases> + <2 bases> is a site seed" (role B).  Every M-mer of every site
// k-mer (both read orientations) is entered twice, once per role.  One 32-bit load then covers the
// windows of two seeds: A (at stream position p) closes the windows starting in [p-5, p], B (at p+2)
// closes [p-3, p+2] -- 8 windows, so the kernel probes one PAIR per 8 positions: 4 loads per
// 32-position chunk, pairs always at local positions 0, 8, 16, 24 (no phase arithmetic).  (For k > 19
// a seed closes more windows than that; the geometry stays the k = 19 one, which is all a pair needs
// to cover its 8.)  The four windows that contain both seeds need both bits, which also makes the
// level-1 filter ~2x more selective.  Windows that pass (~2 %) go on to the k-mer bitmap (level 2)
// and the exact path (reference hash64 + table + atomicAdd), which is what produces the reference's
// counts (src/FingerPrint.hpp:89-103, vendor/KseqHashIterator.hpp:87-139).
#pragma once
#include "kernels.cuh"

namespace ntsm {

constexpr int kPairMaxM = 14;                               // seed length for k >= 19
constexpr int kPairMinK = 17;                               // below that the k-mer bitmap is cut differently (2k <= 32): generic kernel
NTSM_HD int pair_seed_len(int k) { return k - 5 < kPairMaxM ? k - 5 : kPairMaxM; }
NTSM_HD size_t pair_words(int m) { return (size_t)1 << (2 * (m - 2)); }      // one 32-bit word per (M-2)-base core: M = 14 -> 4^12 words

constexpr int kPairFoldDefault = 1;                         // table folded 2:1 by default (M = 14: 32 MiB), see below
constexpr int kPairFoldMax = 4;

// The full table is 64 MiB, and with the 16 MiB k-mer bitmap next to it that is more than ONE of the
// two L2 partitions of a B200 holds: ncu on the unfolded table (profiles/r01v8_*) shows the L2
// sector hit rate at 49 % and 2.5 B/base of DRAM reads instead of 0.38.  Folding drops the top
// `fold` bits of the word index (the high bit(s) of the last shared base: two cores that differ only
// there share a word, their role bits ORed) -- half the footprint per fold bit for twice the
// level-1 false-positive rate, at no instruction cost (the mask is a register either way).
NTSM_HD uint32_t pair_word_mask(int m, int fold) { return (uint32_t)(pair_words(m) >> fold) - 1u; }

// One pair probe.  x = the 16 bases starting at the pair's position (32 bits, stream order).  Returns,
// for the 8 windows the pair closes (bit t <-> window p - 5 + t), which of them may still be a site
// k-mer; 0 without issuing the load when need == 0.  The address is {base_lo + 4 * key, base_hi}: the
// table never crosses a 4 GiB line (checked at load).  BSHIFT = 2 * M when M is a compile-time 14
// (role B's bases are the top four bits of x, no mask needed), else taken from bshift at run time.
template <int BSHIFT>
__device__ __forceinline__ uint32_t pair_probe(uint32_t x, uint32_t need, uint32_t base_lo, uint32_t base_hi, uint32_t off_mask,
                                               uint32_t bshift)
{
	uint32_t w;
	asm("{\n\t"
	    ".reg .pred p;\n\t"
	    ".reg .u32 off, alo;\n\t"
	    ".reg .u64 a1;\n\t"
	    "setp.ne.u32 p, %2, 0;\n\t"
	    "shr.u32 off, %1, 2;\n\t"
	    "and.b32 off, off, %5;\n\t"                   // 4 * ((x >> 4) & word mask)
	    "add.u32 alo, off, %3;\n\t"
	    "mov.b64 a1, {alo, %4};\n\t"
	    "mov.u32 %0, 0;\n\t"
	    "@p ld.global.nc.u32 %0, [a1];\n\t"
	    "}"
	    : "=r"(w)
	    : "r"(x), "r"(need), "r"(base_lo), "r"(base_hi), "r"(off_mask));
	const uint32_t a = 0u - ((w >> (x & 15u)) & 1u);               // all-ones if seed A is marked
	const uint32_t bsel = BSHIFT == 28 ? (x >> 28) : ((x >> bshift) & 15u);
	const uint32_t b = 0u - ((w >> (bsel + 16u)) & 1u);            // all-ones if seed B is marked
	return (a | 0xC0u) & (b | 0x03u) & 0xFFu;                      // A closes t in [0,6), B closes t in [2,8)
}

constexpr int kCandSlots = 32;      // candidates one warp hands round per pass of its tail
constexpr uint32_t kGroupChunks = 31;

// Tail (level 2 + exact path).  About 2 % of positions pass level 1, in runs of 3-4 inside few lanes; a
// per-lane loop walks them with ~2 lanes active and one L2 round trip per step (ncu, round 1: 36 %
// of the long-scoreboard stalls on that one load).  The warp pools its candidates instead: an
// inclusive scan gives every lane its slots, the owners write (lane, position) into 32 shared-memory
// slots, and lane j takes candidate j -- fetching the owner's four words by shuffle -- so all level-2
// probes of a group are in flight together.
__device__ __forceinline__ void pooled_tail(const CountParams &P, const uint32_t (&w)[4], uint32_t pass, uint32_t lane,
                                            uint16_t *cand, uint32_t wshift, uint32_t k, uint32_t &hits)
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
			const uint32_t mix = lo * kMixA + hi * (kMixB << (64 - 2 * k));      // filter_mix for 2k > 32, unmasked words
			if (filter_test(__ldg(P.filter + (mix >> wshift)), mix)) hits += resolve_one(P, lo, hi, k);
		}
		__syncwarp();
	}
}

// Work layout.  A group is 31 chunks (992 positions) handled by lanes 0-30; lane 31 holds the chunk
// after them, which is only the halo of lane 30 (a window reaches k-1 <= 30 positions past its
// start).  That costs one idle lane (no extra issue slots) and buys a loop with no dependence
// between a group and the next one's data, so every lane loads its words a whole iteration before
// they are used.  Groups are dealt round-robin over all warps of the grid: at any moment the whole
// grid reads one narrow band of the stream (contiguous per-warp runs were 15 % slower -- thousands of
// separate streams cost TLB reach and DRAM page locality).
//
// Pair q of a chunk sits at local position 8 q and closes the windows starting at local positions
// [8q - 5, 8q + 3): in "pair coordinates" t = i + 5 that is t in [8q, 8q + 8).  The five windows
// before position 0 belong to the previous lane (their validity comes over by shuffle and decides
// whether pair 0 is needed at all); this lane's last five windows are closed by the next lane's
// pair 0, whose answer comes over by shuffle.  Lane 31 therefore probes its pair 0 for lane 30, and
// lane 0 of the warp that owns that chunk probes it again for its own windows 0-2.
//
// K = 19 is the compile-time instantiation of the reference's default; K = 0 takes 17 <= k <= 31 from
// P.k at run time (seed length through P.pair_bshift / P.pair_off_mask), same code.
template <int K, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) count_kernel_pair(const CountParams P)
{
	static_assert(K == 0 || (K >= 19 && K <= 31), "compile-time K uses the M = 14 probe");
	__shared__ uint16_t s_cand[THREADS / 32][kCandSlots];
	const uint32_t k = K ? (uint32_t)K : P.k;
	const uint32_t base_lo = (uint32_t)(uintptr_t)P.pair, base_hi = (uint32_t)((uintptr_t)P.pair >> 32);
	const uint32_t wshift = P.filter_shift + 5;
	const uint32_t off_mask = P.pair_off_mask, bshift = P.pair_bshift;
	constexpr int BS = K ? 28 : 0;
	const uint32_t lane = threadIdx.x & 31;
	uint16_t *cand = s_cand[threadIdx.x >> 5];
	uint32_t tk = 0, hits = 0;

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
		const uint32_t m1 = __shfl_down_sync(0xffffffffu, m0, 1);
		if (lane == 31 || c >= P.n_chunks) m0 = 0xFFFFFFFFu;    // lane 31 is halo only; nothing starts in the padding
		const uint32_t w[4] = { own.x, own.y, nxt.x, nxt.y };
		const uint32_t valid = valid_windows(m0, m1, k);
		tk += __popc(valid);

		// valid windows in pair coordinates, the previous lane's last five included
		uint32_t pv = __shfl_up_sync(0xffffffffu, valid, 1);
		if (lane == 0) pv = 0;                                    // closed by lane 31 of the warp that owns that chunk
		const uint32_t n5 = __funnelshift_l(pv, valid, 5);        // bit i + 5 <-> window i, i = -5 .. 26

		const uint32_t r0 = pair_probe<BS>(own.x, n5 & 0xFFu, base_lo, base_hi, off_mask, bshift);
		const uint32_t r1 = pair_probe<BS>(__funnelshift_r(own.x, own.y, 16), n5 & 0xFF00u, base_lo, base_hi, off_mask, bshift);
		const uint32_t r2 = pair_probe<BS>(own.y, n5 & 0xFF0000u, base_lo, base_hi, off_mask, bshift);
		const uint32_t r3 = pair_probe<BS>(__funnelshift_r(own.y, nxt.x, 16), n5 & 0xFF000000u, base_lo, base_hi, off_mask, bshift);
		const uint32_t nb = __shfl_down_sync(0xffffffffu, r0, 1);   // the next chunk's pair 0 closes windows 27-31
		const uint32_t plo = r0 | (r1 << 8) | (r2 << 16) | (r3 << 24);
		const uint32_t pass = __funnelshift_r(plo, nb, 5) & valid;

		pooled_tail(P, w, pass, lane, cand, wshift, k, hits);
	}
	flush_tallies(tk, hits, P.totals);
}

// ---------------------------------------------------------------------------------------------
// Wide paired seeds: an EXPERIMENT for panels whose tables cannot stay in L2 anyway (BASELINE cfg 5:
// 10^6 sites, 26 M k-mers, 64 M distinct seeds), kept as option "kernel" = 2 with its parity tests.
// It is NOT the default for such panels: measured 295 Gbases/s against 319 for count_kernel_pair.
//
// ncu on that panel with the 14-mer table (profiles/r02a_cfg5_count_ncu_summary.txt): DRAM-bound --
// 72 % of peak DRAM throughput, 18.5 bytes read per base, L2 sector hit rate 16 %.  Every level-1
// word, level-2 word and exact-table slot is a random access that misses L2.  The idea here: a sparser
// level 1 sized for HBM instead of L2, so that level 2 and the exact table all but vanish:
//   seeds are 16-mers 4 apart (A at p, B at p + 4), sharing 12 bases; those index one 32-byte entry
//   -- one sector: a 128-bit map of A's four leading bases and a 128-bit map of B's four trailing
//   bases (8 selector bits folded to 7).  4^12 entries = 512 MiB, ~3 % full for 10^6 sites.  A closes
//   the windows starting in [p-3, p], B closes [p+1, p+4]: 8 windows per entry, pairs at local
//   positions 0, 8, 16, 24 as above.
// What the three versions taught (profiles/r02h_cfg5_wide_ncu_summary.txt, r02i_cfg5_wide_ncu_summary.txt):
//   1. 64-byte entries, the two maps in separate sectors, loads behind branches: 172 Gbases/s.  A thread
//      had one load in flight instead of eight, and both sectors of an entry missed at once.
//   2. predicated loads: eight predicates are more than a thread has; ptxas strings the loads out again.
//   3. (this one) one sector per entry, unconditional loads (a probe nobody needs reads word 0): all eight
//      in flight, L1 serves the second map of an entry (41 % L1 hit rate), L2 sector requests fall by 40 % --
//      and DRAM bytes stay where they were: 69.8 GB against 73.8 GB per 4 Gbases.  An L2 miss costs a whole
//      128-byte line of DRAM traffic here whichever sector was asked for (2.32 G sectors read by L2 for 0.63 G
//      that missed; cudaLimitMaxL2FetchGranularity changes nothing, profiles/r02b_l2fetch_granularity.jsonl),
//      so the wall for this panel is HBM bandwidth itself -- the pair kernel moves 5.9 TB/s, 91 % of the measured
//      copy peak, ~46 G line fetches per second, ~4.5 per 32 positions with either table -- and the wide
//      kernel's 48 registers leave it half the warps to hide the latency with (63 vs 72 % of nominal DRAM peak).  Beating it needs fewer MISSES per position, i.e. a selective level 1 that is
//      L2-resident, which 64 M seeds do not allow in ~48 MiB.
// k >= 19 (a window must hold a 16-mer at 4 consecutive offsets).  Same tail, same exact path.
constexpr int kWideM = 16;
constexpr size_t kWideEntries = (size_t)1 << 24;            // 4^12 cores
constexpr size_t kWideBytes = kWideEntries * 32;

NTSM_HD void wide_slots_a(uint32_t v, uint32_t &word, uint32_t &bit)      // 16-mer v in role A: last 12 bases shared
{
	const uint32_t a = v & 0x7Fu;
	word = (v >> 8) * 8 + (a >> 5);
	bit = a & 31u;
}
NTSM_HD void wide_slots_b(uint32_t v, uint32_t &word, uint32_t &bit)      // 16-mer v in role B: first 12 bases shared
{
	const uint32_t b = (v >> 24) & 0x7Fu;
	word = (v & 0xFFFFFFu) * 8 + 4 + (b >> 5);
	bit = b & 31u;
}

// The four wide probes of a chunk.  x[q] = the 16 bases from pair q's position, t8[q] = the 4 bases after
// them (8 bits), need = valid windows in pair coordinates (bit 8q + t <-> window 8q - 3 + t).  Returns
// the windows that may still be site k-mers.  The table lives in HBM: a thread that waits for each DRAM
// round trip in turn is eight times slower than one that has its eight loads in flight together.  So
// nothing here is conditional -- a probe nobody needs reads word 0 of the table (always cached) instead
// of being branched or predicated around (eight predicates are more than an SM has per thread, and
// ptxas then strings the loads out again) -- and the answers are masked by `need` at the end.
__device__ __forceinline__ uint32_t wide_probe4(const uint32_t (&x)[4], const uint32_t (&t8)[4], uint32_t need, const uint32_t *__restrict__ tab)
{
	uint32_t wa[4], wb[4];
#pragma unroll
	for (int q = 0; q < 4; ++q) {
		const uint32_t entry = (x[q] >> 8) << 3;                       // word index of the 32-byte entry
		const uint32_t ia = entry + ((x[q] & 0x7Fu) >> 5), ib = entry + 4u + ((t8[q] & 0x7Fu) >> 5);
		wa[q] = __ldg(tab + ((need >> (8 * q)) & 0x0Fu ? ia : 0u));
		wb[q] = __ldg(tab + ((need >> (8 * q)) & 0xF0u ? ib : 0u));
	}
	uint32_t r = 0;
#pragma unroll
	for (int q = 0; q < 4; ++q) {
		const uint32_t ra = (0u - ((wa[q] >> (x[q] & 31u)) & 1u)) & 0x0Fu;       // bit (a & 31) of A's word, a = x & 0x7F
		const uint32_t rb = (0u - ((wb[q] >> (t8[q] & 31u)) & 1u)) & 0xF0u;
		r |= (ra | rb) << (8 * q);
	}
	return r & need;
}

template <int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) count_kernel_wide(const CountParams P)
{
	__shared__ uint16_t s_cand[THREADS / 32][kCandSlots];
	const uint32_t k = P.k;
	const uint32_t wshift = P.filter_shift + 5;
	const uint32_t lane = threadIdx.x & 31;
	uint16_t *cand = s_cand[threadIdx.x >> 5];
	uint32_t tk = 0, hits = 0;

	const uint64_t n_groups = (P.n_chunks + kGroupChunks - 1) / kGroupChunks;
	const uint64_t n_warps = (uint64_t)gridDim.x * (THREADS / 32);
	const uint64_t gw = (uint64_t)blockIdx.x * (THREADS / 32) + (threadIdx.x >> 5);

	uint2 own_n = make_uint2(0, 0);
	uint32_t m0_n = 0xFFFFFFFFu;
	if (gw < n_groups && gw * kGroupChunks + lane <= P.n_chunks) {
		own_n = __ldcs(P.bases + gw * kGroupChunks + lane);
		m0_n = __ldcs(P.nmask + gw * kGroupChunks + lane);
	}
	for (uint64_t g = gw; g < n_groups; g += n_warps) {
		const uint64_t c = g * kGroupChunks + lane;
		const uint64_t cn = c + n_warps * kGroupChunks;
		const uint2 own = own_n;
		uint32_t m0 = m0_n;
		own_n = make_uint2(0, 0);
		m0_n = 0xFFFFFFFFu;
		if (g + n_warps < n_groups && cn <= P.n_chunks) {
			own_n = __ldcs(P.bases + cn);
			m0_n = __ldcs(P.nmask + cn);
		}
		uint2 nxt;
		nxt.x = __shfl_down_sync(0xffffffffu, own.x, 1);
		nxt.y = __shfl_down_sync(0xffffffffu, own.y, 1);
		const uint32_t m1 = __shfl_down_sync(0xffffffffu, m0, 1);
		if (lane == 31 || c >= P.n_chunks) m0 = 0xFFFFFFFFu;
		const uint32_t w[4] = { own.x, own.y, nxt.x, nxt.y };
		const uint32_t valid = valid_windows(m0, m1, k);
		tk += __popc(valid);

		// pair q (local position 8q) closes the windows starting at [8q - 3, 8q + 5): in pair coordinates
		// t = i + 3 that is t in [8q, 8q + 8).  The three windows before position 0 belong to the previous
		// lane; this lane's last three are closed by the next lane's pair 0.
		uint32_t pv = __shfl_up_sync(0xffffffffu, valid, 1);
		if (lane == 0) pv = 0;
		const uint32_t n3 = __funnelshift_l(pv, valid, 3);        // bit i + 3 <-> window i, i = -3 .. 28

		const uint32_t xs[4] = { own.x, __funnelshift_r(own.x, own.y, 16), own.y, __funnelshift_r(own.y, nxt.x, 16) };
		const uint32_t ts[4] = { own.y & 0xFFu, (own.y >> 16) & 0xFFu, nxt.x & 0xFFu, (nxt.x >> 16) & 0xFFu };
		const uint32_t plo = wide_probe4(xs, ts, n3, P.pair);
		const uint32_t nb = __shfl_down_sync(0xffffffffu, plo, 1);  // the next chunk's pair 0 closes windows 29-31
		const uint32_t pass = __funnelshift_r(plo, nb, 3) & valid;

		pooled_tail(P, w, pass, lane, cand, wshift, k, hits);
	}
	flush_tallies(tk, hits, P.totals);
}

}  // namespace ntsm
