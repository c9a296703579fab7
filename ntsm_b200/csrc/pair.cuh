// pair.cuh -- paired-seed count kernel (sm_100a): the production path for k = 19.
//
// The strided-seed kernel (seed.cuh) probes one 14-mer per W = 6 positions and runs at the chip's
// random-request rate into L2: ncu has l1tex__m_l1tex2xbar_req_cycles_active at 97 % with the ALU
// pipe at 48 % (profiles/r01v7_*), and the microbenchmark (tools/microbench.cu, profiles/
// r01v8_microbench.txt) shows that wall is one 128-byte LINE request per clock per SM, however few
// of the line's bytes are wanted -- and that thread-block-cluster DSMEM probes are slower still.
// The only way left to go faster is to ask fewer questions per position.
//
// Two seeds that start D = 2 positions apart share M - D = 12 bases.  Index a table by those 12
// shared bases (4^12 words of 32 bits = 64 MiB, L2 resident) and let the word answer for BOTH
// seeds: bits 0-15 say "the 14-mer made of <2 bases b0 b1> + <the 12 shared bases> is a site
// seed" (role A, bit = b0 | b1 << 2), bits 16-31 say "<the 12 shared bases> + <2 bases b14 b15> is
// a site seed" (role B, bit = 16 + (b14 | b15 << 2)).  Every 14-mer of every site k-mer (both read
// orientations) is entered twice, once per role.  One 32-bit load then covers the windows of two
// seeds: A (at stream position p) closes the windows starting in [p-5, p], B (at p+2) closes
// [p-3, p+2] -- 8 windows, so the kernel probes one PAIR per 8 positions: 4 loads per 32-position
// chunk instead of 5.33, pairs always at local positions 0, 8, 16, 24 (no phase arithmetic).  The
// four windows that contain both seeds need both bits, which also makes the level-1 filter ~2x
// more selective (fewer level-2 requests).  Windows that pass (~1 %) go on to the k-mer bitmap
// (level 2) and the exact path (reference hash64 + table + atomicAdd), shared with gate2.cuh, which
// is what produces the reference's counts (src/FingerPrint.hpp:89-103,
// vendor/KseqHashIterator.hpp:87-139).
#pragma once
#include "gate2.cuh"

namespace ntsm {

constexpr int kPairM = 14;                                  // seed length
constexpr int kPairD = 2;                                   // distance between the two seeds of a pair
constexpr size_t kPairWords = (size_t)1 << (2 * (kPairM - kPairD));   // 4^12 words

constexpr int kPairFoldDefault = 1;                         // table folded 2:1 by default (32 MiB), see below
constexpr int kPairFoldMax = 4;

// The full table is 64 MiB, and with the 16 MiB k-mer bitmap next to it that is more than ONE of the
// two L2 partitions of a B200 holds: ncu on the unfolded table (profiles/r01v8_*) shows the L2
// sector hit rate at 49 % and 2.5 B/base of DRAM reads instead of 0.38.  Folding drops the top
// `fold` bits of the word index (the high bit(s) of the last shared base: two cores that differ only
// there share a word, their role bits ORed) -- half the footprint per fold bit for twice the
// level-1 false-positive rate, at no instruction cost (the mask is a register either way).
NTSM_HD uint32_t pair_word_mask(int fold) { return (uint32_t)(kPairWords >> fold) - 1u; }

// the two table entries of stream-order 14-mer v (28 bits, first base in the low bits)
NTSM_HD void pair_slots(uint32_t v, uint32_t word_mask, uint32_t &word_a, uint32_t &bit_a, uint32_t &word_b, uint32_t &bit_b)
{
	word_a = (v >> 4) & word_mask;      // role A: v's last 12 bases are the shared ones
	bit_a = v & 15;
	word_b = v & 0xFFFFFFu & word_mask; // role B: v's first 12 bases are the shared ones
	bit_b = 16 + (v >> 24);
}

// One pair probe.  x = the 16 bases starting at the pair's position (32 bits, stream order).  Returns,
// for the 8 windows the pair closes (bit t <-> window p - 5 + t), which of them may still be a site
// k-mer; 0 without issuing the load when need == 0.  The address is {base_lo + 4 * key, base_hi}: the
// table never crosses a 4 GiB line (checked at load).
__device__ __forceinline__ uint32_t pair_probe(uint32_t x, uint32_t need, uint32_t base_lo, uint32_t base_hi, uint32_t off_mask)
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
	const uint32_t b = 0u - ((w >> ((x >> 28) + 16u)) & 1u);       // all-ones if seed B is marked
	return (a | 0xC0u) & (b | 0x03u) & 0xFFu;                      // A closes t in [0,6), B closes t in [2,8)
}

// Work layout as in count_kernel_gate2 / count_kernel_seed: a warp takes groups of 31 chunks (lanes
// 0-30; lane 31 holds the next chunk as halo), groups dealt round-robin over all warps of the grid,
// each lane's words loaded one iteration ahead.
//
// Pair q of a chunk sits at local position 8 q and closes the windows starting at local positions
// [8q - 5, 8q + 3): in "pair coordinates" t = i + 5 that is t in [8q, 8q + 8).  The five windows
// before position 0 belong to the previous lane (their validity comes over by shuffle and decides
// whether pair 0 is needed at all); this lane's last five windows are closed by the next lane's
// pair 0, whose answer comes over by shuffle.  Lane 31 therefore probes its pair 0 for lane 30, and
// lane 0 of the warp that owns that chunk probes it again for its own windows 0-2.
template <int K, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) count_kernel_pair(const CountParams P)
{
	static_assert(K - kPairM + 1 == 6, "pair geometry (8 windows per pair, masks 0x3F / 0xFC) is written out for K - M = 5, D = 2");
	static_assert(2 * K > 32 && K <= 31, "level 2 cuts the k-mer as one full word plus 2K-32 bits");
	__shared__ uint16_t s_cand[THREADS / 32][kCandSlots];
	const uint32_t base_lo = (uint32_t)(uintptr_t)P.minimizer2, base_hi = (uint32_t)((uintptr_t)P.minimizer2 >> 32);
	const uint32_t wshift = P.filter_shift + 5;
	const uint32_t off_mask = P.pair_word_mask << 2;
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
		const uint32_t valid = valid_windows(m0, m1, K);
		tk += __popc(valid);

		// valid windows in pair coordinates, the previous lane's last five included
		uint32_t pv = __shfl_up_sync(0xffffffffu, valid, 1);
		if (lane == 0) pv = 0;                                    // closed by lane 31 of the warp that owns that chunk
		const uint32_t n5 = __funnelshift_l(pv, valid, 5);        // bit i + 5 <-> window i, i = -5 .. 26

		const uint32_t r0 = pair_probe(own.x, n5 & 0xFFu, base_lo, base_hi, off_mask);
		const uint32_t r1 = pair_probe(__funnelshift_r(own.x, own.y, 16), n5 & 0xFF00u, base_lo, base_hi, off_mask);
		const uint32_t r2 = pair_probe(own.y, n5 & 0xFF0000u, base_lo, base_hi, off_mask);
		const uint32_t r3 = pair_probe(__funnelshift_r(own.y, nxt.x, 16), n5 & 0xFF000000u, base_lo, base_hi, off_mask);
		const uint32_t nb = __shfl_down_sync(0xffffffffu, r0, 1);   // the next chunk's pair 0 closes windows 27-31
		const uint32_t plo = r0 | (r1 << 8) | (r2 << 16) | (r3 << 24);
		const uint32_t pass = __funnelshift_r(plo, nb, 5) & valid;

		pooled_tail<K>(P, w, pass, lane, cand, wshift, hits);
	}
	flush_tallies(tk, hits, P.totals);
}

}  // namespace ntsm
