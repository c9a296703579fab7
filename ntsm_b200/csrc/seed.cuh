// seed.cuh -- strided-seed count kernel (sm_100a): the production path for k = 19.
//
// The gated kernels (kernels.cuh, gate2.cuh) spend ~28 instructions per position computing a
// minimizer so that they can probe on change only; ncu put them on the ALU pipe (57-67 %) with
// the L1->L2 request path half idle.  This kernel needs no minimizer and no hash at all.
//
// Every site k-mer (K bases) contains W = K - M + 1 M-mers.  The seed bitmap (4^M bits, direct
// index = the M-mer's 2M bits in stream order, L2 resident: 32 MiB for M = 14) marks EVERY M-mer
// of every site k-mer, both read orientations.  A window [i, i+K) of the read stream contains
// exactly one M-mer that starts at a stream position j = 0 (mod W), j in [i, i+W-1]; if that
// M-mer is not in the bitmap the window cannot be a site k-mer.  So the kernel probes one seed
// per W positions -- 16 probes per 96 positions for W = 6 -- and every probe answers for the W
// windows it closes.  Windows whose seed is marked (~2 %) go on to the k-mer bitmap (level 2) and
// the exact path (reference hash64 + table + atomicAdd), unchanged from gate2.cuh, which is what
// produces the reference's counts (src/FingerPrint.hpp:89-103, vendor/KseqHashIterator.hpp:87-139).
//
// Per position that leaves ~2.5 instructions and 1/6 L2 request (less: probes that close no
// valid window are predicated off), so the kernel runs at the chip's random-request rate into L2
// (one request per clock per SM, profiles/r01_microbench_v1.txt) instead of at ALU throughput.
#pragma once
#include "gate2.cuh"

namespace ntsm {

constexpr int kSeedM = 14;

// where stream-order M-mer v (2M bits) lives in the seed bitmap; the kernel shifts the word LEFT by
// (v & 31), which brings bit 31 - (v & 31) to the sign position
NTSM_HD void seed_slots(uint32_t v, uint32_t &word, uint32_t &bit)
{
	word = v >> 5;
	bit = 31 - (v & 31);
}

// One seed probe.  x holds the M-mer in its top 2M = 28 bits (low 4 bits: the two bases before it,
// ignored).  Returns all-ones if the seed is marked, 0 if not or if need == 0 (no load issued).
// The address is {base_lo + 4 * word, base_hi}: the bitmap never crosses a 4 GiB line (checked at
// load).  Constant shifts are mul.hi so they stay off the ALU pipe, as in gate2_step.
__device__ __forceinline__ uint32_t seed_probe(uint32_t x, uint32_t need, uint32_t base_lo, uint32_t base_hi, uint32_t four)
{
	uint32_t r;
	asm("{\n\t"
	    ".reg .pred p;\n\t"
	    ".reg .u32 i1, alo, w1, s;\n\t"
	    ".reg .u64 a1;\n\t"
	    "setp.ne.u32 p, %2, 0;\n\t"
	    "mul.hi.u32 i1, %1, %3;\n\t"                  // word = x >> 9
	    "mad.lo.u32 alo, i1, %4, %5;\n\t"
	    "mov.b64 a1, {alo, %6};\n\t"
	    "mov.u32 w1, 0;\n\t"
	    "@p ld.global.nc.u32 w1, [a1];\n\t"
	    "mul.hi.u32 s, %1, %7;\n\t"                   // x >> 4: low five bits = bit index
	    "shf.l.wrap.b32 w1, w1, w1, s;\n\t"
	    "shr.s32 %0, w1, 31;\n\t"
	    "}"
	    : "=r"(r)
	    : "r"(x), "r"(need), "r"(1u << 23), "r"(four), "r"(base_lo), "r"(base_hi), "r"(1u << 28));
	return r;
}

// Work layout as in count_kernel_gate2: a warp takes groups of 31 chunks (lanes 0-30; lane 31 holds
// the next chunk as halo), groups dealt round-robin over all warps of the grid, each lane's words
// loaded one iteration ahead.
//
// Seeds sit at stream positions that are multiples of W = 6.  Chunk c starts at position 32 c, so
// its first seed is at local position f = (-32 c) mod 6 = {0, 4, 2}[c mod 3]; a lane probes the
// seeds that START inside its chunk (6, 5 or 5 of them: 16 per 96 positions).  The windows after a
// lane's last seed are closed by the next lane's first seed, whose answer comes over by shuffle
// (lane 31 probes its first seed for lane 30; the same seed is probed again by lane 0 of the warp
// that owns that chunk -- 1/31 of the chunks).
//
// Bookkeeping is done in "seed coordinates" t = i + 5 - f (i = local window start): seed q closes
// t in [6q, 6q+6), so its answer is ANDed with a constant mask, and one funnel shift by 5 - f maps
// the 37 bits back to positions.  The same coordinates tell which seeds close a valid window at
// all (the others are not probed: ~12 % of them in 150-bp reads, the windows that run into a
// read separator).
template <int K, int M, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) count_kernel_seed(const CountParams P)
{
	constexpr int W = K - M + 1;
	static_assert(M == 14 && W == 6, "seed positions and masks below are written out for M = 14, W = 6");
	static_assert(2 * K > 32 && K <= 31, "level 2 cuts the k-mer as one full word plus 2K-32 bits");
	__shared__ uint16_t s_cand[THREADS / 32][kCandSlots];
	const uint32_t base_lo = (uint32_t)(uintptr_t)P.minimizer2, base_hi = (uint32_t)((uintptr_t)P.minimizer2 >> 32);
	const uint32_t wshift = P.filter_shift + 5;
	const uint32_t lane = threadIdx.x & 31;
	uint16_t *cand = s_cand[threadIdx.x >> 5];
	uint32_t tk = 0, hits = 0;

	const uint64_t n_groups = (P.n_chunks + kGroupChunks - 1) / kGroupChunks;
	const uint64_t n_warps = (uint64_t)gridDim.x * (THREADS / 32);
	const uint64_t gw = (uint64_t)blockIdx.x * (THREADS / 32) + (threadIdx.x >> 5);
	uint32_t cm3 = (uint32_t)((gw * kGroupChunks + lane) % 3);            // chunk index mod 3
	const uint32_t step3 = (uint32_t)((n_warps * kGroupChunks) % 3);

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
		const uint32_t f = (0x240u >> (4 * cm3)) & 15u;           // {0, 4, 2}[c mod 3]
		cm3 += step3;
		cm3 -= cm3 >= 3 ? 3 : 0;

		uint2 nxt;
		nxt.x = __shfl_down_sync(0xffffffffu, own.x, 1);
		nxt.y = __shfl_down_sync(0xffffffffu, own.y, 1);
		const uint32_t m1 = __shfl_down_sync(0xffffffffu, m0, 1);
		if (lane == 31 || c >= P.n_chunks) m0 = 0xFFFFFFFFu;    // lane 31 is halo only; nothing starts in the padding
		const uint32_t w[4] = { own.x, own.y, nxt.x, nxt.y };
		const uint32_t valid = valid_windows(m0, m1, K);
		tk += __popc(valid);

		// valid windows in seed coordinates, the previous lane's last five included (t < 5 - f .. they
		// belong to this lane's first seed)
		uint32_t pv = __shfl_up_sync(0xffffffffu, valid, 1);
		if (lane == 0) pv = 0;                                    // closed by lane 31 of the warp that owns that chunk
		const uint32_t n5 = __funnelshift_l(pv, valid, 5);        // bit i + 5 <-> window i, i = -5 .. 26
		const uint32_t nlo = __funnelshift_r(n5, valid >> 27, f); // bit t <-> window t - 5 + f
		const uint32_t nhi = valid >> (27 + f);

		// the stream shifted up by two bases, so that a funnel shift by 2 j lands seed j in the top 28 bits
		const uint32_t v0 = own.x << 4, v1 = __funnelshift_l(own.x, own.y, 4), v2 = __funnelshift_l(own.y, nxt.x, 4);
		const uint32_t f2 = 2 * f;
		const uint32_t r0 = seed_probe(__funnelshift_rc(v0, v1, f2), nlo & 0x3Fu, base_lo, base_hi, P.four);
		const uint32_t r1 = seed_probe(__funnelshift_rc(v0, v1, f2 + 12), nlo & 0xFC0u, base_lo, base_hi, P.four);
		const uint32_t r2 = seed_probe(__funnelshift_rc(v0, v1, f2 + 24), nlo & 0x3F000u, base_lo, base_hi, P.four);
		const uint32_t r3 = seed_probe(__funnelshift_rc(v1, v2, f2 + 4), nlo & 0xFC0000u, base_lo, base_hi, P.four);
		const uint32_t r4 = seed_probe(__funnelshift_rc(v1, v2, f2 + 16), nlo & 0x3F000000u, base_lo, base_hi, P.four);
		// seed 5 starts inside this chunk only when f == 0 (local position 30)
		const uint32_t need5 = f == 0 ? ((nlo & 0xC0000000u) | (nhi & 0xFu)) : 0u;
		const uint32_t r5own = seed_probe(__funnelshift_rc(v1, v2, 28), need5, base_lo, base_hi, P.four);
		const uint32_t nb = __shfl_down_sync(0xffffffffu, r0, 1);   // the next chunk's first seed
		const uint32_t r5 = f == 0 ? r5own : nb;
		const uint32_t plo = (r0 & 0x3Fu) | (r1 & 0xFC0u) | (r2 & 0x3F000u) | (r3 & 0xFC0000u) | (r4 & 0x3F000000u) | (r5 & 0xC0000000u);
		const uint32_t phi = (r5 & 0xFu) | (nb & 0x10u);            // t = 36: window 31 of an f == 0 chunk
		const uint32_t pass = __funnelshift_r(plo, phi, 5 - f) & valid;

		pooled_tail<K>(P, w, pass, lane, cand, wshift, hits);
	}
	flush_tallies(tk, hits, P.totals);
}

}  // namespace ntsm
