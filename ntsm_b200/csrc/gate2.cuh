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

template <int K, int M>
__global__ void __launch_bounds__(kGateThreads, 1) count_kernel_gate2(const CountParams P)
{
	constexpr int W = K - M + 1;
	constexpr int NH = 32 + W - 1;
	constexpr int SH = 32 - 2 * M;
	constexpr uint32_t kHashMul = kMinHashMul << SH;
	static_assert(2 * K > 32 && K <= 31, "level 2 cuts the k-mer as one full word plus 2K-32 bits");
	static_assert(M >= 11 && M <= 15 && W >= 2 && 2 * (NH - 1) + 32 <= 128, "window does not fit the 128-bit register view");
	extern __shared__ uint32_t s_l0[];
	{
		const uint4 *src = reinterpret_cast<const uint4 *>(P.level0);
		uint4 *dst = reinterpret_cast<uint4 *>(s_l0);
		for (uint32_t i = threadIdx.x; i < kL0Words / 4; i += kGateThreads) dst[i] = __ldg(src + i);
	}
	__syncthreads();
	const uint32_t base_lo = (uint32_t)(uintptr_t)P.minimizer2, base_hi = (uint32_t)((uintptr_t)P.minimizer2 >> 32);
	const uint32_t s_l0_addr = (uint32_t)__cvta_generic_to_shared(s_l0);
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

		// hash of the M-mer starting at each position: the multiplier's low SH zero bits push the
		// bases beyond the M-mer out of the word, so no mask is needed
		uint32_t h[NH];
#pragma unroll
		for (int j = 0; j < NH; ++j) {
			const int a = j >> 4, sh = (2 * j) & 31;
			h[j] = __funnelshift_r(w[a], w[a + 1 < 4 ? a + 1 : 3], sh) * kHashMul;
		}
		// sliding minimum over W consecutive hashes (van Herk / Gil-Werman)
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
		uint32_t pass = 0, bit = 0;
#pragma unroll
		for (int i = 0; i < 32; ++i)
			gate2_step<SH>(win[i], i ? win[i - 1] : ~win[0], s_l0_addr, base_lo, base_hi, P.four, bit, pass);
		pass = __brev(pass);                 // steps pushed position 0 first, so it ended up at bit 31
		pass &= valid;

		// level 2: the k-mer bitmap, for the ~1.5 % of positions whose minimizer is a site minimizer
		uint32_t pass2 = 0;
		const uint32_t wshift = P.filter_shift + 5;
		while (pass) {
			uint32_t i;
			asm("bfind.u32 %0, %1;" : "=r"(i) : "r"(pass));      // highest set bit (one FLO)
			pass ^= 1u << i;
			const bool up = i >= 16;
			const uint32_t x0 = up ? w[1] : w[0], x1 = up ? w[2] : w[1], x2 = up ? w[3] : w[2];
			const uint32_t lo = __funnelshift_r(x0, x1, 2 * i), hi = __funnelshift_r(x1, x2, 2 * i);
			const uint32_t mix = lo * kG2MixA + hi * (kG2MixB << (64 - 2 * K));
			const uint32_t v = __ldg(P.filter + (mix >> wshift));
			const uint32_t t = mix * kG2MixC;
			const uint32_t both = __funnelshift_r(v, v, t) & __funnelshift_r(v, v, t >> 5) & 1u;   // bits (t & 31) and (t >> 5 & 31)
			pass2 |= both << i;
		}
		if (pass2) hits += resolve_survivors<K>(P, w, pass2, K, kmask);
	}
	flush_tallies(tk, hits, P.totals);
}

}  // namespace ntsm
