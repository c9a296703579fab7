// multi.cuh -- device side of the multi-sample matrix path (sm_100a): MultiCount driven by VCFConvert
// (src/MultiCount.hpp:52-70,93-203; src/VCFConvert.hpp:148-169).  SURVEY 8f rank 4.
//
// What the reference does: for every SNP line of a multi-sample VCF it cuts the window around the site out of
// the reference genome twice (ref allele, alt allele), rolls both through KseqHashIterator and, for every
// k-mer and every sample, calls MultiCount::insertCount(sample, hash, multi or 2 * multi) -- a robin_map find
// plus a compare-and-swap on one byte of m_matCounts[sample][k-mer] that only the FIRST writer of a cell wins;
// a later insert of a different value prints a warning.  printNormMatrix then walks sites x samples x k-mers
// through the same map (two `at` per k-mer) to take maxREF / (maxREF + maxVAR).  Everything is lookups into a
// 100 MB hash map from one thread per VCF line.
//
// Here the matrix lives in device memory as uint8 mat[sample][dense k-mer index] (the reference's layout and
// row stride), and a batch of VCF lines is processed by four small kernels:
//   1. vcf_kmerize_kernel   one thread per (line, allele, window offset): decode (nt4 table), validity,
//                           canonical value, the reference's hash64, probe of the ctx's open-addressing
//                           table -> dense index of that occurrence (or none), and a per-k-mer occurrence count;
//   2. an exclusive scan over the k-mer index space (occurrence offsets + rank among the k-mers that occur);
//   3. occ_scatter_kernel   occurrence lists per k-mer (CSR) and the compact list of k-mers that occur;
//   4. multi_fill_kernel    one thread per (occurring k-mer, 16 samples) replays each cell's inserts IN THE
//                           REFERENCE'S SERIAL ORDER (line, ref before alt, window offset) against the cell's
//                           current byte: first non-zero writer wins, later different values are warnings.
//                           Cells are independent, so the result is the one-thread reference's bit for bit
//                           -- including which inserts warn -- however many threads the GPU runs.
// and the two printers are one kernel each over (site, sample) (multi_norm_kernel, multi_counts_max_kernel)
// plus a warp-per-site sum that adds the samples' values in sample order, as the reference's loop does, so
// that the double it hands to the long-double division is the same double.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "kmer_math.h"
#include "table.cuh"

namespace ntsm {

constexpr uint32_t kNoIdx = 0xFFFFFFFFu;
constexpr double kUndef = 1.7976931348623157e308;       // MultiCount::UNDEF = numeric_limits<double>::max() (src/MultiCount.hpp:41)

// dense index of a hashed k-mer, or kNoIdx: m_kmerToHash.find (src/MultiCount.hpp:53)
__device__ __forceinline__ uint32_t table_find(const TableSlot *__restrict__ table, uint32_t table_mask, uint64_t h)
{
	uint32_t slot = (uint32_t)(h ^ (h >> 29)) & table_mask;
	for (;;) {
		const TableSlot e = table[slot];
		if (e.key == h) return e.idx;
		if (e.key == kEmptyKey) return kNoIdx;
		slot = (slot + 1) & table_mask;
	}
}

struct VcfBatch {
	const char *windows;        // [n_lines][2][wstride] ASCII: ref-allele window, alt-allele window
	const uint16_t *lens;       // [n_lines][2] bytes of each window that count (the std::string's size, src/VCFConvert.hpp:148,159)
	uint32_t wstride;           // bytes between windows
	uint32_t n_lines;
	uint32_t J;                 // window offsets per allele = max window length - k + 1
	const uint32_t *geno;       // [n_lines][gwords] 2 bits per sample: 0 hom1, 1 het, 2 hom2
	uint32_t gwords;
};

// occurrence o = (line * 2 + allele) * J + j   <->  the k-mer at window offset j: the order of the reference's loops
__global__ void vcf_kmerize_kernel(const VcfBatch B, uint32_t k, const TableSlot *__restrict__ table, uint32_t table_mask,
                                   uint32_t *__restrict__ occ_idx, unsigned long long *__restrict__ cnt)
{
	const uint64_t o = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const uint64_t n_occ = (uint64_t)B.n_lines * 2 * B.J;
	if (o >= n_occ) return;
	const uint32_t j = (uint32_t)(o % B.J), w = (uint32_t)(o / B.J);
	uint32_t idx = kNoIdx;
	if (j + k <= B.lens[w]) {
		const unsigned char *p = reinterpret_cast<const unsigned char *>(B.windows) + (size_t)w * B.wstride + j;
		uint64_t fw = 0, rv = 0;
		bool ok = true;
		for (uint32_t t = 0; t < k; ++t) {
			const uint64_t c = nt4(p[t]);                         // vendor/KseqHashIterator.hpp:96-97
			ok = ok && c < 4;
			fw = (fw << 2) | (c & 3);                             // :99
			rv = (rv >> 2) | ((3 - (c & 3)) << (2 * (k - 1)));    // :100
		}
		if (ok) idx = table_find(table, table_mask, hash64(fw < rv ? fw : rv, kmer_mask(k)));   // :102, MultiCount.hpp:53
	}
	occ_idx[o] = idx;
	if (idx != kNoIdx) atomicAdd(cnt + idx, 1ull);
}

// ---- exclusive scan of packed (occurrences | distinct << 32) over the k-mer index space ----
// in[i] = occurrences of k-mer i; out[i] = (occurrences before i) | (occurring k-mers before i) << 32.
constexpr int kScanThreads = 1024;

__device__ __forceinline__ unsigned long long block_exclusive_scan(unsigned long long v, unsigned long long *total)
{
	__shared__ unsigned long long warp_sum[kScanThreads / 32];
	const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	unsigned long long incl = v;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, d);
		if (lane >= (uint32_t)d) incl += t;
	}
	if (lane == 31) warp_sum[wid] = incl;
	__syncthreads();
	if (wid == 0) {
		unsigned long long s = warp_sum[lane];
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const unsigned long long t = __shfl_up_sync(0xffffffffu, s, d);
			if (lane >= (uint32_t)d) s += t;
		}
		warp_sum[lane] = s;                                       // inclusive over the warps
	}
	__syncthreads();
	const unsigned long long before = wid ? warp_sum[wid - 1] : 0ull;
	if (total) *total = warp_sum[kScanThreads / 32 - 1];
	__syncthreads();
	return before + incl - v;
}

__global__ void __launch_bounds__(kScanThreads) occ_scan_blocks_kernel(const unsigned long long *__restrict__ cnt, unsigned long long *__restrict__ off,
                                                                       unsigned long long *__restrict__ partial, uint32_t n)
{
	const uint32_t i = blockIdx.x * kScanThreads + threadIdx.x;
	const unsigned long long c = i < n ? cnt[i] : 0ull;
	unsigned long long total;
	const unsigned long long ex = block_exclusive_scan(c ? (c | (1ull << 32)) : 0ull, &total);
	if (i < n) off[i] = ex;
	if (threadIdx.x == 0) partial[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kScanThreads) occ_scan_partials_kernel(unsigned long long *partial, uint32_t n_blocks, unsigned long long *grand_total)
{
	unsigned long long carry = 0;
	for (uint32_t base = 0; base < n_blocks; base += kScanThreads) {
		const uint32_t i = base + threadIdx.x;
		const unsigned long long v = i < n_blocks ? partial[i] : 0ull;
		unsigned long long total;
		const unsigned long long ex = block_exclusive_scan(v, &total);
		if (i < n_blocks) partial[i] = carry + ex;
		carry += total;
	}
	if (threadIdx.x == 0) *grand_total = carry;                   // (occurrences with a table hit) | (k-mers that occur) << 32
}

__global__ void __launch_bounds__(kScanThreads) occ_scan_add_kernel(unsigned long long *__restrict__ off, const unsigned long long *__restrict__ partial,
                                                                    const unsigned long long *__restrict__ grand_total, uint32_t n)
{
	const uint32_t i = blockIdx.x * kScanThreads + threadIdx.x;
	if (i < n) off[i] += partial[blockIdx.x];
	if (i == n) off[n] = *grand_total;                            // off has n + 1 entries: a k-mer's count is off[i + 1] - off[i]
}

__global__ void occ_scatter_kernel(const uint32_t *__restrict__ occ_idx, uint64_t n_occ, const unsigned long long *__restrict__ off,
                                   uint32_t *__restrict__ cursor, uint32_t *__restrict__ list, uint32_t *__restrict__ uniq)
{
	const uint64_t o = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (o >= n_occ) return;
	const uint32_t idx = occ_idx[o];
	if (idx == kNoIdx) return;
	const unsigned long long e = off[idx];
	const uint32_t at = atomicAdd(cursor + idx, 1u);
	list[(uint32_t)e + at] = (uint32_t)o;                         // unordered inside a k-mer's list; the fill kernel walks it in ascending o
	if (at == 0) uniq[(uint32_t)(e >> 32)] = idx;
}

// One warning of MultiCount::insertCount (:59-62): which insert (occurrence, sample) found which byte and wanted what
struct MultiWarn {
	uint32_t occ, sample;
	uint32_t old_value, wanted;
};

struct FillParams {
	const uint32_t *uniq;           // [n_uniq] k-mers that occur in this batch
	uint32_t n_uniq;
	const unsigned long long *off;  // [n_kmers + 1] low word: start of the k-mer's list
	const uint32_t *list;
	const uint32_t *geno;
	uint32_t gwords, J;
	uint32_t n_samples, multi;
	uint8_t *mat;
	uint64_t stride;                // m_kmerToHash.size(): bytes between samples (:55)
	unsigned long long *n_warn;     // MODE 0: number of warnings; MODE 1: cursor into warn
	MultiWarn *warn;
};

// MODE 0: dry run, only count the warnings this batch will raise; MODE 1: dry run, record them; MODE 2: write the cells;
// MODE 3: write the cells and count the warnings in one pass.
// With multi <= 127 every inserted value fits the byte, and a replay that starts from the bytes AFTER the batch raises
// exactly the warnings of the replay from the bytes before it (a cell that was empty now holds its first writer's value,
// which that writer's own insert equals): one pass of MODE 3, and MODE 1 afterwards only if it counted any.  With larger
// multi the byte wraps (400 is stored as 144 and a later 400 "differs"), so 0, (1), then 2: the cells are written last.
// MultiCount::insertCount's decision for one insert of value v into a cell that holds x (:58-67)
template <int MODE>
__device__ __forceinline__ void fill_insert(const FillParams &P, uint32_t &x, uint32_t v, uint32_t o, uint32_t s)
{
	if (x > 0) {                                                  // :58
		if (x != v) {                                             // :59
			if (MODE == 0 || MODE == 3) atomicAdd(P.n_warn, 1ull);
			if (MODE == 1) {
				const unsigned long long at = atomicAdd(P.n_warn, 1ull);
				P.warn[at] = MultiWarn{ o, s, x, v };
			}
		}
	} else x = v & 0xFFu;                                         // :65-67, the byte takes the value's low 8 bits
}

template <int MODE>
__global__ void multi_fill_kernel(const FillParams P)
{
	const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
	if (u >= P.n_uniq) return;
	const uint32_t idx = P.uniq[u];
	const uint32_t beg = (uint32_t)P.off[idx], n = (uint32_t)P.off[idx + 1] - beg;
	const uint32_t first = P.list[beg];
	// a thread takes its k-mer through 16 samples at a time (one word of genotypes): the list offsets are read once
	// per 16 cells, and a warp's 32 k-mers are 32 neighbouring bytes of each sample's row
	for (uint32_t sg = blockIdx.y; sg < P.gwords; sg += gridDim.y) {
		const uint32_t s_end = min(P.n_samples, sg * 16 + 16);
		if (n == 1) {
			// the k-mer occurs once in the batch (nearly always): one genotype word decides all 16 cells
			const uint32_t w = first / P.J;                       // line * 2 + allele
			const uint32_t gw = P.geno[(size_t)(w >> 1) * P.gwords + sg];
			const uint32_t hom = (w & 1) ? 2u : 0u;
			uint8_t *cell = P.mat + P.stride * (sg * 16) + idx;   // :56-57
			// all the bytes this thread needs are requested before the first one is looked at (ncu on the load-per-cell-in-turn
			// form: 26 of 30 stall cycles on the long scoreboard, profiles/r02z_matrix_ncu_summary.txt; this form: 6.3 -> 6.1 ms
			// for the whole job -- what is left is the index chain in front of the cells and 32-byte partial-sector writes)
			uint32_t xs[16];
#pragma unroll
			for (uint32_t t = 0; t < 16; ++t) {
				const uint32_t g = (gw >> (2 * t)) & 3u;
				xs[t] = sg * 16 + t < s_end && (g == 1 || g == hom) ? cell[(size_t)t * P.stride] : 0u;
			}
#pragma unroll
			for (uint32_t t = 0; t < 16; ++t) {
				const uint32_t g = (gw >> (2 * t)) & 3u, s = sg * 16 + t;
				if (s >= s_end || (g != 1 && g != hom)) continue; // out of range / this sample does not carry this window's allele
				uint32_t x = xs[t];
				fill_insert<MODE>(P, x, g == 1 ? P.multi : P.multi * 2, first, s);   // VCFConvert.hpp:151-155,162-166
				if ((MODE == 2 || MODE == 3) && x != xs[t]) cell[(size_t)t * P.stride] = (uint8_t)x;
			}
			continue;
		}
		for (uint32_t s = sg * 16; s < s_end; ++s) {
			uint8_t *cell = P.mat + P.stride * s + idx;
			const uint32_t x0 = *cell;
			uint32_t x = x0;
			int64_t last = -1;
			for (uint32_t t = 0; t < n; ++t) {
				uint32_t o = 0xFFFFFFFFu;                         // next occurrence in the reference's order
				for (uint32_t q = 0; q < n; ++q) {
					const uint32_t c = P.list[beg + q];
					if ((int64_t)c > last && c < o) o = c;
				}
				last = o;
				const uint32_t w = o / P.J;
				const uint32_t g = (P.geno[(size_t)(w >> 1) * P.gwords + sg] >> (2 * (s & 15))) & 3u;
				if (g == 1) fill_insert<MODE>(P, x, P.multi, o, s);                       // het: both alleles once
				else if (g == ((w & 1) ? 2u : 0u)) fill_insert<MODE>(P, x, P.multi * 2, o, s);   // homozygous for this window's allele
			}
			if ((MODE == 2 || MODE == 3) && x != x0) *cell = (uint8_t)x;
		}
	}
}

// MultiCount::insertCount for ONE (sample, hash) pair (:52-70): the single-call form of the ABI
__global__ void multi_insert_one_kernel(const TableSlot *table, uint32_t table_mask, uint64_t hash, uint8_t *mat, uint64_t stride,
                                        uint32_t sample, uint32_t multi, uint32_t *result /* 0 not in table, 1 written, 2 kept, 3 kept + warning; [1] = old byte */)
{
	const uint32_t idx = table_find(table, table_mask, hash);
	if (idx == kNoIdx) { result[0] = 0; return; }
	uint8_t *cell = mat + stride * sample + idx;
	const uint32_t x = *cell;
	result[1] = x;
	if (x > 0) result[0] = x != multi ? 3 : 2;
	else { *cell = (uint8_t)multi; result[0] = 1; }
}

// MultiCount::printCountsMax for one sample (:93-138): site_reduce_kernel over a row of bytes
__global__ void multi_counts_max_kernel(const uint8_t *__restrict__ row, const uint32_t *__restrict__ allele_off, uint32_t n_sites,
                                        uint32_t *__restrict__ max_ref, uint32_t *__restrict__ max_var, uint32_t *__restrict__ sum_ref,
                                        uint32_t *__restrict__ sum_var)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_sites) return;
	const uint32_t r0 = allele_off[2 * i], r1 = allele_off[2 * i + 1], v1 = allele_off[2 * i + 2];
	uint32_t mr = 0, sr = 0, mv = 0, sv = 0;
	for (uint32_t j = r0; j < v1; ++j) {
		const uint32_t c = row[j];
		if (j < r1) { mr = max(mr, c); sr += c; }
		else { mv = max(mv, c); sv += c; }
	}
	max_ref[i] = mr; max_var[i] = mv; sum_ref[i] = sr; sum_var[i] = sv;
}

// MultiCount::printNormMatrix, the per-(site, sample) part (:163-184): values[site][sample] = maxREF / (maxREF + maxVAR)
// as doubles, kUndef where both are zero.  A 32 x 32 tile: threads read the matrix with the SITE index fastest (adjacent
// sites are adjacent bytes of a sample's row) and write the values with the SAMPLE index fastest (rows of the output).
__global__ void __launch_bounds__(1024) multi_norm_kernel(const uint8_t *__restrict__ mat, uint64_t stride, const uint32_t *__restrict__ allele_off,
                                                           uint32_t n_sites, uint32_t n_samples, double *__restrict__ values,
                                                           unsigned long long *__restrict__ first_undef)
{
	__shared__ double tile[32][33];
	const uint32_t i = blockIdx.x * 32 + threadIdx.x, s = blockIdx.y * 32 + threadIdx.y;
	double v = kUndef;
	if (i < n_sites && s < n_samples) {
		const uint32_t r0 = allele_off[2 * i], r1 = allele_off[2 * i + 1], v1 = allele_off[2 * i + 2];
		const uint8_t *row = mat + stride * s;
		uint32_t mr = 0, mv = 0;
		for (uint32_t j = r0; j < r1; ++j) mr = max(mr, (uint32_t)row[j]);
		for (uint32_t j = r1; j < v1; ++j) mv = max(mv, (uint32_t)row[j]);
		const uint32_t denom = mr + mv;                           // :179
		if (denom) v = (double)mr / (double)denom;                // :183 (IEEE division, as the host's)
	}
	tile[threadIdx.y][threadIdx.x] = v;
	// row-major position of the first missing value of the whole matrix: where the reference's output stream starts printing 19
	// digits (MultiCount.hpp:192).  One atomic per warp that holds one.
	unsigned long long at = i < n_sites && s < n_samples && v == kUndef ? (unsigned long long)i * n_samples + s : ~0ull;
#pragma unroll
	for (int d = 16; d; d >>= 1) {
		const unsigned long long o = __shfl_xor_sync(0xffffffffu, at, d);
		at = o < at ? o : at;
	}
	if ((threadIdx.x & 31) == 0 && at != ~0ull) atomicMin(first_undef, at);
	__syncthreads();
	const uint32_t oi = blockIdx.x * 32 + threadIdx.y, os = blockIdx.y * 32 + threadIdx.x;
	if (oi < n_sites && os < n_samples) values[(size_t)oi * n_samples + os] = tile[threadIdx.x][threadIdx.y];
}

// `sum += values.at(j)` over the samples that have a value, in sample order (:184): one warp per site loads 32
// values at a time and every lane adds them in order -- the double handed to the long-double division (:188-189)
// is the reference's, not a tree sum's.
__global__ void multi_norm_sum_kernel(const double *__restrict__ values, uint32_t n_sites, uint32_t n_samples, double *__restrict__ sums)
{
	const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	if (i >= n_sites) return;
	const double *row = values + (size_t)i * n_samples;
	double sum = 0.0;
	for (uint32_t base = 0; base < n_samples; base += 32) {
		const double v = base + lane < n_samples ? row[base + lane] : kUndef;
#pragma unroll 4
		for (int t = 0; t < 32; ++t) {
			const double vt = __shfl_sync(0xffffffffu, v, t);
			if (vt != kUndef) sum += vt;
		}
	}
	if (lane == 0) sums[i] = sum;
}

}  // namespace ntsm
