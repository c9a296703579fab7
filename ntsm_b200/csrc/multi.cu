// multi.cu -- C ABI of the multi-sample matrix path: MultiCount (src/MultiCount.hpp:36-289) on the device.
// The object sits on a ctx that already holds the site table (ntsm_load_siteset = MultiCount::initCountsHash,
// :214-288, which is FingerPrint's); the matrix m_matCounts[sample][k-mer] (:208,266) lives in device memory.
// Kernels and the reasoning behind them: multi.cuh.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/ntsm_b200.h"
#include "internal.h"
#include "multi.cuh"

using namespace ntsm;

struct ntsm_multi {
	ntsm_ctx *ctx = nullptr;
	ntsm_ctx_view v{};
	cudaStream_t stream = nullptr;
	uint32_t n_samples = 0;
	uint64_t stride = 0;                    // m_kmerToHash.size() (:55): the ABI only builds printable matrices, where it equals the listed k-mers
	uint8_t *d_mat = nullptr;               // [n_samples][stride]
	// per-batch scratch, grown on demand
	char *d_windows = nullptr;  size_t windows_cap = 0;
	uint16_t *d_lens = nullptr; size_t lens_cap = 0;
	uint32_t *d_geno = nullptr; size_t geno_cap = 0;
	uint32_t *d_occ = nullptr, *d_list = nullptr, *d_uniq = nullptr; size_t occ_cap = 0;
	unsigned long long *d_cnt = nullptr, *d_off = nullptr, *d_partial = nullptr, *d_scalars = nullptr;   // [n_kmers + 1] x 2, scan partials, {grand total, warnings}
	uint32_t *d_cursor = nullptr;
	uint32_t *d_rows = nullptr;             // 4 * n_sites (printCountsMax)
	uint32_t *d_one = nullptr;              // result of a single insertCount
	double *d_values = nullptr, *d_sums = nullptr;   // printNormMatrix's numbers between ntsm_multi_norm_begin and _end
	std::vector<uint32_t> h_geno;
	// the warnings of insertCount (:59-62) in the reference's serial order: (byte found, value wanted)
	std::vector<std::pair<uint8_t, uint32_t>> warnings;
	// device time of the kernels, by CUDA events on the stream (ntsm_multi_kernel_ms)
	cudaEvent_t ev[6] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr };
	double ms_lists = 0, ms_fill = 0, ms_norm = 0;
	uint64_t cells_touched = 0;             // (occurring k-mer, sample) pairs the fill kernel walked (once with multi <= 127, else twice)
};

static int mfail(ntsm_multi *m, int code, const char *fmt, ...)
{
	char b[512];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(b, sizeof b, fmt, ap);
	va_end(ap);
	ntsm_ctx_set_error(m ? m->ctx : nullptr, b);
	return code;
}

#define MCU(m, call)                                                                                                   \
	do {                                                                                                               \
		cudaError_t e_ = (call);                                                                                       \
		if (e_ != cudaSuccess) return mfail(m, NTSM_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
	} while (0)

template <class T> static int grow(ntsm_multi *m, T **p, size_t *cap, size_t need)
{
	if (need <= *cap) return NTSM_OK;
	if (*p) MCU(m, cudaFree(*p));
	*p = nullptr;
	*cap = 0;
	const size_t n = need + need / 4 + 64;
	MCU(m, cudaMalloc(p, n * sizeof(T)));
	*cap = n;
	return NTSM_OK;
}

extern "C" void ntsm_multi_destroy(ntsm_multi *m)
{
	if (!m) return;
	cudaSetDevice(m->v.device);
	void *ptrs[] = { m->d_mat, m->d_windows, m->d_lens, m->d_geno, m->d_occ, m->d_list, m->d_uniq, m->d_cnt, m->d_off,
	                 m->d_partial, m->d_scalars, m->d_cursor, m->d_rows, m->d_one, m->d_values, m->d_sums };
	for (void *p : ptrs)
		if (p) cudaFree(p);
	for (cudaEvent_t e : m->ev)
		if (e) cudaEventDestroy(e);
	delete m;
}

extern "C" int ntsm_multi_create(ntsm_multi **out, ntsm_ctx *ctx, uint32_t n_samples)
{
	if (!out || !ctx) return mfail(nullptr, NTSM_ERR_ARG, "ntsm_multi_create: null argument");
	*out = nullptr;
	ntsm_multi *m = new ntsm_multi();
	m->ctx = ctx;
	auto init = [&]() -> int {
		const int rc = ntsm_ctx_view_get(ctx, &m->v);
		if (rc) return rc;
		m->stream = (cudaStream_t)m->v.stream;
		m->n_samples = n_samples;
		m->stride = m->v.n_kmers;
		MCU(m, cudaSetDevice(m->v.device));
		const size_t bytes = (size_t)m->stride * n_samples;
		MCU(m, cudaMalloc(&m->d_mat, bytes + 16));
		MCU(m, cudaMemsetAsync(m->d_mat, 0, bytes + 16, m->stream));                   // :266
		const size_t nk = (size_t)m->v.n_kmers + 1;
		MCU(m, cudaMalloc(&m->d_cnt, nk * 8));
		MCU(m, cudaMalloc(&m->d_off, nk * 8));
		MCU(m, cudaMalloc(&m->d_partial, ((nk + kScanThreads - 1) / kScanThreads + 1) * 8));
		MCU(m, cudaMalloc(&m->d_scalars, 2 * 8));
		MCU(m, cudaMalloc(&m->d_cursor, nk * 4));
		MCU(m, cudaMalloc(&m->d_rows, (size_t)4 * (m->v.n_sites + 1) * 4));
		MCU(m, cudaMalloc(&m->d_one, 2 * 4));
		for (cudaEvent_t &e : m->ev) MCU(m, cudaEventCreate(&e));
		MCU(m, cudaStreamSynchronize(m->stream));
		return NTSM_OK;
	};
	const int rc = init();
	if (rc) {
		ntsm_multi_destroy(m);
		return rc;
	}
	*out = m;
	return NTSM_OK;
}

extern "C" uint32_t ntsm_multi_n_samples(const ntsm_multi *m) { return m ? m->n_samples : 0; }

extern "C" int ntsm_multi_insert_count(ntsm_multi *m, uint32_t sample, uint64_t hash, uint32_t multi)
{
	if (!m) return NTSM_ERR_ARG;
	if (sample >= m->n_samples) return mfail(m, NTSM_ERR_ARG, "sample %u out of range (%u samples)", sample, m->n_samples);
	MCU(m, cudaSetDevice(m->v.device));
	multi_insert_one_kernel<<<1, 1, 0, m->stream>>>((const TableSlot *)m->v.d_table, m->v.table_mask, hash, m->d_mat, m->stride, sample,
	                                                 multi, m->d_one);
	MCU(m, cudaGetLastError());
	ntsm_ctx_add_launches(m->ctx, 1);
	uint32_t r[2] = { 0, 0 };
	MCU(m, cudaMemcpyAsync(r, m->d_one, sizeof r, cudaMemcpyDeviceToHost, m->stream));
	MCU(m, cudaStreamSynchronize(m->stream));
	if (r[0] == 3) m->warnings.emplace_back((uint8_t)r[1], multi);
	return NTSM_OK;
}

// one batch of at most kMaxLines lines: the four kernels of multi.cuh
static const uint32_t kMaxLines = 1u << 15;

static int insert_batch(ntsm_multi *m, const char *windows, uint32_t wstride, const uint16_t *lens, const uint32_t *geno2,
                        uint32_t n_lines, uint32_t multi)
{
	const uint32_t k = m->v.k, S = m->n_samples;
	uint32_t maxlen = 0;
	for (uint32_t i = 0; i < 2 * n_lines; ++i) {
		if (lens[i] > wstride) return mfail(m, NTSM_ERR_ARG, "window %u: length %u exceeds the stride %u", i, lens[i], wstride);
		maxlen = std::max<uint32_t>(maxlen, lens[i]);
	}
	if (maxlen < k || S == 0) return NTSM_OK;                     // no window holds a k-mer / nobody to insert for
	const uint32_t J = maxlen - k + 1;
	const uint64_t n_occ = (uint64_t)n_lines * 2 * J;
	if (n_occ >= 0xFFFFFFFFull) return mfail(m, NTSM_ERR_ARG, "batch too large: %llu window offsets", (unsigned long long)n_occ);
	const uint32_t gwords = (S + 15) / 16;

	int rc;
	if ((rc = grow(m, &m->d_windows, &m->windows_cap, (size_t)n_lines * 2 * wstride))) return rc;
	if ((rc = grow(m, &m->d_lens, &m->lens_cap, (size_t)n_lines * 2))) return rc;
	if ((rc = grow(m, &m->d_geno, &m->geno_cap, (size_t)n_lines * gwords))) return rc;
	if (n_occ > m->occ_cap) {
		for (uint32_t **p : { &m->d_occ, &m->d_list, &m->d_uniq }) {
			if (*p) MCU(m, cudaFree(*p));
			*p = nullptr;
		}
		m->occ_cap = 0;
		const size_t n = n_occ + n_occ / 4 + 64;
		MCU(m, cudaMalloc(&m->d_occ, n * 4));
		MCU(m, cudaMalloc(&m->d_list, n * 4));
		MCU(m, cudaMalloc(&m->d_uniq, n * 4));
		m->occ_cap = n;
	}
	cudaStream_t st = m->stream;
	MCU(m, cudaMemcpyAsync(m->d_windows, windows, (size_t)n_lines * 2 * wstride, cudaMemcpyHostToDevice, st));
	MCU(m, cudaMemcpyAsync(m->d_lens, lens, (size_t)n_lines * 2 * sizeof(uint16_t), cudaMemcpyHostToDevice, st));
	MCU(m, cudaMemcpyAsync(m->d_geno, geno2, (size_t)n_lines * gwords * 4, cudaMemcpyHostToDevice, st));
	ntsm_ctx_add_pcie(m->ctx, (uint64_t)n_lines * 2 * wstride + (uint64_t)n_lines * 4 + (uint64_t)n_lines * gwords * 4, 0);
	const uint32_t nk = m->v.n_kmers;
	MCU(m, cudaMemsetAsync(m->d_cnt, 0, ((size_t)nk + 1) * 8, st));
	MCU(m, cudaMemsetAsync(m->d_cursor, 0, ((size_t)nk + 1) * 4, st));
	MCU(m, cudaMemsetAsync(m->d_scalars, 0, 16, st));

	MCU(m, cudaEventRecord(m->ev[0], st));
	VcfBatch B{ m->d_windows, m->d_lens, wstride, n_lines, J, m->d_geno, gwords };
	const uint32_t occ_blocks = (uint32_t)((n_occ + 255) / 256);
	vcf_kmerize_kernel<<<occ_blocks, 256, 0, st>>>(B, k, (const TableSlot *)m->v.d_table, m->v.table_mask, m->d_occ, m->d_cnt);
	const uint32_t scan_blocks = (nk + kScanThreads - 1) / kScanThreads;
	occ_scan_blocks_kernel<<<std::max(scan_blocks, 1u), kScanThreads, 0, st>>>(m->d_cnt, m->d_off, m->d_partial, nk);
	occ_scan_partials_kernel<<<1, kScanThreads, 0, st>>>(m->d_partial, std::max(scan_blocks, 1u), m->d_scalars);
	occ_scan_add_kernel<<<nk / kScanThreads + 1, kScanThreads, 0, st>>>(m->d_off, m->d_partial, m->d_scalars, nk);
	occ_scatter_kernel<<<occ_blocks, 256, 0, st>>>(m->d_occ, n_occ, m->d_off, m->d_cursor, m->d_list, m->d_uniq);
	MCU(m, cudaGetLastError());
	MCU(m, cudaEventRecord(m->ev[1], st));
	unsigned long long grand = 0;
	MCU(m, cudaMemcpyAsync(&grand, m->d_scalars, 8, cudaMemcpyDeviceToHost, st));
	MCU(m, cudaStreamSynchronize(st));
	ntsm_ctx_add_launches(m->ctx, 5);
	float ms = 0;
	if (cudaEventElapsedTime(&ms, m->ev[0], m->ev[1]) == cudaSuccess) m->ms_lists += ms;
	const uint32_t n_uniq = (uint32_t)(grand >> 32);
	if (n_uniq == 0) return NTSM_OK;

	FillParams P{ m->d_uniq, n_uniq, m->d_off, m->d_list, m->d_geno, gwords, J, S, multi, m->d_mat, m->stride, m->d_scalars + 1, nullptr };
	const dim3 grid((n_uniq + 127) / 128, std::min<uint32_t>(gwords, 65535u));
	const bool one_pass = multi <= 127;                           // every value fits the byte: write and count together (multi.cuh)
	MCU(m, cudaEventRecord(m->ev[2], st));
	if (one_pass) multi_fill_kernel<3><<<grid, 128, 0, st>>>(P);
	else multi_fill_kernel<0><<<grid, 128, 0, st>>>(P);
	MCU(m, cudaGetLastError());
	MCU(m, cudaEventRecord(m->ev[3], st));
	unsigned long long n_warn = 0;
	MCU(m, cudaMemcpyAsync(&n_warn, m->d_scalars + 1, 8, cudaMemcpyDeviceToHost, st));
	MCU(m, cudaStreamSynchronize(st));
	ntsm_ctx_add_launches(m->ctx, 1);
	if (cudaEventElapsedTime(&ms, m->ev[2], m->ev[3]) == cudaSuccess) m->ms_fill += ms;
	if (n_warn) {
		MultiWarn *d_warn = nullptr;
		MCU(m, cudaMalloc(&d_warn, n_warn * sizeof(MultiWarn)));
		cudaMemsetAsync(m->d_scalars + 1, 0, 8, st);
		P.warn = d_warn;
		multi_fill_kernel<1><<<grid, 128, 0, st>>>(P);
		std::vector<MultiWarn> w(n_warn);
		cudaError_t e = cudaMemcpyAsync(w.data(), d_warn, n_warn * sizeof(MultiWarn), cudaMemcpyDeviceToHost, st);
		if (e == cudaSuccess) e = cudaStreamSynchronize(st);
		cudaFree(d_warn);
		if (e != cudaSuccess) return mfail(m, NTSM_ERR_CUDA, "recording insertCount warnings: %s", cudaGetErrorString(e));
		ntsm_ctx_add_launches(m->ctx, 1);
		// the order one thread would have raised them in: line, ref before alt, window offset, sample (VCFConvert.hpp:148-169)
		std::sort(w.begin(), w.end(), [](const MultiWarn &a, const MultiWarn &b) { return a.occ != b.occ ? a.occ < b.occ : a.sample < b.sample; });
		for (const MultiWarn &x : w) m->warnings.emplace_back((uint8_t)x.old_value, x.wanted);
	}
	if (!one_pass) {
		MCU(m, cudaEventRecord(m->ev[4], st));
		multi_fill_kernel<2><<<grid, 128, 0, st>>>(P);
		MCU(m, cudaGetLastError());
		MCU(m, cudaEventRecord(m->ev[5], st));
		MCU(m, cudaStreamSynchronize(st));
		ntsm_ctx_add_launches(m->ctx, 1);
		if (cudaEventElapsedTime(&ms, m->ev[4], m->ev[5]) == cudaSuccess) m->ms_fill += ms;
	}
	// (every path above ends on a stream synchronize: the host buffers of this batch are the caller's again)
	m->cells_touched += (uint64_t)n_uniq * S;
	return NTSM_OK;
}

// genotypes as they travel: 2 bits per sample, 16 samples per little-endian uint32, (n_samples + 15) / 16 words per line
int ntsm_multi_insert_windows_packed(ntsm_multi *m, const char *windows, uint32_t wstride, const uint16_t *lens, const uint32_t *geno2,
                                     uint32_t n_lines, uint32_t multi)
{
	if (!m || (n_lines && (!windows || !lens || (!geno2 && m->n_samples)))) return mfail(m, NTSM_ERR_ARG, "ntsm_multi_insert_windows: null argument");
	MCU(m, cudaSetDevice(m->v.device));
	const uint32_t gwords = (m->n_samples + 15) / 16;
	for (uint32_t at = 0; at < n_lines; at += kMaxLines) {
		const uint32_t n = std::min(kMaxLines, n_lines - at);
		const int rc = insert_batch(m, windows + (size_t)at * 2 * wstride, wstride, lens + (size_t)at * 2, geno2 + (size_t)at * gwords, n, multi);
		if (rc) return rc;
	}
	return NTSM_OK;
}

extern "C" int ntsm_multi_insert_windows(ntsm_multi *m, const char *windows, uint32_t wstride, const uint16_t *lens,
                                         const uint8_t *genotypes, uint32_t n_lines, uint32_t multi)
{
	if (!m || (n_lines && (!windows || !lens || (!genotypes && m->n_samples)))) return mfail(m, NTSM_ERR_ARG, "ntsm_multi_insert_windows: null argument");
	const uint32_t S = m->n_samples, gwords = (S + 15) / 16;
	m->h_geno.assign((size_t)n_lines * gwords, 0u);
	for (uint32_t l = 0; l < n_lines; ++l) {
		const uint8_t *g = genotypes + (size_t)l * S;
		uint32_t *w = m->h_geno.data() + (size_t)l * gwords;
		for (uint32_t s = 0; s < S; ++s) {
			if (g[s] > 2) return mfail(m, NTSM_ERR_ARG, "line %u sample %u: genotype code %u (0 hom1, 1 het, 2 hom2)", l, s, g[s]);
			w[s >> 4] |= (uint32_t)g[s] << (2 * (s & 15));
		}
	}
	return ntsm_multi_insert_windows_packed(m, windows, wstride, lens, m->h_geno.data(), n_lines, multi);
}

extern "C" uint64_t ntsm_multi_n_warnings(const ntsm_multi *m) { return m ? m->warnings.size() : 0; }

extern "C" int64_t ntsm_multi_warnings_text(const ntsm_multi *m, char *buf, size_t cap)
{
	if (!m) return NTSM_ERR_ARG;
	std::string o;
	for (const auto &w : m->warnings) {                           // :60-61; the byte goes out as a character
		o += "Warning: Inconsistent k-mer counts, check for overlapping sites: ";
		o.push_back((char)w.first);
		o += " vs " + std::to_string(w.second) + "\n";
	}
	if (buf) memcpy(buf, o.data(), std::min(cap, o.size()));
	return (int64_t)o.size();
}

extern "C" int ntsm_multi_get_matrix(ntsm_multi *m, uint8_t *out)
{
	if (!m || !out) return NTSM_ERR_ARG;
	MCU(m, cudaSetDevice(m->v.device));
	const size_t bytes = (size_t)m->stride * m->n_samples;
	MCU(m, cudaMemcpyAsync(out, m->d_mat, bytes, cudaMemcpyDeviceToHost, m->stream));
	MCU(m, cudaStreamSynchronize(m->stream));
	ntsm_ctx_add_pcie(m->ctx, 0, bytes);
	return NTSM_OK;
}

extern "C" int ntsm_multi_counts_max(ntsm_multi *m, uint32_t sample, uint32_t *max_ref, uint32_t *max_var, uint32_t *sum_ref,
                                     uint32_t *sum_var)
{
	if (!m) return NTSM_ERR_ARG;
	if (sample >= m->n_samples) return mfail(m, NTSM_ERR_ARG, "sample %u out of range (%u samples)", sample, m->n_samples);
	MCU(m, cudaSetDevice(m->v.device));
	const uint32_t S = m->v.n_sites;
	if (S == 0) return NTSM_OK;
	uint32_t *r = m->d_rows;
	multi_counts_max_kernel<<<(S + 255) / 256, 256, 0, m->stream>>>(m->d_mat + m->stride * sample, m->v.d_allele_off, S, r, r + S, r + 2 * (size_t)S,
	                                                                 r + 3 * (size_t)S);
	MCU(m, cudaGetLastError());
	ntsm_ctx_add_launches(m->ctx, 1);
	uint32_t *outs[4] = { max_ref, max_var, sum_ref, sum_var };
	for (int i = 0; i < 4; ++i)
		if (outs[i]) MCU(m, cudaMemcpyAsync(outs[i], r + (size_t)i * S, (size_t)S * 4, cudaMemcpyDeviceToHost, m->stream));
	MCU(m, cudaStreamSynchronize(m->stream));
	ntsm_ctx_add_pcie(m->ctx, 0, (uint64_t)S * 16);
	return NTSM_OK;
}

// printNormMatrix in three steps, so that a caller can take the rows in blocks while it formats the previous ones:
// begin runs the kernels and leaves values[site][sample] and sums[site] on the device (first_undef, nullable: the
// row-major index of the first missing value, UINT64_MAX if there is none -- where the reference's stream switches to
// 19 digits, :192), fetch copies rows [row0, row0 + n_rows) into caller memory, end frees the device copies.
int ntsm_multi_norm_begin(ntsm_multi *m, uint64_t *first_undef)
{
	if (!m) return NTSM_ERR_ARG;
	MCU(m, cudaSetDevice(m->v.device));
	ntsm_multi_norm_end(m);
	const uint32_t S = m->v.n_sites, N = m->n_samples;
	if (first_undef) *first_undef = UINT64_MAX;
	if (S == 0) return NTSM_OK;
	MCU(m, cudaMalloc(&m->d_values, ((size_t)S * N + 1) * sizeof(double)));
	MCU(m, cudaMalloc(&m->d_sums, (size_t)S * sizeof(double)));
	MCU(m, cudaMemsetAsync(m->d_scalars, 0xFF, 8, m->stream));
	MCU(m, cudaEventRecord(m->ev[0], m->stream));
	if (N) {
		const dim3 grid((S + 31) / 32, (N + 31) / 32);
		multi_norm_kernel<<<grid, dim3(32, 32), 0, m->stream>>>(m->d_mat, m->stride, m->v.d_allele_off, S, N, m->d_values, m->d_scalars);
	}
	multi_norm_sum_kernel<<<(S + 7) / 8, 256, 0, m->stream>>>(m->d_values, S, N, m->d_sums);
	MCU(m, cudaGetLastError());
	MCU(m, cudaEventRecord(m->ev[1], m->stream));
	unsigned long long first = 0;
	MCU(m, cudaMemcpyAsync(&first, m->d_scalars, 8, cudaMemcpyDeviceToHost, m->stream));
	MCU(m, cudaStreamSynchronize(m->stream));
	if (first_undef) *first_undef = first;
	float ms = 0;
	if (cudaEventElapsedTime(&ms, m->ev[0], m->ev[1]) == cudaSuccess) m->ms_norm += ms;
	ntsm_ctx_add_launches(m->ctx, N ? 2 : 1);
	return NTSM_OK;
}

int ntsm_multi_norm_fetch(ntsm_multi *m, uint32_t row0, uint32_t n_rows, double *values, double *sums)
{
	if (!m) return NTSM_ERR_ARG;
	const uint32_t S = m->v.n_sites, N = m->n_samples;
	if (n_rows == 0) return NTSM_OK;
	if (!m->d_values || row0 > S || n_rows > S - row0) return mfail(m, NTSM_ERR_ARG, "ntsm_multi_norm_fetch: rows %u..%u of %u (begin first)", row0, row0 + n_rows, S);
	MCU(m, cudaSetDevice(m->v.device));
	if (values && N) MCU(m, cudaMemcpyAsync(values, m->d_values + (size_t)row0 * N, (size_t)n_rows * N * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
	if (sums) MCU(m, cudaMemcpyAsync(sums, m->d_sums + row0, (size_t)n_rows * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
	MCU(m, cudaStreamSynchronize(m->stream));
	ntsm_ctx_add_pcie(m->ctx, 0, (values ? (uint64_t)n_rows * N * 8 : 0) + (sums ? (uint64_t)n_rows * 8 : 0));
	return NTSM_OK;
}

void ntsm_multi_norm_end(ntsm_multi *m)
{
	if (!m) return;
	if (m->d_values) cudaFree(m->d_values);
	if (m->d_sums) cudaFree(m->d_sums);
	m->d_values = m->d_sums = nullptr;
}

extern "C" int ntsm_multi_norm_matrix(ntsm_multi *m, double *values, double *sums)
{
	if (!m) return NTSM_ERR_ARG;
	int rc = ntsm_multi_norm_begin(m, nullptr);
	if (rc == NTSM_OK) rc = ntsm_multi_norm_fetch(m, 0, m->v.n_sites, values, sums);
	ntsm_multi_norm_end(m);
	return rc;
}

// device time so far, by CUDA events on the stream: ms[0] the k-merize + table lookup + occurrence lists of every batch,
// ms[1] the fill kernel's passes, ms[2] the norm-matrix kernels; cells = (occurring k-mer, sample) pairs per fill pass
extern "C" void ntsm_multi_kernel_ms(const ntsm_multi *m, double ms[3], uint64_t *cells)
{
	if (!m) return;
	if (ms) { ms[0] = m->ms_lists; ms[1] = m->ms_fill; ms[2] = m->ms_norm; }
	if (cells) *cells = m->cells_touched;
}
