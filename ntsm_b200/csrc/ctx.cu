// ctx.cu -- the C ABI over the device path: context, site table, pinned batch ring, streams,
// NCCL combine, results.  Mirrors the FingerPrint object (src/FingerPrint.hpp:32-566) and the
// bulk-buffer recycling of vendor/ProdConKseqRunner.hpp:34-46.
#include <cuda_runtime.h>
#include <nccl.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>

#include <algorithm>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/ntsm_b200.h"
#include "internal.h"
#include "kernels.cuh"
#include "gate2.cuh"
#include "seed.cuh"
#include "pair.cuh"
#include "pack.h"

using namespace ntsm;

static thread_local std::string t_last_error;

struct ntsm_batch {
	ntsm_ctx *ctx = nullptr;
	uint64_t *h_bases = nullptr;      // pinned
	uint32_t *h_mask = nullptr;       // pinned
	uint64_t *h_snap = nullptr;       // pinned {TK, hits} snapshot taken after this batch's kernel
	uint2 *d_bases = nullptr;
	uint32_t *d_mask = nullptr;
	cudaEvent_t copied = nullptr, done = nullptr;
	Packer pk;
	uint64_t cap_pos = 0;             // usable positions
	uint64_t n_bases = 0, n_reads = 0;
	int state = 0;                    // 0 free, 1 acquired, 2 in flight
	// -m only: where every read of this batch ends (stream position of its separator) and the bases
	// of the batch up to and including it -- what an exact stop after read i needs (ntsm_trim_to_cap)
	std::vector<uint32_t> read_end;
	std::vector<uint64_t> read_bases;
	uint64_t n_pos = 0;               // data positions as submitted
};

struct ntsm_ctx {
	ntsm_cfg cfg{};
	int device = 0;
	cudaStream_t copy_stream = nullptr, compute_stream = nullptr, own_compute = nullptr;
	int sm_count = 148;
	// site table
	uint32_t n_kmers = 0, n_sites = 0;
	uint32_t *d_filter = nullptr;
	uint32_t *d_level1 = nullptr;           // 4^M-bit minimizer bitmap of the k = 19 kernels (layout per variant)
	uint32_t *d_level0 = nullptr;           // image of the gated kernels' shared-memory level-0 bitmap
	int kernel_variant = 5;                 // k = 19: 0 plain, 1 minimizer, 2 smem-gated minimizer, 3 gate2, 4 strided seeds, 5 paired seeds (default)
	int pair_fold = kPairFoldDefault;       // paired-seed table folded 2^pair_fold : 1 (NTSM_PAIR_FOLD)
	int seed_cfg = 1;                       // seed kernel launch shape (NTSM_SEED_CFG): 0 = 1024x1, 1 = 1024x2 (default), 2 = 512x4, 3 = 256x8
	int gate_m = 14;                        // M-mer length of gate2 (13 or 14)
	int gate_threads = 1024;                // gate2 CTA size (NTSM_GATE_THREADS: 512 / 768 / 1024)
	bool pool_tail = true;                  // gate2: warp-pooled level-2/exact tail (NTSM_TAIL_POOL=0 -> per-lane loop)
	uint32_t filter_bits = 0;
	TableSlot *d_table = nullptr;
	uint32_t table_cap = 0;
	uint32_t *d_counts = nullptr;
	uint32_t *d_allele_off = nullptr;
	uint32_t *d_rows = nullptr;             // 4 * n_sites
	unsigned long long *d_totals = nullptr; // {TK, hits, bases}
	// batches
	std::vector<ntsm_batch *> batches;
	std::deque<ntsm_batch *> inflight;
	std::mutex mu;
	std::condition_variable cv;
	ntsm_batch *current = nullptr;          // ntsm_insert_count's open batch
	// tallies
	uint64_t done_kmers = 0, done_hits = 0, done_bases = 0;   // over completed batches
	uint64_t submitted_bases = 0;
	uint64_t launches = 0;
	// exact -m stop: the last batch submitted, a scratch tally, and what launch_count adds per hit
	ntsm_batch *last_batch = nullptr;
	unsigned long long *d_scratch_totals = nullptr;
	uint32_t launch_delta = 1;
	unsigned long long *launch_totals = nullptr;   // nullptr = d_totals
	bool reduced = false;
	ncclComm_t comm = nullptr;
	int rank = 0, n_ranks = 1;
	std::string err;
};

static int fail(ntsm_ctx *c, int code, const char *fmt, ...)
{
	char b[512];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(b, sizeof b, fmt, ap);
	va_end(ap);
	t_last_error = b;
	if (c) c->err = b;
	return code;
}

#define CU(c, call)                                                                                      \
	do {                                                                                                 \
		cudaError_t e_ = (call);                                                                         \
		if (e_ != cudaSuccess) return fail(c, NTSM_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
	} while (0)
#define NC(c, call)                                                                                      \
	do {                                                                                                 \
		ncclResult_t e_ = (call);                                                                        \
		if (e_ != ncclSuccess) return fail(c, NTSM_ERR_NCCL, "%s: %s (%s:%d)", #call, ncclGetErrorString(e_), __FILE__, __LINE__); \
	} while (0)

extern "C" const char *ntsm_version(void) { return "ntsm_b200 0.1 (reference ntsm 663f9a5 v1.2.1 counting path, sm_100a)"; }

extern "C" const char *ntsm_last_error(const ntsm_ctx *ctx) { return ctx ? ctx->err.c_str() : t_last_error.c_str(); }

extern "C" int ntsm_device_count(void)
{
	int n = 0;
	return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0;
}

extern "C" int ntsm_device_warmup(int device)
{
	// creating the primary context is most of a short run's wall time; a caller can start it on a
	// thread of its own while it still reads the site file (the CLI does)
	if (cudaSetDevice(device) != cudaSuccess) return NTSM_ERR_CUDA;
	return cudaFree(nullptr) == cudaSuccess ? NTSM_OK : NTSM_ERR_CUDA;
}

// ------------------------------------------------------------------ context
extern "C" int ntsm_ctx_create(ntsm_ctx **out, const ntsm_cfg *cfg)
{
	if (!out || !cfg) return fail(nullptr, NTSM_ERR_ARG, "ntsm_ctx_create: null argument");
	if (cfg->k < 1 || cfg->k > 31) return fail(nullptr, NTSM_ERR_ARG, "k must be in 1..31 (got %u)", cfg->k);
	int n_dev = 0;
	cudaError_t e = cudaGetDeviceCount(&n_dev);
	if (e != cudaSuccess || n_dev == 0)
		return fail(nullptr, NTSM_ERR_CUDA, "no CUDA device (%s); this library has no CPU path", cudaGetErrorString(e));
	if (cfg->device < 0 || cfg->device >= n_dev) return fail(nullptr, NTSM_ERR_ARG, "device %d out of range (%d devices)", cfg->device, n_dev);
	ntsm_ctx *c = new ntsm_ctx();
	c->cfg = *cfg;
	if (c->cfg.n_buffers == 0) c->cfg.n_buffers = 3;
	if (c->cfg.n_buffers < 2) c->cfg.n_buffers = 2;
	if (c->cfg.batch_bases == 0) c->cfg.batch_bases = 1ull << 25;
	if (c->cfg.batch_bases < 4096) c->cfg.batch_bases = 4096;
	c->device = cfg->device;
	CU(c, cudaSetDevice(c->device));
	CU(c, cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, c->device));
	CU(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
	CU(c, cudaStreamCreateWithFlags(&c->own_compute, cudaStreamNonBlocking));
	c->compute_stream = c->own_compute;
	CU(c, cudaMalloc(&c->d_totals, 3 * sizeof(unsigned long long)));
	CU(c, cudaMemset(c->d_totals, 0, 3 * sizeof(unsigned long long)));
	*out = c;
	return NTSM_OK;
}

static void free_batch(ntsm_batch *b)
{
	if (!b) return;
	cudaFreeHost(b->h_bases);
	cudaFreeHost(b->h_mask);
	cudaFreeHost(b->h_snap);
	cudaFree(b->d_bases);
	cudaFree(b->d_mask);
	if (b->copied) cudaEventDestroy(b->copied);
	if (b->done) cudaEventDestroy(b->done);
	delete b;
}

extern "C" void ntsm_ctx_destroy(ntsm_ctx *c)
{
	if (!c) return;
	cudaSetDevice(c->device);
	cudaDeviceSynchronize();
	if (c->comm) ncclCommDestroy(c->comm);
	for (ntsm_batch *b : c->batches) free_batch(b);
	cudaFree(c->d_filter);
	cudaFree(c->d_level1);
	cudaFree(c->d_level0);
	cudaFree(c->d_table);
	cudaFree(c->d_counts);
	cudaFree(c->d_allele_off);
	cudaFree(c->d_rows);
	cudaFree(c->d_totals);
	if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
	if (c->own_compute) cudaStreamDestroy(c->own_compute);
	delete c;
}

// ------------------------------------------------------------------ site table
extern "C" int ntsm_load_sites(ntsm_ctx *c, const uint64_t *kmer_hash, const uint8_t *erased, uint32_t n_kmers,
                               const uint32_t *allele_off, uint32_t n_sites)
{
	if (!c || (!kmer_hash && n_kmers) || !allele_off) return fail(c, NTSM_ERR_ARG, "ntsm_load_sites: null argument");
	CU(c, cudaSetDevice(c->device));
	const uint32_t k = c->cfg.k;
	uint64_t live = 0;
	for (uint32_t i = 0; i < n_kmers; ++i) live += !(erased && erased[i]);

	// exact table: capacity = power of two >= 2 * keys (load <= 0.5, like robin_map's default, robin_map.h:90)
	uint64_t cap = 1024;
	while (cap < 2 * live) cap <<= 1;
	if (cap > (1ull << 31)) return fail(c, NTSM_ERR_ARG, "too many site k-mers (%llu)", (unsigned long long)live);

	// which count kernel will run decides which pre-filter structures are built (NTSM_KERNEL / NTSM_GATE_M
	// are measurement knobs: every variant gives the same counts)
	int variant = k == 19 ? 5 : 0, gm = 14;
	if (const char *e = getenv("NTSM_KERNEL")) variant = atoi(e);
	if (const char *e = getenv("NTSM_GATE_M")) gm = atoi(e);
	if (k != 19 || variant < 0 || variant > 5) variant = 0;
	if (gm < 13 || gm > 14 || variant >= 4) gm = 14;
	if (const char *e = getenv("NTSM_SEED_CFG")) c->seed_cfg = std::min(3, std::max(0, atoi(e)));
	// panels far larger than the human one (cfg 5: 26 M k-mers) saturate a folded pair table and nothing
	// stays in L2 anyway: measured 318 (unfolded) vs 241 Gbases/s (profiles/r01v12_sweep_cfg5.jsonl)
	c->pair_fold = live > 5000000 ? 0 : kPairFoldDefault;
	if (const char *e = getenv("NTSM_PAIR_FOLD")) c->pair_fold = std::min(kPairFoldMax, std::max(0, atoi(e)));
	if (const char *e = getenv("NTSM_TAIL_POOL")) c->pool_tail = atoi(e) != 0;
	if (const char *e = getenv("NTSM_GATE_THREADS")) c->gate_threads = atoi(e);

	// k-mer bitmap holding both orientations of every live k-mer
	uint32_t fbits = 16;
	while (fbits < 30 && (1ull << fbits) < 40ull * 2ull * live) ++fbits;
	if (const char *e = getenv("NTSM_FILTER_BITS")) fbits = (uint32_t)std::min(32, std::max(10, atoi(e)));
	const size_t filter_words = (1ull << fbits) / 32;                 // layout depends on the variant
	// minimizer bitmaps: level 1 = 4^M bits in global memory, level 0 = image of the shared-memory bitmap
	const int mm_len = variant == 1 ? kMinimizerM : variant == 2 ? kGateM : gm;
	const size_t level1_words = variant == 5 ? kPairWords >> c->pair_fold : variant ? (1ull << (2 * mm_len)) / 32 : 0;
	const size_t level0_words = variant == 2 || variant == 3 ? kL0Words : 0;

	cudaFree(c->d_filter); cudaFree(c->d_level1); cudaFree(c->d_level0); cudaFree(c->d_table); cudaFree(c->d_counts); cudaFree(c->d_allele_off); cudaFree(c->d_rows);
	c->d_level1 = nullptr; c->d_level0 = nullptr; c->d_filter = nullptr; c->d_table = nullptr; c->d_counts = nullptr; c->d_allele_off = nullptr; c->d_rows = nullptr;
	CU(c, cudaMalloc(&c->d_filter, filter_words * 4));
	CU(c, cudaMalloc(&c->d_table, cap * sizeof(TableSlot)));
	CU(c, cudaMalloc(&c->d_counts, std::max<size_t>(1, n_kmers) * 4));
	CU(c, cudaMalloc(&c->d_allele_off, (2 * (size_t)n_sites + 1) * 4));
	CU(c, cudaMalloc(&c->d_rows, std::max<size_t>(1, n_sites) * 16));
	if (level1_words) {
		// gate2 forms probe addresses as {lo32(base) + offset, hi32(base)}: the bitmap must not cross a
		// 4 GiB line.  cudaMalloc hands out 2 MiB-aligned blocks, so a second try always fits.
		std::vector<void *> rejected;
		for (int attempt = 0; attempt < 8; ++attempt) {
			CU(c, cudaMalloc(&c->d_level1, level1_words * 4));
			const uintptr_t a = (uintptr_t)c->d_level1, z = a + level1_words * 4 - 1;
			if ((a >> 32) == (z >> 32)) break;
			rejected.push_back(c->d_level1);
			c->d_level1 = nullptr;
		}
		for (void *r : rejected) cudaFree(r);
		if (!c->d_level1) return fail(c, NTSM_ERR_CUDA, "could not place the minimizer bitmap inside one 4 GiB window");
	}
	if (level0_words) {
		CU(c, cudaMalloc(&c->d_level0, level0_words * 4));
		CU(c, cudaFuncSetAttribute(count_kernel_gate<19>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kL0Words * 4)));
		CU(c, cudaFuncSetAttribute(count_kernel_gate2<19, 13, true, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kL0Words * 4)));
		CU(c, cudaFuncSetAttribute(count_kernel_gate2<19, 14, true, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kL0Words * 4)));
		CU(c, cudaFuncSetAttribute(count_kernel_gate2<19, 14, false, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kL0Words * 4)));
		CU(c, cudaFuncSetAttribute(count_kernel_gate2<19, 14, true, 768>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kL0Words * 4)));
		CU(c, cudaFuncSetAttribute(count_kernel_gate2<19, 14, true, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kL0Words * 4)));
	}

	// Build on the device (initCountsHash's insert loop, src/FingerPrint.hpp:506-552, with the host
	// having applied the first-wins / dupes rules): the hashes go up once, one thread per k-mer
	// claims a table slot with a 64-bit CAS and sets its bits in the bitmaps with atomicOr.
	uint64_t *d_hash = nullptr;
	uint8_t *d_erased = nullptr;
	int *d_err = nullptr;
	int h_err[2] = { 0, 0 };
	cudaStream_t st = c->copy_stream;
	auto cleanup = [&]() { cudaFree(d_hash); cudaFree(d_erased); cudaFree(d_err); };
	auto build = [&]() -> int {
		CU(c, cudaMalloc(&d_hash, std::max<size_t>(1, n_kmers) * 8));
		CU(c, cudaMalloc(&d_erased, std::max<size_t>(1, n_kmers)));
		CU(c, cudaMalloc(&d_err, sizeof h_err));
		CU(c, cudaMemcpyAsync(d_hash, kmer_hash, (size_t)n_kmers * 8, cudaMemcpyHostToDevice, st));
		if (erased) CU(c, cudaMemcpyAsync(d_erased, erased, n_kmers, cudaMemcpyHostToDevice, st));
		else CU(c, cudaMemsetAsync(d_erased, 0, std::max<size_t>(1, n_kmers), st));
		CU(c, cudaMemsetAsync(d_err, 0, sizeof h_err, st));
		CU(c, cudaMemsetAsync(c->d_table, 0xFF, cap * sizeof(TableSlot), st));          // key = kEmptyKey everywhere
		CU(c, cudaMemsetAsync(c->d_filter, 0, filter_words * 4, st));
		if (level1_words) CU(c, cudaMemsetAsync(c->d_level1, 0, level1_words * 4, st));
		if (level0_words) CU(c, cudaMemsetAsync(c->d_level0, 0, level0_words * 4, st));
		BuildParams B;
		B.hash = d_hash; B.erased = d_erased; B.n_kmers = n_kmers; B.k = k; B.variant = variant; B.gate_m = gm;
		B.table = c->d_table; B.table_mask = (uint32_t)(cap - 1); B.filter = c->d_filter; B.filter_shift = 32 - fbits;
		B.level1 = c->d_level1; B.level0 = c->d_level0; B.err = d_err; B.pair_word_mask = pair_word_mask(c->pair_fold);
		if (n_kmers) {
			build_tables_kernel<<<(n_kmers + 255) / 256, 256, 0, st>>>(B);
			CU(c, cudaGetLastError());
			c->launches++;
		}
		CU(c, cudaMemcpyAsync(c->d_allele_off, allele_off, (2 * (size_t)n_sites + 1) * 4, cudaMemcpyHostToDevice, st));
		CU(c, cudaMemcpyAsync(h_err, d_err, sizeof h_err, cudaMemcpyDeviceToHost, st));
		CU(c, cudaStreamSynchronize(st));
		return NTSM_OK;
	};
	const int brc = build();
	cleanup();
	if (brc) return brc;
	if (h_err[0] == 1) return fail(c, NTSM_ERR_ARG, "k-mer hash %d out of range for k=%u", h_err[1], k);
	if (h_err[0] == 2) return fail(c, NTSM_ERR_ARG, "duplicate k-mer hash at index %d", h_err[1]);
	c->kernel_variant = variant;
	c->gate_m = gm;
	c->n_kmers = n_kmers;
	c->n_sites = n_sites;
	c->filter_bits = fbits;
	c->table_cap = (uint32_t)cap;
	return ntsm_reset_counts(c);
}

extern "C" int ntsm_load_siteset(ntsm_ctx *c, const ntsm_sites *s)
{
	if (!c || !s) return fail(c, NTSM_ERR_ARG, "ntsm_load_siteset: null argument");
	if (ntsm_sites_k(s) != c->cfg.k) return fail(c, NTSM_ERR_ARG, "site set k=%u but ctx k=%u", ntsm_sites_k(s), c->cfg.k);
	return ntsm_load_sites(c, ntsm_sites_hashes(s), ntsm_sites_erased(s), ntsm_sites_n_kmers(s), ntsm_sites_allele_off(s),
	                       ntsm_sites_n_sites(s));
}

static int reset_tallies(ntsm_ctx *c)
{
	std::lock_guard<std::mutex> g(c->mu);
	if (!c->inflight.empty()) return fail(c, NTSM_ERR_ARG, "reset with batches in flight; ntsm_sync first");
	if (c->d_counts) CU(c, cudaMemsetAsync(c->d_counts, 0, std::max<size_t>(1, c->n_kmers) * 4, c->compute_stream));
	CU(c, cudaMemsetAsync(c->d_totals, 0, 3 * sizeof(unsigned long long), c->compute_stream));
	c->done_kmers = c->done_hits = c->done_bases = c->submitted_bases = 0;
	c->reduced = false;
	return NTSM_OK;
}

extern "C" int ntsm_reset_counts_async(ntsm_ctx *c)
{
	if (!c) return NTSM_ERR_ARG;
	CU(c, cudaSetDevice(c->device));
	return reset_tallies(c);
}

extern "C" int ntsm_reset_counts(ntsm_ctx *c)
{
	if (!c) return NTSM_ERR_ARG;
	int rc = ntsm_sync(c);
	if (rc) return rc;
	rc = reset_tallies(c);
	if (rc) return rc;
	CU(c, cudaStreamSynchronize(c->compute_stream));
	return NTSM_OK;
}

extern "C" int ntsm_set_stream(ntsm_ctx *c, void *cuda_stream)
{
	if (!c) return NTSM_ERR_ARG;
	int rc = ntsm_sync(c);
	if (rc) return rc;
	c->compute_stream = cuda_stream ? (cudaStream_t)cuda_stream : c->own_compute;
	return NTSM_OK;
}

// ------------------------------------------------------------------ kernel launch
static int launch_count(ntsm_ctx *c, const uint2 *d_bases, const uint32_t *d_mask, uint64_t n_pos, cudaStream_t st)
{
	if (!c->d_table) return fail(c, NTSM_ERR_ARG, "no site table loaded");
	if (c->reduced) return fail(c, NTSM_ERR_ARG, "counts were already all-reduced; ntsm_reset_counts first");
	if (n_pos == 0) return NTSM_OK;
	CountParams P;
	P.bases = d_bases;
	P.nmask = d_mask;
	P.n_chunks = (n_pos + 31) / 32;
	P.minimizer = c->d_level1;
	P.minimizer2 = c->d_level1;
	P.level0 = c->d_level0;
	P.filter = c->d_filter;
	P.filter_shift = 32 - c->filter_bits;
	P.table = c->d_table;
	P.table_mask = c->table_cap - 1;
	P.k = c->cfg.k;
	P.four = 4;
	P.pair_word_mask = pair_word_mask(c->pair_fold);
	P.delta = c->launch_delta;
	P.counts = c->d_counts;
	P.totals = c->launch_totals ? c->launch_totals : c->d_totals;
	const uint64_t tiles = (P.n_chunks + kCountThreads - 1) / kCountThreads;
	const unsigned grid = (unsigned)std::min<uint64_t>(tiles, (uint64_t)c->sm_count * 16);
	const unsigned g2 = (unsigned)std::min<uint64_t>((P.n_chunks + kGateThreads - 1) / kGateThreads, (uint64_t)c->sm_count);
	if (c->cfg.k == 19 && c->kernel_variant >= 4) {
		// persistent: MINB CTAs per SM (fewer when the batch has fewer 31-chunk groups than that many CTAs have warps)
		const uint64_t groups = (P.n_chunks + kGroupChunks - 1) / kGroupChunks;
		static const int shape[4][2] = { { 1024, 1 }, { 1024, 2 }, { 512, 4 }, { 256, 8 } };
		const int th = shape[c->seed_cfg][0], mb = shape[c->seed_cfg][1];
		const uint64_t wpc = th / 32;
		const unsigned gs = (unsigned)std::min<uint64_t>((groups + wpc - 1) / wpc, (uint64_t)c->sm_count * mb);
		if (c->kernel_variant == 5) {
			if (c->seed_cfg == 1) count_kernel_pair<19, 1024, 2><<<gs, 1024, 0, st>>>(P);
			else if (c->seed_cfg == 2) count_kernel_pair<19, 512, 4><<<gs, 512, 0, st>>>(P);
			else if (c->seed_cfg == 3) count_kernel_pair<19, 256, 8><<<gs, 256, 0, st>>>(P);
			else count_kernel_pair<19, 1024, 1><<<gs, 1024, 0, st>>>(P);
		} else if (c->seed_cfg == 1) count_kernel_seed<19, kSeedM, 1024, 2><<<gs, 1024, 0, st>>>(P);
		else if (c->seed_cfg == 2) count_kernel_seed<19, kSeedM, 512, 4><<<gs, 512, 0, st>>>(P);
		else if (c->seed_cfg == 3) count_kernel_seed<19, kSeedM, 256, 8><<<gs, 256, 0, st>>>(P);
		else count_kernel_seed<19, kSeedM, 1024, 1><<<gs, 1024, 0, st>>>(P);
	} else if (c->cfg.k == 19 && c->kernel_variant == 3) {
		// one persistent CTA per SM (fewer when the batch has fewer 31-chunk groups than that many CTAs have warps)
		const uint64_t groups = (P.n_chunks + kGroupChunks - 1) / kGroupChunks, wpc = (uint64_t)(c->gate_threads == 768 || c->gate_threads == 512 ? c->gate_threads : 1024) / 32;
		const unsigned g3 = (unsigned)std::min<uint64_t>((groups + wpc - 1) / wpc, (uint64_t)c->sm_count);
		if (c->gate_m == 13) count_kernel_gate2<19, 13, true, 1024><<<g3, 1024, kL0Words * 4, st>>>(P);
		else if (!c->pool_tail) count_kernel_gate2<19, 14, false, 1024><<<g3, 1024, kL0Words * 4, st>>>(P);
		else if (c->gate_threads == 768) count_kernel_gate2<19, 14, true, 768><<<g3, 768, kL0Words * 4, st>>>(P);
		else if (c->gate_threads == 512) count_kernel_gate2<19, 14, true, 512><<<g3, 512, kL0Words * 4, st>>>(P);
		else count_kernel_gate2<19, 14, true, 1024><<<g3, 1024, kL0Words * 4, st>>>(P);
	} else if (c->cfg.k == 19 && c->kernel_variant == 2) count_kernel_gate<19><<<g2, kGateThreads, kL0Words * 4, st>>>(P);
	else if (c->cfg.k == 19 && c->kernel_variant == 1) count_kernel_min<19, kMinimizerM><<<grid, kCountThreads, 0, st>>>(P);
	else if (c->cfg.k == 19) count_kernel<19><<<grid, kCountThreads, 0, st>>>(P);
	else count_kernel<0><<<grid, kCountThreads, 0, st>>>(P);
	CU(c, cudaGetLastError());
	c->launches++;
	return NTSM_OK;
}

extern "C" int ntsm_count_packed_device(ntsm_ctx *c, const uint32_t *d_bases2, const uint32_t *d_nmask, uint64_t n_pos,
                                        uint64_t n_bases, void *cuda_stream)
{
	if (!c || !d_bases2 || !d_nmask) return fail(c, NTSM_ERR_ARG, "ntsm_count_packed_device: null argument");
	CU(c, cudaSetDevice(c->device));
	cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : c->compute_stream;
	const int rc = launch_count(c, reinterpret_cast<const uint2 *>(d_bases2), d_nmask, n_pos, st);
	if (rc == NTSM_OK) {
		std::lock_guard<std::mutex> g(c->mu);
		c->submitted_bases += n_bases;
	}
	return rc;
}

// ------------------------------------------------------------------ batches
static int make_batch(ntsm_ctx *c, ntsm_batch **out)
{
	ntsm_batch *b = new ntsm_batch();
	b->ctx = c;
	b->cap_pos = c->cfg.batch_bases & ~(kReadAlign - 1);
	const uint64_t padded = padded_positions(b->cap_pos);
	CU(c, cudaMallocHost(&b->h_bases, padded / 32 * 8));
	CU(c, cudaMallocHost(&b->h_mask, padded / 32 * 4));
	CU(c, cudaMallocHost(&b->h_snap, 16));
	CU(c, cudaMalloc(&b->d_bases, padded / 32 * 8));
	CU(c, cudaMalloc(&b->d_mask, padded / 32 * 4));
	CU(c, cudaEventCreateWithFlags(&b->copied, cudaEventDisableTiming));
	CU(c, cudaEventCreateWithFlags(&b->done, cudaEventDisableTiming));
	*out = b;
	return NTSM_OK;
}

// fold finished batches into the completed tallies; caller holds c->mu
static void reap(ntsm_ctx *c, bool block_on_oldest)
{
	while (!c->inflight.empty()) {
		ntsm_batch *b = c->inflight.front();
		cudaError_t q = cudaEventQuery(b->done);
		if (q == cudaErrorNotReady) {
			if (!block_on_oldest) break;
			cudaEventSynchronize(b->done);
			block_on_oldest = false;
		}
		c->done_kmers = b->h_snap[0];      // d_totals is cumulative and kernels run in submit order
		c->done_hits = b->h_snap[1];
		c->done_bases += b->n_bases;
		b->state = 0;
		c->inflight.pop_front();
	}
}

extern "C" int ntsm_acquire_batch(ntsm_ctx *c, ntsm_batch **out)
{
	if (!c || !out) return fail(c, NTSM_ERR_ARG, "ntsm_acquire_batch: null argument");
	CU(c, cudaSetDevice(c->device));
	std::unique_lock<std::mutex> g(c->mu);
	for (;;) {
		reap(c, false);
		for (ntsm_batch *b : c->batches)
			if (b->state == 0) {
				b->state = 1;
				b->pk.reset(b->h_bases, b->h_mask);
				b->n_bases = b->n_reads = 0;
				b->read_end.clear();
				b->read_bases.clear();
				*out = b;
				return NTSM_OK;
			}
		if (c->batches.size() < c->cfg.n_buffers) {
			ntsm_batch *b = nullptr;
			const int rc = make_batch(c, &b);
			if (rc) { free_batch(b); return rc; }
			c->batches.push_back(b);
			continue;
		}
		if (!c->inflight.empty()) reap(c, true);        // wait for the oldest kernel
		else c->cv.wait(g);                             // every buffer is held by another producer
	}
}

extern "C" int ntsm_batch_append(ntsm_batch *b, const char *seq, uint64_t len, uint64_t *pos)
{
	if (!b || !pos || (!seq && len) || *pos > len) return fail(b ? b->ctx : nullptr, NTSM_ERR_ARG, "ntsm_batch_append: bad argument");
	const uint32_t k = b->ctx->cfg.k;
	// a continued read re-packs its last k-1 consumed bases so the windows that span the cut are seen once
	const uint64_t from = *pos >= (uint64_t)(k - 1) ? *pos - (k - 1) : 0;
	const uint64_t start = *pos == 0 ? 0 : from;
	const uint64_t need = len - start;
	const uint64_t room = b->cap_pos - b->pk.pos;       // positions left (a multiple of 8, like every read's span)
	const bool capped = b->ctx->cfg.max_counts != 0;
	if (read_span(need) <= room) {
		b->pk.put_read(seq + start, need);
		b->n_bases += len - *pos;
		b->n_reads += (*pos == 0);
		*pos = len;
		if (capped) {                                        // the read's separator sits right after its last base
			b->read_end.push_back((uint32_t)(b->pk.pos - read_span(need) + need));
			b->read_bases.push_back(b->n_bases);
		}
		return 1;
	}
	// with a -m cap a read is kept whole (the stop is decided read by read) unless it is longer than a batch
	if (capped && b->pk.pos != 0) return 0;
	const uint64_t split_min = std::max<uint64_t>(2 * k, std::min<uint64_t>(4096, b->cap_pos / 4));
	if (b->pk.pos != 0 && room < split_min + 1) return 0;   // full: submit and come back
	const uint64_t take = room - 1;                          // spans exactly `room`; >= 2k > k-1, so the read always advances
	b->pk.put_read(seq + start, take);
	b->n_bases += start + take - *pos;
	b->n_reads += (*pos == 0);
	*pos = start + take;
	return 0;
}

extern "C" uint64_t ntsm_batch_positions(const ntsm_batch *b) { return b->pk.pos; }
extern "C" uint64_t ntsm_batch_bases(const ntsm_batch *b) { return b->n_bases; }
extern "C" uint64_t ntsm_batch_reads(const ntsm_batch *b) { return b->n_reads; }

// enqueue H2D of a packed stream (src_* = host memory holding at least padded/halo words) into the
// batch's device buffers, the count kernel and the tally snapshot; caller holds no lock
static int enqueue_batch(ntsm_ctx *c, ntsm_batch *b, const void *src_bases, const void *src_mask, uint64_t n_pos,
                         uint64_t copy_pos, uint64_t n_bases)
{
	std::lock_guard<std::mutex> g(c->mu);
	if (n_pos == 0) {
		b->state = 0;
		c->cv.notify_all();
		return NTSM_OK;
	}
	CU(c, cudaMemcpyAsync(b->d_bases, src_bases, copy_pos / 32 * 8, cudaMemcpyHostToDevice, c->copy_stream));
	CU(c, cudaMemcpyAsync(b->d_mask, src_mask, copy_pos / 32 * 4, cudaMemcpyHostToDevice, c->copy_stream));
	CU(c, cudaEventRecord(b->copied, c->copy_stream));
	CU(c, cudaStreamWaitEvent(c->compute_stream, b->copied, 0));
	const int rc = launch_count(c, b->d_bases, b->d_mask, n_pos, c->compute_stream);
	if (rc) return rc;
	CU(c, cudaMemcpyAsync(b->h_snap, c->d_totals, 16, cudaMemcpyDeviceToHost, c->compute_stream));
	CU(c, cudaEventRecord(b->done, c->compute_stream));
	b->state = 2;
	b->n_bases = n_bases;
	b->n_pos = n_pos;
	c->last_batch = b;
	c->submitted_bases += n_bases;
	c->inflight.push_back(b);
	c->cv.notify_all();
	return NTSM_OK;
}

extern "C" int ntsm_submit_batch(ntsm_ctx *c, ntsm_batch *b)
{
	if (!c || !b || b->ctx != c || b->state != 1) return fail(c, NTSM_ERR_ARG, "ntsm_submit_batch: batch not acquired from this ctx");
	CU(c, cudaSetDevice(c->device));
	const uint64_t n_pos = b->pk.finish();
	return enqueue_batch(c, b, b->h_bases, b->h_mask, n_pos, padded_positions(n_pos), b->n_bases);
}

extern "C" int ntsm_count_packed_host(ntsm_ctx *c, const uint32_t *h_bases2, const uint32_t *h_nmask, uint64_t n_pos,
                                      uint64_t n_bases)
{
	if (!c || !h_bases2 || !h_nmask) return fail(c, NTSM_ERR_ARG, "ntsm_count_packed_host: null argument");
	CU(c, cudaSetDevice(c->device));
	const uint64_t cap = c->cfg.batch_bases / kTilePositions * kTilePositions;   // slice length, a multiple of 32
	if (cap == 0) return fail(c, NTSM_ERR_ARG, "batch_bases too small");
	for (uint64_t off = 0; off < n_pos; off += cap) {
		const uint64_t len = std::min(cap, n_pos - off);
		ntsm_batch *b = nullptr;
		int rc = ntsm_acquire_batch(c, &b);
		if (rc) return rc;
		// a slice needs the 64 positions after it (windows that start inside and end outside)
		const uint64_t copy_pos = (len + 31) / 32 * 32 + kHaloPositions;
		rc = enqueue_batch(c, b, h_bases2 + off / 16, h_nmask + off / 32, len, copy_pos, off == 0 ? n_bases : 0);
		if (rc) return rc;
	}
	return NTSM_OK;
}

extern "C" int ntsm_release_batch(ntsm_ctx *c, ntsm_batch *b)
{
	if (!c || !b || b->ctx != c || b->state != 1) return fail(c, NTSM_ERR_ARG, "ntsm_release_batch: batch not acquired from this ctx");
	std::lock_guard<std::mutex> g(c->mu);
	b->state = 0;
	c->cv.notify_one();
	return NTSM_OK;
}

extern "C" int ntsm_insert_count(ntsm_ctx *c, const char *seq, uint64_t len)
{
	if (!c) return NTSM_ERR_ARG;
	uint64_t pos = 0;
	for (;;) {
		if (!c->current) {
			const int rc = ntsm_acquire_batch(c, &c->current);
			if (rc) return rc;
		}
		const int r = ntsm_batch_append(c->current, seq, len, &pos);
		if (r < 0) return r;
		if (r == 1) return NTSM_OK;
		ntsm_batch *b = c->current;
		c->current = nullptr;
		const int rc = ntsm_submit_batch(c, b);
		if (rc) return rc;
	}
}

extern "C" int ntsm_flush(ntsm_ctx *c)
{
	if (!c) return NTSM_ERR_ARG;
	if (!c->current) return NTSM_OK;
	ntsm_batch *b = c->current;
	c->current = nullptr;
	return ntsm_submit_batch(c, b);
}

extern "C" int ntsm_poll_totals(ntsm_ctx *c, uint64_t *total_kmers, uint64_t *total_hits, uint64_t *total_bases,
                                int *cap_reached)
{
	if (!c) return NTSM_ERR_ARG;
	CU(c, cudaSetDevice(c->device));
	std::lock_guard<std::mutex> g(c->mu);
	reap(c, false);
	if (total_kmers) *total_kmers = c->done_kmers;
	if (total_hits) *total_hits = c->done_hits;
	if (total_bases) *total_bases = c->done_bases;
	if (cap_reached) *cap_reached = c->cfg.max_counts != 0 && c->done_hits > c->cfg.max_counts;   // FingerPrint.hpp:476
	return NTSM_OK;
}

extern "C" int ntsm_sync(ntsm_ctx *c)
{
	if (!c) return NTSM_ERR_ARG;
	CU(c, cudaSetDevice(c->device));
	CU(c, cudaDeviceSynchronize());
	std::lock_guard<std::mutex> g(c->mu);
	reap(c, false);
	c->cv.notify_all();
	return NTSM_OK;
}

namespace ntsm {
__global__ void set_u64_kernel(unsigned long long *p, unsigned long long v) { *p = v; }
}  // namespace ntsm

// ------------------------------------------------------------------ exact -m stop
// The reference checks the cap after EVERY read (processSingleRead, src/FingerPrint.hpp:473-488): the
// counted reads are the shortest prefix whose hits exceed it.  Batches are what the GPU sees, so when
// the batch just completed pushed the summed tally over the cap, find the read inside it where that
// happened and make the counters what they would be had the batch ended there: re-run prefixes of the
// batch (still resident on the device) with delta 0 into a scratch tally to bisect over the recorded
// read ends, then take the whole batch back out (delta -1) and put the prefix back in (delta +1).
// A prefix ends at a read's separator; the positions after it that share its 32-position chunk are
// masked off by patching that one mask word for the duration of a launch.
// hits_elsewhere = hits on the other GPUs.  Returns 1 if the counters now stop exactly after the
// deciding read, 0 if nothing could be done (no read ends recorded: the stop stays batch-granular).
int ntsm_trim_to_cap(ntsm_ctx *c, uint64_t hits_elsewhere, uint64_t cap)
{
	if (!c) return NTSM_ERR_ARG;
	CU(c, cudaSetDevice(c->device));
	std::lock_guard<std::mutex> g(c->mu);
	ntsm_batch *b = c->last_batch;
	if (!b || b->state != 0 || b->read_end.empty() || !c->inflight.empty()) return 0;
	if (!c->d_scratch_totals) CU(c, cudaMalloc(&c->d_scratch_totals, 3 * sizeof(unsigned long long)));
	cudaStream_t st = c->compute_stream;
	unsigned long long out[2] = { 0, 0 };
	// tallies {TK, hits} of positions [0, e) of the batch, adding `delta` per hit to the counters
	auto run_prefix = [&](uint64_t e, uint32_t delta) -> int {
		const uint64_t ce = e / 32;
		const uint32_t r = (uint32_t)(e & 31);
		const uint32_t patched = r ? (b->h_mask[ce] | (~0u << r)) : 0;
		if (r) CU(c, cudaMemcpyAsync(b->d_mask + ce, &patched, 4, cudaMemcpyHostToDevice, st));
		CU(c, cudaMemsetAsync(c->d_scratch_totals, 0, 3 * sizeof(unsigned long long), st));
		c->launch_delta = delta;
		c->launch_totals = c->d_scratch_totals;
		const int rc = launch_count(c, b->d_bases, b->d_mask, e, st);
		c->launch_delta = 1;
		c->launch_totals = nullptr;
		if (rc) return rc;
		CU(c, cudaMemcpyAsync(out, c->d_scratch_totals, sizeof out, cudaMemcpyDeviceToHost, st));
		if (r) CU(c, cudaMemcpyAsync(b->d_mask + ce, b->h_mask + ce, 4, cudaMemcpyHostToDevice, st));
		CU(c, cudaStreamSynchronize(st));
		return NTSM_OK;
	};
	int rc = run_prefix(b->n_pos, 0);
	if (rc) return rc;
	const uint64_t tk_b = out[0], hits_b = out[1];
	const uint64_t total = hits_elsewhere + c->done_hits;
	if (total <= cap || hits_b > total) return 0;
	const uint64_t base = total - hits_b;                      // summed tally before this batch
	if (base > cap) return 0;                                   // an earlier batch already decided; nothing to trim here
	const size_t n = b->read_end.size();
	if ((rc = run_prefix(b->read_end[n - 1], 0))) return rc;
	if (base + out[1] <= cap) return 0;                         // the cap falls inside a read longer than a batch
	size_t lo = 0, hi = n - 1;                                  // smallest i with base + hits(prefix i) > cap
	while (lo < hi) {
		const size_t mid = lo + (hi - lo) / 2;
		if ((rc = run_prefix(b->read_end[mid], 0))) return rc;
		if (base + out[1] > cap) hi = mid;
		else lo = mid + 1;
	}
	if ((rc = run_prefix(b->n_pos, 0xFFFFFFFFu))) return rc;   // the whole batch out ...
	if ((rc = run_prefix(b->read_end[lo], 1))) return rc;      // ... its deciding prefix back in
	const uint64_t tk_p = out[0], hits_p = out[1];
	c->done_kmers = c->done_kmers - tk_b + tk_p;
	c->done_hits = c->done_hits - hits_b + hits_p;
	set_u64_kernel<<<1, 1, 0, st>>>(c->d_totals + 0, (unsigned long long)c->done_kmers);
	set_u64_kernel<<<1, 1, 0, st>>>(c->d_totals + 1, (unsigned long long)c->done_hits);
	CU(c, cudaGetLastError());
	CU(c, cudaStreamSynchronize(st));
	c->launches += 2;
	const uint64_t dropped = b->n_bases - b->read_bases[lo];
	c->done_bases -= dropped;
	c->submitted_bases -= dropped;
	b->read_end.clear();
	return 1;
}

// ------------------------------------------------------------------ multi-GPU
// stdout is the counts file, so NCCL's version banner must not land there.  NCCL honours
// NCCL_DEBUG_FILE only for levels above VERSION (2.27: debug.cc), and GPU hosts commonly export
// NCCL_DEBUG=VERSION: raise that to WARN (which prints the same banner) so the file setting applies.
static void nccl_banner_to_stderr()
{
	const char *lvl = getenv("NCCL_DEBUG");
	if (lvl && !strcasecmp(lvl, "VERSION")) setenv("NCCL_DEBUG", "WARN", 1);
	setenv("NCCL_DEBUG_FILE", "/dev/stderr", 0);
}

extern "C" int ntsm_nccl_unique_id(void *id_out)
{
	static_assert(sizeof(ncclUniqueId) == NTSM_NCCL_ID_BYTES, "ncclUniqueId size");
	if (!id_out) return NTSM_ERR_ARG;
	ncclUniqueId id;
	nccl_banner_to_stderr();
	NC(nullptr, ncclGetUniqueId(&id));
	memcpy(id_out, &id, sizeof id);
	return NTSM_OK;
}

extern "C" int ntsm_comm_init(ntsm_ctx *c, const void *id, int rank, int n_ranks)
{
	if (!c || !id || rank < 0 || rank >= n_ranks) return fail(c, NTSM_ERR_ARG, "ntsm_comm_init: bad argument");
	CU(c, cudaSetDevice(c->device));
	ncclUniqueId uid;
	memcpy(&uid, id, sizeof uid);
	nccl_banner_to_stderr();
	NC(c, ncclCommInitRank(&c->comm, n_ranks, uid, rank));
	c->rank = rank;
	c->n_ranks = n_ranks;
	return NTSM_OK;
}

// enqueue on the compute stream: base tally -> device, the two all-reduces (if a communicator is
// attached), the per-site reduce.  No host synchronisation.
extern "C" int ntsm_reduce_async(ntsm_ctx *c)
{
	if (!c) return NTSM_ERR_ARG;
	if (!c->d_table) return fail(c, NTSM_ERR_ARG, "no site table loaded");
	if (c->reduced) return NTSM_OK;
	CU(c, cudaSetDevice(c->device));
	// the private base tally joins the two device tallies so ONE u64 all-reduce covers TK/hits/bases
	set_u64_kernel<<<1, 1, 0, c->compute_stream>>>(c->d_totals + 2, (unsigned long long)c->submitted_bases);
	CU(c, cudaGetLastError());
	c->launches++;
	if (c->comm && c->n_ranks > 1) {
		// sums first, per-site max afterwards: max of sums != sum of maxes
		NC(c, ncclGroupStart());
		NC(c, ncclAllReduce(c->d_counts, c->d_counts, c->n_kmers, ncclUint32, ncclSum, c->comm, c->compute_stream));
		NC(c, ncclAllReduce(c->d_totals, c->d_totals, 3, ncclUint64, ncclSum, c->comm, c->compute_stream));
		NC(c, ncclGroupEnd());
	}
	const uint32_t S = c->n_sites;
	if (S) {
		uint32_t *r = c->d_rows;
		site_reduce_kernel<<<(S + 255) / 256, 256, 0, c->compute_stream>>>(c->d_counts, c->d_allele_off, S, r, r + S,
		                                                                   r + 2 * (size_t)S, r + 3 * (size_t)S);
		CU(c, cudaGetLastError());
		c->launches++;
	}
	c->reduced = true;
	return NTSM_OK;
}

extern "C" int ntsm_allreduce(ntsm_ctx *c)
{
	if (!c) return NTSM_ERR_ARG;
	int rc = ntsm_flush(c);
	if (rc) return rc;
	rc = ntsm_sync(c);
	if (rc) return rc;
	rc = ntsm_reduce_async(c);
	if (rc) return rc;
	CU(c, cudaStreamSynchronize(c->compute_stream));
	return NTSM_OK;
}

// ------------------------------------------------------------------ results
extern "C" int ntsm_finalize(ntsm_ctx *c, uint32_t *max_ref, uint32_t *max_var, uint32_t *sum_ref, uint32_t *sum_var,
                             uint64_t totals[3])
{
	if (!c) return NTSM_ERR_ARG;
	int rc = ntsm_flush(c);
	if (rc) return rc;
	{   // drain the batch ring without a device-wide sync (other streams may be busy)
		std::unique_lock<std::mutex> g(c->mu);
		while (!c->inflight.empty()) reap(c, true);
		c->cv.notify_all();
	}
	rc = ntsm_reduce_async(c);
	if (rc) return rc;
	const uint32_t S = c->n_sites;
	uint32_t *dst[4] = { max_ref, max_var, sum_ref, sum_var };
	for (int i = 0; i < 4 && S; ++i)
		if (dst[i]) CU(c, cudaMemcpyAsync(dst[i], c->d_rows + (size_t)i * S, (size_t)S * 4, cudaMemcpyDeviceToHost, c->compute_stream));
	unsigned long long t[3] = { 0, 0, 0 };
	CU(c, cudaMemcpyAsync(t, c->d_totals, sizeof t, cudaMemcpyDeviceToHost, c->compute_stream));
	CU(c, cudaStreamSynchronize(c->compute_stream));
	if (totals) { totals[0] = t[0]; totals[1] = t[1]; totals[2] = t[2]; }
	return NTSM_OK;
}

extern "C" int ntsm_get_counts(ntsm_ctx *c, uint32_t *counts)
{
	if (!c || !counts) return NTSM_ERR_ARG;
	int rc = ntsm_flush(c);
	if (rc) return rc;
	rc = ntsm_sync(c);
	if (rc) return rc;
	CU(c, cudaMemcpy(counts, c->d_counts, (size_t)c->n_kmers * 4, cudaMemcpyDeviceToHost));
	return NTSM_OK;
}

// library-internal helpers (not part of the public header)
uint64_t ntsm_ctx_max_counts(const ntsm_ctx *c) { return c->cfg.max_counts; }
uint64_t ntsm_ctx_batch_bases(const ntsm_ctx *c) { return c->cfg.batch_bases; }
void ntsm_set_thread_error(const char *text) { t_last_error = text; }

extern "C" uint64_t ntsm_ctx_launches(const ntsm_ctx *c) { return c ? c->launches : 0; }
extern "C" const char *ntsm_ctx_kernel_name(const ntsm_ctx *c)
{
	if (!c) return "";
	if (c->cfg.k != 19 || c->kernel_variant == 0) return c->cfg.k == 19 ? "count_kernel<19>" : "count_kernel<0>";
	if (c->kernel_variant == 1) return "count_kernel_min<19,13>";
	if (c->kernel_variant == 2) return "count_kernel_gate<19>";
	if (c->kernel_variant == 5) {
		static const char *names[4] = { "count_kernel_pair<19,1024,1>", "count_kernel_pair<19,1024,2>", "count_kernel_pair<19,512,4>", "count_kernel_pair<19,256,8>" };
		return names[c->seed_cfg];
	}
	if (c->kernel_variant == 4) {
		static const char *names[4] = { "count_kernel_seed<19,14,1024,1>", "count_kernel_seed<19,14,1024,2>", "count_kernel_seed<19,14,512,4>", "count_kernel_seed<19,14,256,8>" };
		return names[c->seed_cfg];
	}
	return c->gate_m == 13 ? "count_kernel_gate2<19,13,1>" : c->pool_tail ? "count_kernel_gate2<19,14,1>" : "count_kernel_gate2<19,14,0>";
}
extern "C" uint32_t ntsm_ctx_filter_bits(const ntsm_ctx *c) { return c ? c->filter_bits : 0; }
extern "C" uint32_t ntsm_ctx_table_capacity(const ntsm_ctx *c) { return c ? c->table_cap : 0; }
