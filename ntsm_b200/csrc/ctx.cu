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
#include <atomic>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/ntsm_b200.h"
#include "internal.h"
#include "kernels.cuh"
#include "pair.cuh"
#include "devpack.cuh"
#include "pack.h"
#include "numa.h"

using namespace ntsm;

static thread_local std::string t_last_error;

struct ntsm_batch {
	ntsm_ctx *ctx = nullptr;
	uint64_t *h_bases = nullptr;      // pinned
	uint32_t *h_mask = nullptr;       // pinned
	uint64_t *h_snap = nullptr;       // pinned {TK, hits} snapshot taken after this batch's kernel
	uint2 *d_bases = nullptr;
	uint32_t *d_mask = nullptr;
	uint8_t *d_ascii = nullptr;       // device staging for reads that arrive as ASCII and are packed on the GPU (lazily allocated)
	uint32_t *d_aux = nullptr;        // per-read input offsets + output positions of a variable-length ASCII batch
	uint32_t *h_aux = nullptr;        // pinned twin of d_aux, filled by the feeder thread
	uint64_t aux_cap = 0;             // reads d_aux / h_aux hold
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
	int numa_node = -1;                     // the GPU's NUMA node when the host has more than one and says which (numa.h)
	// options (ntsm_ctx_set_option, before ntsm_load_sites); -1 / 0 = decide from the panel
	int opt_kernel = -1;                    // 0 generic (one k-mer-bitmap probe per position), 1 paired seeds, 2 wide paired seeds (large panels)
	int opt_pair_fold = -1;                 // paired-seed table folded 2^fold : 1
	int opt_filter_bits = 0;                // log2 bits of the k-mer bitmap
	int opt_shape = 1;                      // pair kernel launch shape: 0 = 1024x1, 1 = 1024x2 (default), 2 = 512x4, 3 = 256x8
	int opt_l2_persist = 0;                 // launch with an L2 access-policy window over the probe tables (measured: no effect, see ntsm_load_sites)
	int opt_parser_procs = -1;              // file pipeline: parsers as worker processes (procpipe.h): -1 = beyond 6 parser threads on plain files
	int opt_device_pack = -1;               // bulk inserts from page-locked memory also feed ASCII to the GPU packer: -1 = when host packers are few
	// site table
	uint32_t n_kmers = 0, n_sites = 0;
	uint32_t *d_probe = nullptr;            // ONE allocation: paired-seed table, then the k-mer bitmap (one L2 access-policy window covers both)
	uint32_t *d_pair = nullptr, *d_filter = nullptr;   // into d_probe
	size_t probe_bytes = 0;
	int kernel = 0;                         // what launch_count runs: 0 generic, 1 paired seeds
	int pair_m = 0, pair_fold = 0;
	bool l2_window = false;                 // launches carry an access-policy window over d_probe
	float l2_hit_ratio = 1.0f;
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
	std::mutex mu;                          // batch states, the in-flight queue, the completed tallies
	std::mutex submit_mu;                   // one producer at a time enqueues copy + kernel + snapshot (stream order = queue order)
	std::condition_variable cv;
	ntsm_batch *current = nullptr;          // ntsm_insert_count's open batch
	// tallies
	uint64_t done_kmers = 0, done_hits = 0, done_bases = 0;   // over completed batches
	uint64_t submitted_bases = 0;
	uint64_t submitted_reads = 0;           // reads begun in the packed batches submitted so far (m_totalReads, src/FingerPrint.hpp:72)
	std::atomic<uint64_t> launches{0};
	std::atomic<uint64_t> h2d_bytes{0}, d2h_bytes{0};  // copied over PCIe by this ctx's data path so far (batches, snapshots, result rows)
	int async_error = 0;                    // a batch's event reported a device fault: sticky until the ctx is destroyed
	// exact -m stop: the last batch submitted, a scratch tally, and what launch_count adds per hit
	ntsm_batch *last_batch = nullptr;
	unsigned long long *d_scratch_totals = nullptr;
	uint32_t launch_delta = 1;
	unsigned long long *launch_totals = nullptr;   // nullptr = d_totals
	bool reduced = false;
	cudaEvent_t drained = nullptr;          // ntsm_group_finalize: this ctx's counts are final
	// ASCII (device-packed) batches are 2.7x the bytes of a packed one: at most two of their copies are queued at
	// a time, so the host packers' small batches never wait behind a long line of them on the copy stream
	cudaEvent_t ascii_ev[2] = { nullptr, nullptr };
	uint64_t ascii_jobs = 0;
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
extern "C" void ntsm_ctx_destroy(ntsm_ctx *c);

extern "C" int ntsm_ctx_create(ntsm_ctx **out, const ntsm_cfg *cfg)
{
	if (!out || !cfg) return fail(nullptr, NTSM_ERR_ARG, "ntsm_ctx_create: null argument");
	*out = nullptr;
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
	// stream positions inside one batch are 32-bit (read ends recorded for the exact -m stop, the
	// device packer's per-read offsets): 2^31 positions per batch is far beyond any useful size
	if (c->cfg.batch_bases > (1ull << 31)) c->cfg.batch_bases = 1ull << 31;
	c->device = cfg->device;
	auto init = [&]() -> int {
		CU(c, cudaSetDevice(c->device));
		CU(c, cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, c->device));
		c->numa_node = gpu_numa_node(c->device);
		CU(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
		CU(c, cudaStreamCreateWithFlags(&c->own_compute, cudaStreamNonBlocking));
		c->compute_stream = c->own_compute;
		CU(c, cudaEventCreateWithFlags(&c->drained, cudaEventDisableTiming));
		CU(c, cudaEventCreateWithFlags(&c->ascii_ev[0], cudaEventDisableTiming | cudaEventBlockingSync));
		CU(c, cudaEventCreateWithFlags(&c->ascii_ev[1], cudaEventDisableTiming | cudaEventBlockingSync));
		CU(c, cudaMalloc(&c->d_totals, 3 * sizeof(unsigned long long)));
		CU(c, cudaMemset(c->d_totals, 0, 3 * sizeof(unsigned long long)));
		return NTSM_OK;
	};
	const int rc = init();
	if (rc) {
		ntsm_ctx_destroy(c);      // the text of the failure stays in this thread's last error
		return rc;
	}
	*out = c;
	return NTSM_OK;
}

static void free_batch(ntsm_batch *b)
{
	if (!b) return;
	cudaFreeHost(b->h_bases);
	cudaFreeHost(b->h_mask);
	cudaFreeHost(b->h_snap);
	cudaFreeHost(b->h_aux);
	cudaFree(b->d_bases);
	cudaFree(b->d_mask);
	cudaFree(b->d_ascii);
	cudaFree(b->d_aux);
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
	cudaFree(c->d_probe);
	cudaFree(c->d_table);
	cudaFree(c->d_counts);
	cudaFree(c->d_allele_off);
	cudaFree(c->d_rows);
	cudaFree(c->d_totals);
	cudaFree(c->d_scratch_totals);
	if (c->drained) cudaEventDestroy(c->drained);
	for (cudaEvent_t e : c->ascii_ev) if (e) cudaEventDestroy(e);
	if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
	if (c->own_compute) cudaStreamDestroy(c->own_compute);
	delete c;
}

// Measurement / test knobs, set before ntsm_load_sites (every setting gives the same counts: the
// pre-filters only ever let too many windows through).  Replaces round 1's NTSM_* environment reads.
extern "C" int ntsm_ctx_set_option(ntsm_ctx *c, const char *name, int value)
{
	if (!c || !name) return fail(c, NTSM_ERR_ARG, "ntsm_ctx_set_option: null argument");
	if (!strcmp(name, "kernel")) c->opt_kernel = value < 0 ? -1 : std::min(2, value);
	else if (!strcmp(name, "pair_fold")) c->opt_pair_fold = value < 0 ? -1 : std::min(kPairFoldMax, value);
	else if (!strcmp(name, "filter_bits")) c->opt_filter_bits = value <= 0 ? 0 : std::min(32, std::max(10, value));
	else if (!strcmp(name, "launch_shape")) c->opt_shape = std::min(3, std::max(0, value));
	else if (!strcmp(name, "l2_persist")) c->opt_l2_persist = value;
	else if (!strcmp(name, "device_pack")) c->opt_device_pack = value;
	else if (!strcmp(name, "parser_procs")) c->opt_parser_procs = value;
	else return fail(c, NTSM_ERR_ARG, "ntsm_ctx_set_option: unknown option '%s'", name);
	return NTSM_OK;
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

	// which count kernel will run decides which pre-filter structures are built
	// (k >= 17: paired seeds.  The wide table sized for HBM -- option kernel = 2, k >= 19 -- was built for panels of
	// millions of k-mers and measured no faster than the 14-mer table there: 295 vs 319 Gbases/s, pair.cuh.)
	int kernel = k >= (uint32_t)kPairMinK ? 1 : 0;
	if (c->opt_kernel >= 0 && kernel) kernel = c->opt_kernel == 2 && k < 19 ? 1 : c->opt_kernel;
	const int pm = kernel == 2 ? kWideM : pair_seed_len((int)k);
	// panels far larger than the human one (cfg 5: 26 M k-mers) saturate a folded pair table and nothing
	// stays in L2 anyway: measured 318 (unfolded) vs 241 Gbases/s (profiles/r01v12_sweep_cfg5.jsonl)
	int fold = live > 5000000 ? 0 : kPairFoldDefault;
	if (c->opt_pair_fold >= 0) fold = c->opt_pair_fold;
	while (fold > 0 && (pair_words(pm) >> fold) < 1024) --fold;

	// k-mer bitmap holding both orientations of every live k-mer, two bits each: 20-40 bits per key (8 MiB for
	// the human panel; 2^24..2^28 bits measured within 3 % of each other, 2^26 best: profiles/r02e_sweep_filterbits.jsonl)
	uint32_t fbits = 16;
	while (fbits < 30 && (1ull << fbits) < 20ull * 2ull * live) ++fbits;
	if (c->opt_filter_bits) fbits = (uint32_t)c->opt_filter_bits;
	const size_t filter_words = (1ull << fbits) / 32;
	const size_t pair_n = kernel == 2 ? kWideBytes / 4 : kernel == 1 ? pair_words(pm) >> fold : 0;

	cudaFree(c->d_probe); cudaFree(c->d_table); cudaFree(c->d_counts); cudaFree(c->d_allele_off); cudaFree(c->d_rows);
	c->d_probe = c->d_pair = c->d_filter = nullptr; c->d_table = nullptr; c->d_counts = nullptr; c->d_allele_off = nullptr; c->d_rows = nullptr;
	{
		// The pair probe forms addresses as {lo32(base) + offset, hi32(base)}: the table must not cross a
		// 4 GiB line.  cudaMalloc hands out 2 MiB-aligned blocks, so a second try always fits.
		const size_t bytes = (pair_n + filter_words) * 4;
		std::vector<void *> rejected;
		for (int attempt = 0; attempt < 8; ++attempt) {
			CU(c, cudaMalloc(&c->d_probe, bytes));
			const uintptr_t a = (uintptr_t)c->d_probe, z = a + std::max<size_t>(4, pair_n * 4) - 1;
			if (kernel != 1 || (a >> 32) == (z >> 32)) break;
			rejected.push_back(c->d_probe);
			c->d_probe = nullptr;
		}
		for (void *r : rejected) cudaFree(r);
		if (!c->d_probe) return fail(c, NTSM_ERR_CUDA, "could not place the paired-seed table inside one 4 GiB window");
		c->probe_bytes = bytes;
		c->d_pair = pair_n ? c->d_probe : nullptr;
		c->d_filter = c->d_probe + pair_n;
	}
	CU(c, cudaMalloc(&c->d_table, cap * sizeof(TableSlot)));
	CU(c, cudaMalloc(&c->d_counts, std::max<size_t>(1, n_kmers) * 4));
	CU(c, cudaMalloc(&c->d_allele_off, (2 * (size_t)n_sites + 1) * 4));
	CU(c, cudaMalloc(&c->d_rows, std::max<size_t>(1, n_sites) * 16));

	// L2 residency: the probe tables are hit at random by every warp while the packed reads stream
	// through the same L2 once.  The reads are loaded with an evict-first hint (ld.global.cs), which is
	// what keeps the tables in.  An access-policy window that marks the probe tables as persisting on
	// top of that ("l2_persist", off by default) was measured and changes nothing for the human panel --
	// same 83 % sector hit rate, same 3.51 vs 3.52 GB of DRAM reads per 6 Gbases, same 3.05 ms
	// (profiles/r02d_count_ncu_persist{0,1}.txt): the misses that remain are capacity (48 MiB of tables
	// against an L2 whose two halves each keep their own copy of far-die lines), not eviction by the
	// stream -- and costs 7 % for the 10^6-site panel, whose tables do not fit the set-aside (296 vs 318 Gbases/s).
	c->l2_window = false;
	if (c->opt_l2_persist) {
		int max_persist = 0, max_window = 0;
		cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, c->device);
		cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, c->device);
		if (max_persist > 0 && max_window > 0) {
			const size_t want = std::min<size_t>(c->probe_bytes, (size_t)max_persist);
			size_t have = 0;
			cudaDeviceGetLimit(&have, cudaLimitPersistingL2CacheSize);
			if (have < want && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) != cudaSuccess) cudaGetLastError();
			cudaDeviceGetLimit(&have, cudaLimitPersistingL2CacheSize);
			if (have > 0 && c->probe_bytes <= (size_t)max_window) {
				c->l2_window = true;
				c->l2_hit_ratio = (float)std::min(1.0, (double)have / (double)c->probe_bytes);
			}
		}
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
		CU(c, cudaMemsetAsync(c->d_probe, 0, c->probe_bytes, st));
		BuildParams B;
		B.hash = d_hash; B.erased = d_erased; B.n_kmers = n_kmers; B.k = k;
		B.table = c->d_table; B.table_mask = (uint32_t)(cap - 1); B.filter = c->d_filter; B.filter_shift = 32 - fbits;
		B.pair = c->d_pair; B.wide = kernel == 2; B.pair_m = (uint32_t)pm; B.pair_word_mask = kernel == 1 ? pair_word_mask(pm, fold) : 0; B.err = d_err;
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
	c->kernel = kernel;
	c->pair_m = pm;
	c->pair_fold = fold;
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
	c->done_kmers = c->done_hits = c->done_bases = c->submitted_bases = c->submitted_reads = 0;
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
template <class Kern>
static cudaError_t launch_with_window(ntsm_ctx *c, Kern kern, unsigned grid, unsigned block, cudaStream_t st, const CountParams &P)
{
	cudaLaunchConfig_t lc = {};
	lc.gridDim = dim3(grid);
	lc.blockDim = dim3(block);
	lc.stream = st;
	cudaLaunchAttribute at[1];
	if (c->l2_window) {
		at[0].id = cudaLaunchAttributeAccessPolicyWindow;
		at[0].val.accessPolicyWindow.base_ptr = c->d_probe;
		at[0].val.accessPolicyWindow.num_bytes = c->probe_bytes;
		at[0].val.accessPolicyWindow.hitRatio = c->l2_hit_ratio;
		at[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
		at[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
		lc.attrs = at;
		lc.numAttrs = 1;
	}
	return cudaLaunchKernelEx(&lc, kern, P);
}

static int launch_count(ntsm_ctx *c, const uint2 *d_bases, const uint32_t *d_mask, uint64_t n_pos, cudaStream_t st)
{
	if (!c->d_table) return fail(c, NTSM_ERR_ARG, "no site table loaded");
	if (c->reduced) return fail(c, NTSM_ERR_ARG, "counts were already combined; ntsm_reset_counts first");
	if (n_pos == 0) return NTSM_OK;
	CountParams P;
	P.bases = d_bases;
	P.nmask = d_mask;
	P.n_chunks = (n_pos + 31) / 32;
	P.pair = c->d_pair;
	P.pair_off_mask = c->kernel == 1 ? pair_word_mask(c->pair_m, c->pair_fold) << 2 : 0;
	P.pair_bshift = 2 * (uint32_t)c->pair_m;
	P.filter = c->d_filter;
	P.filter_shift = 32 - c->filter_bits;
	P.table = c->d_table;
	P.table_mask = c->table_cap - 1;
	P.k = c->cfg.k;
	P.delta = c->launch_delta;
	P.counts = c->d_counts;
	P.totals = c->launch_totals ? c->launch_totals : c->d_totals;
	cudaError_t le;
	if (c->kernel == 2) {
		const uint64_t groups = (P.n_chunks + kGroupChunks - 1) / kGroupChunks;
		// one CTA of 1024 threads per SM with up to 64 registers each: the eight probe loads of a chunk need their
		// own registers to be in flight together (at 32 registers ptxas reuses them and the loads queue up)
		const unsigned gs1 = (unsigned)std::min<uint64_t>((groups + 31) / 32, (uint64_t)c->sm_count);
		le = launch_with_window(c, count_kernel_wide<1024, 1>, gs1, 1024, st, P);
	} else if (c->kernel == 1) {
		// persistent: MINB CTAs per SM (fewer when the batch has fewer 31-chunk groups than that many CTAs have warps)
		const uint64_t groups = (P.n_chunks + kGroupChunks - 1) / kGroupChunks;
		static const int shape[4][2] = { { 1024, 1 }, { 1024, 2 }, { 512, 4 }, { 256, 8 } };
		const int th = shape[c->opt_shape][0], mb = shape[c->opt_shape][1];
		const uint64_t wpc = th / 32;
		const unsigned gs = (unsigned)std::min<uint64_t>((groups + wpc - 1) / wpc, (uint64_t)c->sm_count * mb);
		if (c->cfg.k == 19) {
			if (c->opt_shape == 1) le = launch_with_window(c, count_kernel_pair<19, 1024, 2>, gs, 1024, st, P);
			else if (c->opt_shape == 2) le = launch_with_window(c, count_kernel_pair<19, 512, 4>, gs, 512, st, P);
			else if (c->opt_shape == 3) le = launch_with_window(c, count_kernel_pair<19, 256, 8>, gs, 256, st, P);
			else le = launch_with_window(c, count_kernel_pair<19, 1024, 1>, gs, 1024, st, P);
		} else {
			const unsigned g0 = (unsigned)std::min<uint64_t>((groups + 31) / 32, (uint64_t)c->sm_count * 2);
			le = launch_with_window(c, count_kernel_pair<0, 1024, 2>, g0, 1024, st, P);
		}
	} else {
		const uint64_t tiles = (P.n_chunks + kCountThreads - 1) / kCountThreads;
		const unsigned grid = (unsigned)std::min<uint64_t>(tiles, (uint64_t)c->sm_count * 16);
		le = launch_with_window(c, count_kernel_generic, grid, kCountThreads, st, P);
	}
	if (le != cudaSuccess) return fail(c, NTSM_ERR_CUDA, "count kernel launch: %s", cudaGetErrorString(le));
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
	*out = b;                         // the caller frees it (free_batch) when any step below fails
	b->ctx = c;
	b->cap_pos = c->cfg.batch_bases & ~(kReadAlign - 1);
	const uint64_t padded = padded_positions(b->cap_pos);
	PreferNode on_gpu_node(c->numa_node);     // the ring's pages come from the GPU's own node (no-op on one-node hosts)
	CU(c, cudaMallocHost(&b->h_bases, padded / 32 * 8));
	CU(c, cudaMallocHost(&b->h_mask, padded / 32 * 4));
	CU(c, cudaMallocHost(&b->h_snap, 16));
	CU(c, cudaMalloc(&b->d_bases, padded / 32 * 8));
	CU(c, cudaMalloc(&b->d_mask, padded / 32 * 4));
	CU(c, cudaEventCreateWithFlags(&b->copied, cudaEventDisableTiming));
	CU(c, cudaEventCreateWithFlags(&b->done, cudaEventDisableTiming));
	return NTSM_OK;
}

// fold finished batches into the completed tallies; caller holds c->mu.  A device fault reported by a
// batch's event makes the ctx fail for good (async_error): the tallies of that batch are not taken.
static void reap(ntsm_ctx *c, bool block_on_oldest)
{
	while (!c->inflight.empty()) {
		ntsm_batch *b = c->inflight.front();
		cudaError_t q = cudaEventQuery(b->done);
		if (q == cudaErrorNotReady) {
			if (!block_on_oldest) break;
			q = cudaEventSynchronize(b->done);
			block_on_oldest = false;
		}
		if (q != cudaSuccess) {
			if (!c->async_error) fail(c, NTSM_ERR_CUDA, "a submitted batch failed on the device: %s", cudaGetErrorString(q));
			c->async_error = NTSM_ERR_CUDA;
		} else {
			c->done_kmers = b->h_snap[0];      // d_totals is cumulative and kernels run in submit order
			c->done_hits = b->h_snap[1];
			c->done_bases += b->n_bases;
		}
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
		if (c->async_error) return c->async_error;
		for (ntsm_batch *b : c->batches)
			if (b->state == 0) {
				b->state = 1;
				b->pk.reset_streaming(b->h_bases, b->h_mask);      // pinned, written once, read only by the DMA engine
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

// Enqueue one batch: H2D of a packed stream (src_* = host memory holding at least copy_pos positions)
// into the batch's device buffers -- or, when `ascii` is set, H2D of the reads' ASCII bytes and the
// decode + pack kernel (devpack.cuh) -- then the count kernel and the tally snapshot.  One producer
// at a time is in here (submit_mu): the order of the kernels on the compute stream is the order of
// the in-flight queue; c->mu is only taken for the queue itself, so producers waiting in
// ntsm_acquire_batch and pollers never wait behind CUDA calls.
struct AsciiJob {
	const void *src = nullptr;        // host bytes to copy (pinned)
	uint64_t bytes = 0;
	DevPackParams pp;                 // bases / mask / ascii pointers are filled in here
	bool fixed = true;
};

static int enqueue_batch(ntsm_ctx *c, ntsm_batch *b, const void *src_bases, const void *src_mask, uint64_t n_pos,
                         uint64_t copy_pos, uint64_t n_bases, const AsciiJob *ascii = nullptr)
{
	auto give_back = [&]() {
		std::lock_guard<std::mutex> g(c->mu);
		b->state = 0;
		c->cv.notify_all();
	};
	if (n_pos == 0) {
		give_back();
		return NTSM_OK;
	}
	auto enqueue = [&]() -> int {
		cudaEvent_t slot = nullptr;
		if (ascii) {                      // one feeder per ctx calls this: wait until the ASCII copy before the last one is through
			slot = c->ascii_ev[c->ascii_jobs++ & 1];
			CU(c, cudaEventSynchronize(slot));
		}
		std::lock_guard<std::mutex> sg(c->submit_mu);
		if (ascii) {
			CU(c, cudaMemcpyAsync(b->d_ascii, ascii->src, ascii->bytes, cudaMemcpyHostToDevice, c->copy_stream));
			CU(c, cudaEventRecord(slot, c->copy_stream));
			c->h2d_bytes += ascii->bytes;
			if (!ascii->fixed) {
				CU(c, cudaMemcpyAsync(b->d_aux, b->h_aux, (2 * (uint64_t)ascii->pp.n_reads + 2) * 4, cudaMemcpyHostToDevice, c->copy_stream));
				c->h2d_bytes += (2 * (uint64_t)ascii->pp.n_reads + 2) * 4;
			}
			CU(c, cudaEventRecord(b->copied, c->copy_stream));
			CU(c, cudaStreamWaitEvent(c->compute_stream, b->copied, 0));
			DevPackParams pp = ascii->pp;
			pp.ascii = b->d_ascii;
			pp.bases = b->d_bases;
			pp.mask = b->d_mask;
			const unsigned grid = (unsigned)std::min<uint64_t>((pp.n_chunks_out + 255) / 256, (uint64_t)c->sm_count * 8);
			if (ascii->fixed) pack_ascii_kernel<true><<<grid, 256, 0, c->compute_stream>>>(pp);
			else pack_ascii_kernel<false><<<grid, 256, 0, c->compute_stream>>>(pp);
			CU(c, cudaGetLastError());
			c->launches++;
		} else {
			CU(c, cudaMemcpyAsync(b->d_bases, src_bases, copy_pos / 32 * 8, cudaMemcpyHostToDevice, c->copy_stream));
			CU(c, cudaMemcpyAsync(b->d_mask, src_mask, copy_pos / 32 * 4, cudaMemcpyHostToDevice, c->copy_stream));
			c->h2d_bytes += copy_pos / 32 * 12;
			CU(c, cudaEventRecord(b->copied, c->copy_stream));
			CU(c, cudaStreamWaitEvent(c->compute_stream, b->copied, 0));
		}
		const int rc = launch_count(c, b->d_bases, b->d_mask, n_pos, c->compute_stream);
		if (rc) return rc;
		CU(c, cudaMemcpyAsync(b->h_snap, c->d_totals, 16, cudaMemcpyDeviceToHost, c->compute_stream));
		c->d2h_bytes += 16;
		CU(c, cudaEventRecord(b->done, c->compute_stream));
		std::lock_guard<std::mutex> g(c->mu);
		b->state = 2;
		b->n_bases = n_bases;
		b->n_pos = n_pos;
		c->last_batch = b;
		c->submitted_bases += n_bases;
		c->submitted_reads += ascii ? ascii->pp.n_reads : b->n_reads;
		c->inflight.push_back(b);
		c->cv.notify_all();
		return NTSM_OK;
	};
	const int rc = enqueue();
	if (rc) give_back();              // the buffer returns to the ring; the error stays with the ctx
	return rc;
}

// ASCII reads in PINNED host memory, decoded and packed on the device.  Fixed-length form: n_reads
// rows of `stride` bytes, read_len of them bases.  The caller (bulk.cpp's feeder thread) sizes n_reads
// with ntsm_ascii_capacity so that the packed result fits one batch.
uint64_t ntsm_ascii_capacity_fixed(const ntsm_ctx *c, uint64_t read_len, uint64_t stride)
{
	const uint64_t cap_pos = c->cfg.batch_bases & ~(kReadAlign - 1);
	const uint64_t by_pos = cap_pos / read_span(read_len);
	const uint64_t by_bytes = (cap_pos + 4096) / std::max<uint64_t>(1, stride);      // d_ascii holds cap_pos + 4096 bytes
	return std::min(by_pos, by_bytes);
}

static int ensure_ascii(ntsm_ctx *c, ntsm_batch *b, uint64_t n_reads_var)
{
	if (!b->d_ascii) CU(c, cudaMalloc(&b->d_ascii, b->cap_pos + 4096));
	if (n_reads_var + 1 > b->aux_cap) {
		cudaFree(b->d_aux);
		cudaFreeHost(b->h_aux);
		b->d_aux = b->h_aux = nullptr;
		b->aux_cap = 0;
		const uint64_t cap = std::max<uint64_t>(n_reads_var + 1, b->cap_pos / 64);
		CU(c, cudaMalloc(&b->d_aux, (2 * cap + 2) * 4));
		CU(c, cudaMallocHost(&b->h_aux, (2 * cap + 2) * 4));
		b->aux_cap = cap;
	}
	return NTSM_OK;
}

int ntsm_submit_ascii_fixed(ntsm_ctx *c, const char *rows, uint64_t read_len, uint64_t stride, uint64_t n_reads)
{
	if (!c || !rows || n_reads == 0) return NTSM_OK;
	if (n_reads > ntsm_ascii_capacity_fixed(c, read_len, stride)) return fail(c, NTSM_ERR_ARG, "ntsm_submit_ascii_fixed: block larger than a batch");
	ntsm_batch *b = nullptr;
	int rc = ntsm_acquire_batch(c, &b);
	if (rc) return rc;
	if ((rc = ensure_ascii(c, b, 0))) { ntsm_release_batch(c, b); return rc; }
	AsciiJob j;
	j.src = rows;
	j.bytes = (n_reads - 1) * stride + read_len;
	j.fixed = true;
	memset(&j.pp, 0, sizeof j.pp);
	j.pp.read_len = (uint32_t)read_len;
	j.pp.stride = (uint32_t)stride;
	j.pp.span = (uint32_t)read_span(read_len);
	j.pp.groups_per_read = j.pp.span / 8;
	j.pp.n_reads = (uint32_t)n_reads;
	j.pp.n_pos = n_reads * read_span(read_len);
	j.pp.n_chunks_out = padded_positions(j.pp.n_pos) / 32;
	return enqueue_batch(c, b, nullptr, nullptr, j.pp.n_pos, 0, n_reads * read_len, &j);
}

// Variable-length form: read r = buf[off[r], off[r+1]) for r in [0, n_reads); takes as many whole reads
// from the front as fit one batch (at least one unless the first read alone is too long: returns 0 reads
// taken, and the caller packs that read on the host, which splits long reads).  *taken = reads consumed.
int ntsm_submit_ascii_var(ntsm_ctx *c, const char *buf, const uint64_t *off, uint64_t n_reads, uint64_t *taken)
{
	*taken = 0;
	if (!c || !buf || !off || n_reads == 0) return NTSM_OK;
	const uint64_t cap_pos = c->cfg.batch_bases & ~(kReadAlign - 1);
	// how many reads fit: positions (spans) and staged bytes both bounded by the batch
	uint64_t n = 0, pos = 0;
	while (n < n_reads) {
		const uint64_t len = off[n + 1] - off[n];
		if (pos + read_span(len) > cap_pos || off[n + 1] - off[0] > cap_pos + 4096) break;
		pos += read_span(len);
		++n;
	}
	if (n == 0) return NTSM_OK;
	ntsm_batch *b = nullptr;
	int rc = ntsm_acquire_batch(c, &b);
	if (rc) return rc;
	if ((rc = ensure_ascii(c, b, n))) { ntsm_release_batch(c, b); return rc; }
	uint32_t *in_off = b->h_aux, *out_pos = b->h_aux + (n + 1);
	uint64_t p = 0;
	for (uint64_t r = 0; r < n; ++r) {
		in_off[r] = (uint32_t)(off[r] - off[0]);
		out_pos[r] = (uint32_t)p;
		p += read_span(off[r + 1] - off[r]);
	}
	in_off[n] = (uint32_t)(off[n] - off[0]);
	out_pos[n] = (uint32_t)p;
	AsciiJob j;
	j.src = buf + off[0];
	j.bytes = off[n] - off[0];
	j.fixed = false;
	memset(&j.pp, 0, sizeof j.pp);
	j.pp.in_off = b->d_aux;
	j.pp.out_pos = b->d_aux + (n + 1);
	j.pp.n_reads = (uint32_t)n;
	j.pp.n_pos = p;
	j.pp.n_chunks_out = padded_positions(p) / 32;
	*taken = n;
	if (j.bytes == 0) {              // nothing but empty reads: no bases, no windows
		ntsm_release_batch(c, b);
		return NTSM_OK;
	}
	return enqueue_batch(c, b, nullptr, nullptr, p, 0, off[n] - off[0], &j);
}

// 1 when `p` points into page-locked host memory the DMA engines can read directly (cudaMallocHost /
// cudaHostAlloc / cudaHostRegister), else 0
// Who packs a bulk of reads that lies in page-locked memory?  Returns 0 = the host packers, 1 = host packers and
// a device feeder per GPU together, 2 = the feeders alone.  Measured (bench.py e2e_ascii legs, Gbases/s per box):
//                              host packers   both    feeders alone
//   1 GPU,  16 cores per GPU       109         106         51-54     (PCIe: a device-packed base costs 1 B of it, a host-packed one 0.375)
//   2 GPUs, 12 cores per GPU       112         130          95
//   4 GPUs,  8 cores per GPU       112         154         176
//   8 GPUs,  4 cores per GPU       113         164-167     183-184   (host DMA: 183 GB/s is all the box gives; packers only take from it)
// The packers are bound by host DRAM (2.1 B of traffic per base) at ~110 Gbases/s per BOX however many GPUs there
// are; the feeders by PCIe per GPU and by what the host can DMA in total, which the library cannot see.  What it
// can see is how many packer threads the caller gives each GPU, and the four rows above say: 14 or more -- the
// packers alone saturate the host; 10 to 13 -- both; fewer -- the feeders alone (the few packers would only take
// DRAM bandwidth from the DMA engines).  Option "device_pack" forces 0 or 1 (= both); threads = 0 forces the feeders.
int ntsm_ctx_device_pack(const ntsm_ctx *c, uint32_t host_packers_per_ctx)
{
	if (c->opt_device_pack >= 0) return c->opt_device_pack ? 1 : 0;
	return host_packers_per_ctx >= 14 ? 0 : host_packers_per_ctx >= 10 ? 1 : 2;
}

int ntsm_host_is_pinned(const void *p)
{
	cudaPointerAttributes at;
	if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
		cudaGetLastError();
		return 0;
	}
	return at.type == cudaMemoryTypeHost;
}

extern "C" int ntsm_host_register(void *p, uint64_t bytes)
{
	const cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterPortable);
	if (e != cudaSuccess) { cudaGetLastError(); return fail(nullptr, NTSM_ERR_CUDA, "cudaHostRegister: %s", cudaGetErrorString(e)); }
	return NTSM_OK;
}
extern "C" int ntsm_host_unregister(void *p)
{
	const cudaError_t e = cudaHostUnregister(p);
	if (e != cudaSuccess) { cudaGetLastError(); return fail(nullptr, NTSM_ERR_CUDA, "cudaHostUnregister: %s", cudaGetErrorString(e)); }
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

// A packed batch that lies in page-locked memory the ctx does not own (a slot of the parser processes' shared
// mapping, procpipe.h), padded by its packer up to padded_positions(n_pos): copy + count it through one of the ctx's
// device buffers.  *out is the batch whose `copied` event tells when the slot may be reused (ntsm_batch_copy_done).
int ntsm_submit_foreign(ntsm_ctx *c, const void *bases, const void *mask, uint64_t n_pos, uint64_t n_bases, uint64_t n_reads,
                        ntsm_batch **out)
{
	*out = nullptr;
	if (!c || !bases || !mask) return NTSM_ERR_ARG;
	if (n_pos > (c->cfg.batch_bases & ~(kReadAlign - 1))) return fail(c, NTSM_ERR_ARG, "ntsm_submit_foreign: batch larger than the ctx's buffers");
	ntsm_batch *b = nullptr;
	int rc = ntsm_acquire_batch(c, &b);
	if (rc) return rc;
	b->n_reads = n_reads;
	rc = enqueue_batch(c, b, bases, mask, n_pos, padded_positions(n_pos), n_bases);
	if (rc == NTSM_OK && n_pos) *out = b;
	return rc;
}
// 1 once the host-to-device copy of the batch last submitted through `b` has finished, 0 while it has not, < 0 on a device fault
int ntsm_batch_copy_done(ntsm_batch *b)
{
	cudaSetDevice(b->ctx->device);
	const cudaError_t q = cudaEventQuery(b->copied);
	if (q == cudaSuccess) return 1;
	if (q == cudaErrorNotReady) return 0;
	return fail(b->ctx, NTSM_ERR_CUDA, "batch copy failed on the device: %s", cudaGetErrorString(q));
}
int ntsm_ctx_parser_procs(const ntsm_ctx *c) { return c->opt_parser_procs; }
uint32_t ntsm_ctx_k(const ntsm_ctx *c) { return c->cfg.k; }
// what the multi-sample matrix path (multi.cu) needs of a ctx: the exact table, the site lists, the stream
int ntsm_ctx_view_get(ntsm_ctx *c, ntsm_ctx_view *v)
{
	if (!c || !v) return NTSM_ERR_ARG;
	if (!c->d_table) return fail(c, NTSM_ERR_ARG, "no site table loaded");
	v->device = c->device;
	v->k = c->cfg.k;
	v->n_kmers = c->n_kmers;
	v->n_sites = c->n_sites;
	v->table_mask = c->table_cap - 1;
	v->d_table = c->d_table;
	v->d_allele_off = c->d_allele_off;
	v->stream = c->compute_stream;
	return NTSM_OK;
}
void ntsm_ctx_add_launches(ntsm_ctx *c, uint64_t n) { c->launches += n; }
void ntsm_ctx_add_pcie(ntsm_ctx *c, uint64_t h2d, uint64_t d2h) { c->h2d_bytes += h2d; c->d2h_bytes += d2h; }
void ntsm_ctx_set_error(ntsm_ctx *c, const char *text) { t_last_error = text; if (c) c->err = text; }

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
	{   // every submitted batch, then whatever else sits on the ctx's two streams -- not the whole device
		std::unique_lock<std::mutex> g(c->mu);
		while (!c->inflight.empty()) reap(c, true);
		c->cv.notify_all();
		if (c->async_error) return c->async_error;
	}
	CU(c, cudaStreamSynchronize(c->copy_stream));
	CU(c, cudaStreamSynchronize(c->compute_stream));
	return NTSM_OK;
}

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
	c->submitted_reads -= b->n_reads - (lo + 1);
	b->read_end.clear();
	return 1;
}

// ------------------------------------------------------------------ multi-GPU
// stdout is the counts file, so NCCL's version banner must not land there.  NCCL honours
// NCCL_DEBUG_FILE only for levels above VERSION (2.27: debug.cc), and GPU hosts commonly export
// NCCL_DEBUG=VERSION: raise that to WARN (which prints the same banner) so the file setting applies.
static void nccl_banner_to_stderr()
{
	// once per process, before the first NCCL call (setenv is not thread-safe and ntsm_comm_init is
	// meant to run on one thread per GPU at once); documented in the header as a side effect
	static std::once_flag once;
	std::call_once(once, [] {
		const char *lvl = getenv("NCCL_DEBUG");
		if (lvl && !strcasecmp(lvl, "VERSION")) setenv("NCCL_DEBUG", "WARN", 1);
		setenv("NCCL_DEBUG_FILE", "/dev/stderr", 0);
	});
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
		PeerCounts pc;
		memset(&pc, 0, sizeof pc);
		pc.counts[0] = c->d_counts;
		pc.n_ranks = 1;
		site_reduce_kernel<<<(S + 255) / 256, 256, 0, c->compute_stream>>>(pc, c->d_allele_off, S, r, r + S, r + 2 * (size_t)S,
		                                                                   r + 3 * (size_t)S, nullptr);
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

// ------------------------------------------------------------------ several GPUs of ONE process
// Combine + per-site reduce for the ctxs of one process (the CLI's --gpus N: one ctx per GPU) without
// NCCL: ctx 0's GPU reads every other ctx's private counts straight out of peer memory over NVLink
// inside the per-site reduce kernel (site_reduce_kernel: sum over GPUs first, max afterwards), so the
// all-reduce and the reduce that follows it are ONE kernel, nothing is staged, and no communicator has
// to be brought up (ncclCommInitRank was most of a short run's wall time in round 1).  Each ctx's
// stream records "my counts are final"; ctx 0's stream waits for all of them.  ctxs on the same
// device (tests on a one-GPU box) simply read each other's arrays.
extern "C" int ntsm_group_finalize(ntsm_ctx *const *ctxs, uint32_t n_ctx, uint32_t *max_ref, uint32_t *max_var,
                                   uint32_t *sum_ref, uint32_t *sum_var, uint64_t totals[3])
{
	if (!ctxs || n_ctx == 0 || n_ctx > (uint32_t)kMaxPeers) return fail(nullptr, NTSM_ERR_ARG, "ntsm_group_finalize: 1..%d ctxs", kMaxPeers);
	ntsm_ctx *c0 = ctxs[0];
	if (n_ctx == 1) return ntsm_finalize(c0, max_ref, max_var, sum_ref, sum_var, totals);
	for (uint32_t i = 0; i < n_ctx; ++i) {
		ntsm_ctx *c = ctxs[i];
		if (!c || !c->d_table || c->n_kmers != c0->n_kmers || c->n_sites != c0->n_sites || c->cfg.k != c0->cfg.k)
			return fail(c0, NTSM_ERR_ARG, "ntsm_group_finalize: ctx %u does not hold the same site table", i);
		if (c->reduced) return fail(c0, NTSM_ERR_ARG, "ntsm_group_finalize: ctx %u was already combined", i);
	}
	PeerCounts pc;
	PeerTotals pt;
	memset(&pc, 0, sizeof pc);
	memset(&pt, 0, sizeof pt);
	pc.n_ranks = pt.n_ranks = (int)n_ctx;
	for (uint32_t i = 0; i < n_ctx; ++i) {
		ntsm_ctx *c = ctxs[i];
		int rc = ntsm_flush(c);
		if (rc) return rc;
		CU(c, cudaSetDevice(c->device));
		{
			std::unique_lock<std::mutex> g(c->mu);
			while (!c->inflight.empty()) reap(c, true);
			c->cv.notify_all();
			if (c->async_error) return c->async_error;
		}
		set_u64_kernel<<<1, 1, 0, c->compute_stream>>>(c->d_totals + 2, (unsigned long long)c->submitted_bases);
		CU(c, cudaGetLastError());
		c->launches++;
		CU(c, cudaEventRecord(c->drained, c->compute_stream));
		pc.counts[i] = c->d_counts;
		pt.totals[i] = c->d_totals;
		c->reduced = true;
	}
	CU(c0, cudaSetDevice(c0->device));
	for (uint32_t i = 1; i < n_ctx; ++i) {
		if (ctxs[i]->device != c0->device) {
			int can = 0;
			CU(c0, cudaDeviceCanAccessPeer(&can, c0->device, ctxs[i]->device));
			if (!can) return fail(c0, NTSM_ERR_CUDA, "GPU %d cannot read GPU %d's memory (no peer access)", c0->device, ctxs[i]->device);
			const cudaError_t pe = cudaDeviceEnablePeerAccess(ctxs[i]->device, 0);
			if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) return fail(c0, NTSM_ERR_CUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(pe));
			cudaGetLastError();
		}
		CU(c0, cudaStreamWaitEvent(c0->compute_stream, ctxs[i]->drained, 0));
	}
	const uint32_t S = c0->n_sites;
	cudaStream_t st = c0->compute_stream;
	if (S) {
		uint32_t *r = c0->d_rows;
		// the k-mer-level sums replace ctx 0's private counts (ntsm_get_counts on ctx 0 then returns the combined array)
		site_reduce_kernel<<<(S + 255) / 256, 256, 0, st>>>(pc, c0->d_allele_off, S, r, r + S, r + 2 * (size_t)S, r + 3 * (size_t)S, c0->d_counts);
		CU(c0, cudaGetLastError());
		c0->launches++;
	}
	sum_totals_kernel<<<1, 32, 0, st>>>(pt, c0->d_totals);
	CU(c0, cudaGetLastError());
	c0->launches++;
	uint32_t *dst[4] = { max_ref, max_var, sum_ref, sum_var };
	for (int i = 0; i < 4 && S; ++i)
		if (dst[i]) CU(c0, cudaMemcpyAsync(dst[i], c0->d_rows + (size_t)i * S, (size_t)S * 4, cudaMemcpyDeviceToHost, st));
	unsigned long long t[3] = { 0, 0, 0 };
	CU(c0, cudaMemcpyAsync(t, c0->d_totals, sizeof t, cudaMemcpyDeviceToHost, st));
	CU(c0, cudaStreamSynchronize(st));
	if (totals) { totals[0] = t[0]; totals[1] = t[1]; totals[2] = t[2]; }
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
		if (c->async_error) return c->async_error;
	}
	rc = ntsm_reduce_async(c);
	if (rc) return rc;
	const uint32_t S = c->n_sites;
	uint32_t *dst[4] = { max_ref, max_var, sum_ref, sum_var };
	for (int i = 0; i < 4 && S; ++i)
		if (dst[i]) {
			CU(c, cudaMemcpyAsync(dst[i], c->d_rows + (size_t)i * S, (size_t)S * 4, cudaMemcpyDeviceToHost, c->compute_stream));
			c->d2h_bytes += (size_t)S * 4;
		}
	unsigned long long t[3] = { 0, 0, 0 };
	CU(c, cudaMemcpyAsync(t, c->d_totals, sizeof t, cudaMemcpyDeviceToHost, c->compute_stream));
	c->d2h_bytes += sizeof t;
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
	CU(c, cudaMemcpyAsync(counts, c->d_counts, (size_t)c->n_kmers * 4, cudaMemcpyDeviceToHost, c->compute_stream));
	CU(c, cudaStreamSynchronize(c->compute_stream));
	return NTSM_OK;
}

// A shard's k-mer-level result joins this context's: counts[i] += add[i], tallies += totals.  This is the
// exact merge: shards are summed per k-mer and the per-site max is taken afterwards by
// ntsm_finalize, where `ntsmEval --merge` sums per-site maxima (src/CompareCounts.hpp:648-657).
extern "C" int ntsm_add_counts(ntsm_ctx *c, const uint32_t *counts, const uint64_t totals[3])
{
	if (!c || !counts || !totals) return fail(c, NTSM_ERR_ARG, "ntsm_add_counts: null argument");
	if (!c->d_table) return fail(c, NTSM_ERR_ARG, "no site table loaded");
	if (c->reduced) return fail(c, NTSM_ERR_ARG, "counts were already combined; ntsm_reset_counts first");
	int rc = ntsm_flush(c);
	if (rc) return rc;
	if ((rc = ntsm_sync(c))) return rc;
	uint32_t *d_add = nullptr;
	const uint32_t n = c->n_kmers;
	if (n) {
		CU(c, cudaMalloc(&d_add, (size_t)n * 4));
		cudaError_t e = cudaMemcpyAsync(d_add, counts, (size_t)n * 4, cudaMemcpyHostToDevice, c->compute_stream);
		if (e == cudaSuccess) {
			add_counts_kernel<<<(n + 255) / 256, 256, 0, c->compute_stream>>>(c->d_counts, d_add, n);
			e = cudaGetLastError();
		}
		if (e == cudaSuccess) e = cudaStreamSynchronize(c->compute_stream);
		cudaFree(d_add);
		if (e != cudaSuccess) return fail(c, NTSM_ERR_CUDA, "ntsm_add_counts: %s", cudaGetErrorString(e));
		c->launches++;
		c->h2d_bytes += (size_t)n * 4;
	}
	add_totals_kernel<<<1, 1, 0, c->compute_stream>>>(c->d_totals, (unsigned long long)totals[0], (unsigned long long)totals[1]);
	CU(c, cudaGetLastError());
	CU(c, cudaStreamSynchronize(c->compute_stream));
	c->launches++;
	std::lock_guard<std::mutex> g(c->mu);
	c->done_kmers += totals[0];
	c->done_hits += totals[1];
	c->done_bases += totals[2];
	c->submitted_bases += totals[2];
	return NTSM_OK;
}

// the three tallies as they stand, without combining anything (ntsm_counts_save)
extern "C" int ntsm_get_totals(ntsm_ctx *c, uint64_t totals[3])
{
	if (!c || !totals) return NTSM_ERR_ARG;
	int rc = ntsm_flush(c);
	if (rc) return rc;
	if ((rc = ntsm_sync(c))) return rc;
	unsigned long long t[3] = { 0, 0, 0 };
	CU(c, cudaMemcpyAsync(t, c->d_totals, sizeof t, cudaMemcpyDeviceToHost, c->compute_stream));
	CU(c, cudaStreamSynchronize(c->compute_stream));
	totals[0] = t[0];
	totals[1] = t[1];
	totals[2] = c->reduced ? t[2] : c->submitted_bases;      // the base tally joins the device tallies when the result is combined
	return NTSM_OK;
}

// library-internal helpers (not part of the public header)
uint64_t ntsm_ctx_max_counts(const ntsm_ctx *c) { return c->cfg.max_counts; }
uint64_t ntsm_ctx_reads(const ntsm_ctx *c) { return c->submitted_reads; }
int ntsm_ctx_numa_node(const ntsm_ctx *c) { return c->numa_node; }
uint64_t ntsm_ctx_batch_bases(const ntsm_ctx *c) { return c->cfg.batch_bases; }
void ntsm_set_thread_error(const char *text) { t_last_error = text; }

extern "C" uint64_t ntsm_ctx_launches(const ntsm_ctx *c) { return c ? c->launches.load() : 0; }
extern "C" const char *ntsm_ctx_kernel_name(const ntsm_ctx *c)
{
	if (!c) return "";
	if (c->kernel == 0) return "count_kernel_generic";
	if (c->kernel == 2) return "count_kernel_wide<1024,1>";
	if (c->cfg.k != 19) return "count_kernel_pair<0,1024,2>";
	static const char *names[4] = { "count_kernel_pair<19,1024,1>", "count_kernel_pair<19,1024,2>", "count_kernel_pair<19,512,4>", "count_kernel_pair<19,256,8>" };
	return names[c->opt_shape];
}
extern "C" void ntsm_ctx_pcie_bytes(const ntsm_ctx *c, uint64_t *h2d, uint64_t *d2h)
{
	if (h2d) *h2d = c ? c->h2d_bytes.load() : 0;
	if (d2h) *d2h = c ? c->d2h_bytes.load() : 0;
}
extern "C" int ntsm_ctx_l2_window(const ntsm_ctx *c) { return c && c->l2_window ? (int)(c->l2_hit_ratio * 100.0f + 0.5f) : 0; }
extern "C" uint64_t ntsm_ctx_probe_bytes(const ntsm_ctx *c) { return c ? c->probe_bytes : 0; }
extern "C" uint32_t ntsm_ctx_filter_bits(const ntsm_ctx *c) { return c ? c->filter_bits : 0; }
extern "C" uint32_t ntsm_ctx_table_capacity(const ntsm_ctx *c) { return c ? c->table_cap : 0; }
