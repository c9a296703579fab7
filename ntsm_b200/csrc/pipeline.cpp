// pipeline.cpp -- FingerPrint::computeCounts (src/FingerPrint.hpp:46-87) as a host ingest pipeline:
// parser threads (one file each at a time) -> pinned packed batches -> round-robin over the GPUs.
// This is what the dead ProdConKseqRunner (vendor/ProdConKseqRunner.hpp:30-184) intended: file
// readers produce bulks, buffers are recycled; the "workers" are now CUDA streams.
#include <stdio.h>

#include <atomic>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/ntsm_b200.h"
#include "batch_writer.h"
#include "fastx.h"
#include "internal.h"

namespace {

struct Shared {
	ntsm_ctx *const *ctxs;
	uint32_t n_ctx;
	const char *const *paths;
	uint32_t n_paths;
	int verbose;
	uint64_t max_counts;
	bool exact_cap = false;                     // one parser thread: the -m stop is trimmed to the deciding read
	bool map_plain = true;                      // few parser threads: plain files are scanned in place through a mapping (gzsource.cpp)
	uint32_t helpers = 0, helpers_extra = 0;    // idle -t threads lent to each parser for block-parallel inflate
	std::atomic<uint32_t> next_file{0};
	std::atomic<uint64_t> next_batch{0};
	std::atomic<bool> early{false};
	std::atomic<int> error{0};
	std::mutex err_mu;
	std::string err_text;
};

// -m check (FingerPrint.hpp:476-487) at batch granularity: sum the hit tallies of all GPUs over
// the batches that have completed.
void check_cap(Shared &sh)
{
	if (!sh.max_counts || sh.early.load()) return;
	uint64_t hits = 0;
	for (uint32_t i = 0; i < sh.n_ctx; ++i) {
		uint64_t h = 0;
		ntsm_poll_totals(sh.ctxs[i], nullptr, &h, nullptr, nullptr);
		hits += h;
	}
	if (hits > sh.max_counts) sh.early.store(true);
}

void worker(Shared &sh, uint32_t wi)
{
	const int helpers = (int)(sh.helpers + (wi < sh.helpers_extra ? 1 : 0));
	ntsm::FastxReader rd;
	ntsm::BatchWriter bw(sh.ctxs, sh.n_ctx, &sh.next_batch);
	auto set_error = [&](int code, const std::string &text) {
		std::lock_guard<std::mutex> g(sh.err_mu);
		if (!sh.error.load()) { sh.error.store(code); sh.err_text = text; }
	};
	// -m: wait for the batch just submitted so the stop decision does not depend on timing
	// (deterministic for one parser thread, like the reference's -t 1)
	auto after_submit = [&](ntsm_ctx *went) -> bool {
		if (!sh.max_counts) return false;
		if (!went || sh.early.load()) return sh.early.load();
		ntsm_sync(went);
		uint64_t hits = 0, own = 0;
		for (uint32_t i = 0; i < sh.n_ctx; ++i) {
			uint64_t h = 0;
			ntsm_poll_totals(sh.ctxs[i], nullptr, &h, nullptr, nullptr);
			hits += h;
			if (sh.ctxs[i] == went) own = h;
		}
		if (hits <= sh.max_counts) return false;
		// FingerPrint.hpp:476-487 checks after every read: with one parser thread (the reference's
		// deterministic -t 1) the batch is cut back to the read that crossed the cap
		if (sh.exact_cap) {
			const int rc = ntsm_trim_to_cap(went, hits - own, sh.max_counts);
			if (rc < 0) set_error(rc, ntsm_last_error(went));
		}
		sh.early.store(true);
		return true;
	};
	for (;;) {
		const uint32_t fi = sh.next_file.fetch_add(1);
		if (fi >= sh.n_paths || sh.error.load()) break;
		if (!rd.open(sh.paths[fi], helpers, sh.map_plain)) {                               // FingerPrint.hpp:51-57
			set_error(NTSM_ERR_IO, std::string("file ") + sh.paths[fi] + " cannot be opened");
			break;
		}
		if (sh.verbose) fprintf(stderr, "Opening %s\n", sh.paths[fi]);   // :58-62
		int64_t l;
		while (!sh.early.load() && !sh.error.load() && (l = rd.next()) >= 0) {   // :67 (any negative code ends the file)
			if (!bw.append(rd.seq(), (uint64_t)l, after_submit)) { set_error(bw.error, bw.error_text); return; }
		}
		rd.close();
	}
	ntsm_ctx *went = nullptr;
	if (!bw.submit(&went)) { set_error(bw.error, bw.error_text); return; }
	after_submit(went);
}

}  // namespace

extern "C" int ntsm_count_files(ntsm_ctx *const *ctxs, uint32_t n_ctx, const char *const *paths, uint32_t n_paths,
                                uint32_t threads, int verbose, int *early)
{
	if (!ctxs || !n_ctx || (!paths && n_paths)) return NTSM_ERR_ARG;
	Shared sh;
	sh.ctxs = ctxs;
	sh.n_ctx = n_ctx;
	sh.paths = paths;
	sh.n_paths = n_paths;
	sh.verbose = verbose;
	sh.max_counts = 0;
	sh.max_counts = ntsm_ctx_max_counts(ctxs[0]);   // every ctx carries the same cap
	uint32_t nt = threads ? threads : 1;
	if (nt > n_paths) nt = n_paths ? n_paths : 1;                   // the reference never uses more than #files (:47-48) ...
	const uint32_t spare = (threads ? threads : 1) - nt;            // ... the rest of -t inflates BGZF blocks for the parsers (gzsource.h)
	sh.exact_cap = nt == 1;
	sh.map_plain = nt <= 6;
	sh.helpers = spare / nt;
	sh.helpers_extra = spare % nt;
	std::vector<std::thread> pool;
	for (uint32_t t = 1; t < nt; ++t) pool.emplace_back(worker, std::ref(sh), t);
	worker(sh, 0);
	for (auto &t : pool) t.join();
	for (uint32_t i = 0; i < n_ctx; ++i) {
		const int rc = ntsm_sync(ctxs[i]);
		if (rc && !sh.error.load()) { sh.error.store(rc); sh.err_text = ntsm_last_error(ctxs[i]); }
	}
	check_cap(sh);
	if (early) *early = sh.early.load() ? 1 : 0;
	if (sh.error.load()) {
		ntsm_set_thread_error(sh.err_text.c_str());
		return sh.error.load();
	}
	return NTSM_OK;
}
