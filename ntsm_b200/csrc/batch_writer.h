// batch_writer.h -- one producer's view of the pinned batch rings: "append this read", with the
// acquire / full / submit / next-GPU steps hidden.  Shared by the file parser threads
// (pipeline.cpp) and the bulk in-memory producers (bulk.cpp).  It is the hand-off the dead
// ProdConKseqRunner meant to do with its bulk and recycle queues
// (vendor/ProdConKseqRunner.hpp:34-46,76-110): a producer fills a bulk, passes it on, and gets a
// recycled one back.
#pragma once
#include <atomic>
#include <string>

#include "../../include/ntsm_b200.h"

namespace ntsm {

struct BatchWriter {
	ntsm_ctx *const *ctxs;
	uint32_t n_ctx;
	std::atomic<uint64_t> *next_batch;      // shared round-robin ticket: batch i goes to ctxs[i % n_ctx]
	ntsm_batch *b = nullptr;
	ntsm_ctx *bctx = nullptr;
	int error = 0;
	std::string error_text;

	BatchWriter(ntsm_ctx *const *c, uint32_t n, std::atomic<uint64_t> *ticket) : ctxs(c), n_ctx(n), next_batch(ticket) {}

	bool fail(int code, ntsm_ctx *c)
	{
		error = code;
		error_text = ntsm_last_error(c);
		return false;
	}

	// submit the open batch (if any); `submitted` (nullable) receives the ctx it went to
	bool submit(ntsm_ctx **submitted = nullptr)
	{
		if (submitted) *submitted = nullptr;
		if (!b) return true;
		ntsm_ctx *c = bctx;
		const int rc = ntsm_submit_batch(c, b);
		b = nullptr;
		if (rc) return fail(rc, c);
		if (submitted) *submitted = c;
		return true;
	}

	// FingerPrint::insertCount(seq, len) (src/FingerPrint.hpp:89) for this producer.  on_submit is
	// called after every batch this read caused to be submitted (the -m check hooks in there).  If it
	// returns true the cap was reached by what is already submitted: a read that has not begun is
	// dropped (the reference stops before it, src/FingerPrint.hpp:67); a read longer than a batch that
	// is part-way in is finished first, because the reference counts a whole read before it looks at
	// the cap (processSingleRead, :473-487).
	template <class F> bool append(const char *seq, uint64_t len, F &&on_submit)
	{
		uint64_t pos = 0;
		bool finishing = false;
		for (;;) {
			if (!b) {
				bctx = ctxs[next_batch->fetch_add(1) % n_ctx];
				const int rc = ntsm_acquire_batch(bctx, &b);
				if (rc) { b = nullptr; return fail(rc, bctx); }
			}
			const int r = ntsm_batch_append(b, seq, len, &pos);
			if (r < 0) return fail(r, bctx);
			if (r == 1) return true;
			ntsm_ctx *went = nullptr;
			if (!submit(&went)) return false;
			if (!finishing && on_submit(went)) {
				if (pos == 0) return true;
				finishing = true;
			}
		}
	}
	bool append(const char *seq, uint64_t len)
	{
		return append(seq, len, [](ntsm_ctx *) { return false; });
	}
};

}  // namespace ntsm
