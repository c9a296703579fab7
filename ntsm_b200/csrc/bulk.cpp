// bulk.cpp -- FingerPrint::insertCount (src/FingerPrint.hpp:89-103) for a whole bulk of reads that
// is already in host memory: what a consumer of the ProdConKseqRunner bulk queue would receive
// (vendor/ProdConKseqRunner.hpp:34-46 hands kseq_t records over 256 at a time).  `threads`
// producers each take contiguous ranges of reads, decode + pack them (pack.cpp) straight into
// pinned batches and submit those to the GPUs round-robin; the host never touches a base again.
#include <algorithm>
#include <atomic>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/ntsm_b200.h"
#include "batch_writer.h"
#include "internal.h"

namespace {

struct Bulk {
	ntsm_ctx *const *ctxs;
	uint32_t n_ctx;
	const char *buf;
	const uint64_t *off;     // variable-length reads: read r = buf[off[r], off[r+1]); NULL for the matrix form
	uint64_t read_len, stride;
	uint64_t n_reads;
	uint64_t reads_per_block;
	std::atomic<uint64_t> next_block{0};
	std::atomic<uint64_t> next_batch{0};
	std::atomic<int> error{0};
	std::mutex err_mu;
	std::string err_text;
};

void producer(Bulk &bk)
{
	ntsm::BatchWriter bw(bk.ctxs, bk.n_ctx, &bk.next_batch);
	for (;;) {
		const uint64_t blk = bk.next_block.fetch_add(1);
		const uint64_t r0 = blk * bk.reads_per_block;
		if (r0 >= bk.n_reads || bk.error.load()) break;
		const uint64_t r1 = std::min(bk.n_reads, r0 + bk.reads_per_block);
		bool ok = true;
		if (bk.off) {
			for (uint64_t r = r0; r < r1 && ok; ++r) ok = bw.append(bk.buf + bk.off[r], bk.off[r + 1] - bk.off[r]);
		} else {
			for (uint64_t r = r0; r < r1 && ok; ++r) ok = bw.append(bk.buf + r * bk.stride, bk.read_len);
		}
		// one block = one batch worth of reads: submit now so the GPU starts while the next block packs
		if (ok) ok = bw.submit();
		if (!ok) {
			std::lock_guard<std::mutex> g(bk.err_mu);
			if (!bk.error.load()) { bk.error.store(bw.error); bk.err_text = bw.error_text; }
			return;
		}
	}
}

int run(Bulk &bk, uint32_t threads, uint64_t total_bases)
{
	if (bk.n_reads == 0) return NTSM_OK;
	// size blocks so that a block fills about one batch
	const uint64_t cap = ntsm_ctx_batch_bases(bk.ctxs[0]);
	const uint64_t avg = total_bases / bk.n_reads + 8;               // a read spans its bases + separator, rounded up to 8 positions
	bk.reads_per_block = std::max<uint64_t>(1, (cap - cap / 64) / avg);
	const uint64_t n_blocks = (bk.n_reads + bk.reads_per_block - 1) / bk.reads_per_block;
	uint32_t nt = threads ? threads : 1;
	if (nt > n_blocks) nt = (uint32_t)n_blocks;
	std::vector<std::thread> pool;
	for (uint32_t t = 1; t < nt; ++t) pool.emplace_back(producer, std::ref(bk));
	producer(bk);
	for (auto &t : pool) t.join();
	if (bk.error.load()) {
		ntsm_set_thread_error(bk.err_text.c_str());
		return bk.error.load();
	}
	return NTSM_OK;
}

}  // namespace

extern "C" int ntsm_insert_reads(ntsm_ctx *const *ctxs, uint32_t n_ctx, const char *buf, const uint64_t *off,
                                 uint64_t n_reads, uint32_t threads)
{
	if (!ctxs || !n_ctx || (n_reads && (!buf || !off))) return NTSM_ERR_ARG;
	Bulk bk;
	bk.ctxs = ctxs; bk.n_ctx = n_ctx; bk.buf = buf; bk.off = off; bk.read_len = bk.stride = 0; bk.n_reads = n_reads;
	return run(bk, threads, n_reads ? off[n_reads] - off[0] : 0);
}

extern "C" int ntsm_insert_reads_fixed(ntsm_ctx *const *ctxs, uint32_t n_ctx, const char *buf, uint64_t read_len,
                                       uint64_t stride, uint64_t n_reads, uint32_t threads)
{
	if (!ctxs || !n_ctx || (n_reads && !buf) || stride < read_len) return NTSM_ERR_ARG;
	Bulk bk;
	bk.ctxs = ctxs; bk.n_ctx = n_ctx; bk.buf = buf; bk.off = nullptr; bk.read_len = read_len; bk.stride = stride; bk.n_reads = n_reads;
	return run(bk, threads, n_reads * read_len);
}
