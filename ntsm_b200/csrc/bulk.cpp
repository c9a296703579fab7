// bulk.cpp -- FingerPrint::insertCount (src/FingerPrint.hpp:89-103) for a whole bulk of reads that
// is already in host memory: what a consumer of the ProdConKseqRunner bulk queue would receive
// (vendor/ProdConKseqRunner.hpp:34-46 hands kseq_t records over 256 at a time).  `threads`
// producers each take contiguous ranges of reads, decode + pack them (pack.cpp) straight into
// pinned batches and submit those to the GPUs round-robin; the host never touches a base again.
// When the caller's buffer is page-locked, one FEEDER thread per GPU works the same queue of read
// blocks: it hands a block's ASCII bytes to the DMA engine as they are and the GPU decodes + packs
// them (devpack.cuh) -- those bases are never touched by a host core at all.  Packers and feeders
// balance themselves: whoever finishes a block takes the next one (the packers cost the host ~2.1
// bytes of DRAM traffic per base but only 0.375 bytes of PCIe, the feeders 1 byte of each).
#include <algorithm>
#include <atomic>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/ntsm_b200.h"
#include "batch_writer.h"
#include "internal.h"
#include "numa.h"

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

// Feeder: blocks of reads go to ctx `ci` as ASCII (see above).  A block is cut into as many ASCII
// batches as it needs; a read too long for one batch goes through the host packer, which splits it.
void feeder(Bulk &bk, uint32_t ci)
{
	ntsm_ctx *c = bk.ctxs[ci];
	ntsm::run_on_node(ntsm_ctx_numa_node(c));           // the feeder only issues DMA, but its pinned aux tables are touched here
	ntsm::BatchWriter bw(bk.ctxs, bk.n_ctx, &bk.next_batch);
	auto fail = [&](int code, const char *text) {
		std::lock_guard<std::mutex> g(bk.err_mu);
		if (!bk.error.load()) { bk.error.store(code); bk.err_text = text ? text : ""; }
	};
	const uint64_t rows_per_batch = bk.off ? 0 : ntsm_ascii_capacity_fixed(c, bk.read_len, bk.stride);
	for (;;) {
		const uint64_t blk = bk.next_block.fetch_add(1);
		const uint64_t r0 = blk * bk.reads_per_block;
		if (r0 >= bk.n_reads || bk.error.load()) break;
		const uint64_t r1 = std::min(bk.n_reads, r0 + bk.reads_per_block);
		uint64_t r = r0;
		while (r < r1) {
			int rc;
			if (!bk.off) {
				if (rows_per_batch == 0) {      // a row longer than a batch: host packer (splits reads)
					if (!bw.append(bk.buf + r * bk.stride, bk.read_len)) { fail(bw.error, bw.error_text.c_str()); return; }
					++r;
					continue;
				}
				const uint64_t n = std::min(rows_per_batch, r1 - r);
				rc = ntsm_submit_ascii_fixed(c, bk.buf + r * bk.stride, bk.read_len, bk.stride, n);
				r += n;
			} else {
				uint64_t taken = 0;
				rc = ntsm_submit_ascii_var(c, bk.buf, bk.off + r, r1 - r, &taken);
				if (rc == 0 && taken == 0) {
					if (!bw.append(bk.buf + bk.off[r], bk.off[r + 1] - bk.off[r])) { fail(bw.error, bw.error_text.c_str()); return; }
					taken = 1;
				}
				r += taken;
			}
			if (rc) { fail(rc, ntsm_last_error(c)); return; }
		}
		if (!bw.submit()) { fail(bw.error, bw.error_text.c_str()); return; }
	}
}

void producer(Bulk &bk)
{
	// one GPU: the packers run on its node; several GPUs: batches go round-robin, any node is as good as another
	if (bk.n_ctx == 1) ntsm::run_on_node(ntsm_ctx_numa_node(bk.ctxs[0]));
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
	// page-locked input: the GPUs can decode + pack it themselves (one feeder thread per GPU) -- instead of the host
	// packers, next to them, or not at all: ntsm_ctx_device_pack (ctx.cu) holds the rule and the measurements behind it
	const int mode = ntsm_host_is_pinned(bk.buf) ? ntsm_ctx_device_pack(bk.ctxs[0], threads / bk.n_ctx) : 0;
	const bool pinned = mode != 0;
	uint32_t nt = mode == 2 ? 0 : (threads ? threads : (pinned ? 0 : 1));
	if (nt > n_blocks) nt = (uint32_t)n_blocks;
	std::vector<std::thread> pool;
	if (pinned)
		for (uint32_t g = 0; g < bk.n_ctx; ++g) pool.emplace_back(feeder, std::ref(bk), g);
	// every producer runs on a thread of its own (they may pin themselves to the GPU's NUMA node; the
	// caller's thread keeps its affinity)
	for (uint32_t t = 0; t < nt; ++t) pool.emplace_back(producer, std::ref(bk));
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
