// pipeline.cpp -- FingerPrint::computeCounts (src/FingerPrint.hpp:46-87) as a host ingest pipeline:
// parser threads (one file each at a time) -> pinned packed batches -> round-robin over the GPUs.
// This is what the dead ProdConKseqRunner (vendor/ProdConKseqRunner.hpp:30-184) intended: file
// readers produce bulks, buffers are recycled; the "workers" are now CUDA streams.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <fcntl.h>
#include <spawn.h>
#include <stdio.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/syscall.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>

#include <atomic>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/ntsm_b200.h"
#include "batch_writer.h"
#include "fastx.h"
#include "internal.h"
#include "pack.h"
#include "procpipe.h"

extern char **environ;

namespace {

struct Shared {
	ntsm_ctx *const *ctxs;
	uint32_t n_ctx;
	const char *const *paths;
	uint32_t n_paths;
	int verbose;
	uint64_t max_counts;
	bool exact_cap = false;                     // one parser thread: the -m stop is trimmed to the deciding read
	uint32_t k = 19;
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

// ------------------------------------------------------------------------------------------------
// Parsers as worker processes (procpipe.h).  The shared mapping -- header, slot table, paths, batch slots -- is
// created once per geometry and stays page-locked for the life of the process; workers are spawned per call.
struct ProcShm {
	int fd = -1;
	uint8_t *base = nullptr;
	size_t bytes = 0;
	uint64_t cap_pos = 0;
	uint32_t n_slots = 0;
	bool registered = false;
};
std::mutex g_proc_mu;          // one ntsm_count_files at a time uses the worker mapping
ProcShm g_proc;

std::string worker_exe()
{
	Dl_info info;
	if (!dladdr((void *)&ntsm_count_files, &info) || !info.dli_fname) return "";
	std::string lib = info.dli_fname;                 // .../ntsm_b200/lib/libntsm_b200.so
	const size_t slash = lib.rfind('/');
	if (slash == std::string::npos) return "";
	std::string exe = lib.substr(0, slash) + "/../bin/ntsm_parse_worker";
	return access(exe.c_str(), X_OK) == 0 ? exe : "";
}

bool plain_regular_file(const char *path)
{
	struct stat sb;
	if (stat(path, &sb) != 0 || !S_ISREG(sb.st_mode) || sb.st_size == 0) return false;
	const int fd = open(path, O_RDONLY);
	if (fd < 0) return false;
	unsigned char m[2] = { 0, 0 };
	const bool gz = pread(fd, m, 2, 0) == 2 && m[0] == 0x1f && m[1] == 0x8b;
	close(fd);
	return !gz;
}

// makes (or reuses) the shared mapping for n_slots slots of cap_pos positions; page-locks the slot area
bool proc_shm_prepare(uint64_t cap_pos, uint32_t n_slots)
{
	using namespace ntsm;
	const ProcLayout L = proc_layout(cap_pos, n_slots);
	if (g_proc.base && (g_proc.cap_pos != cap_pos || g_proc.n_slots != n_slots)) {
		if (g_proc.registered) cudaHostUnregister(g_proc.base + ((ProcHeader *)g_proc.base)->data_off);
		munmap(g_proc.base, g_proc.bytes);
		close(g_proc.fd);
		g_proc = ProcShm();
	}
	if (!g_proc.base) {
		const int fd = (int)syscall(SYS_memfd_create, "ntsm_parse", 1u /* MFD_CLOEXEC: only the workers get it, through dup2 */);
		if (fd < 0) return false;
		if (ftruncate(fd, (off_t)L.total) != 0) { close(fd); return false; }
		void *m = mmap(nullptr, L.total, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
		if (m == MAP_FAILED) { close(fd); return false; }
		g_proc.fd = fd;
		g_proc.base = (uint8_t *)m;
		g_proc.bytes = L.total;
		g_proc.cap_pos = cap_pos;
		g_proc.n_slots = n_slots;
		// portable: every CUDA context of the process (all the GPUs) may DMA out of it
		g_proc.registered = cudaHostRegister(g_proc.base + L.data_off, L.slot_stride * n_slots, cudaHostRegisterPortable) == cudaSuccess;
		if (!g_proc.registered) {
			cudaGetLastError();
			munmap(g_proc.base, g_proc.bytes);
			close(g_proc.fd);
			g_proc = ProcShm();
			return false;
		}
	}
	return true;
}

// 1 = done through worker processes (sh.error set on failure), 0 = not applicable / could not start: use the threads
int count_files_procs(Shared &sh, uint32_t nt)
{
	using namespace ntsm;
	const int want = ntsm_ctx_parser_procs(sh.ctxs[0]);
	if (want == 0 || sh.max_counts != 0 || sh.n_paths == 0) return 0;
	if (want < 0 && nt <= 6) return 0;                     // few parsers: mapped in-process threads are as good
	for (uint32_t i = 0; i < sh.n_paths; ++i)
		if (!plain_regular_file(sh.paths[i])) return 0;    // gzip input is inflate-bound, pipes cannot be mapped: threads
	const std::string exe = worker_exe();
	if (exe.empty()) return 0;
	std::lock_guard<std::mutex> guard(g_proc_mu);
	const uint64_t cap_pos = ntsm_ctx_batch_bases(sh.ctxs[0]) & ~7ull;
	const uint32_t n_slots = 2 * nt + 2 * sh.n_ctx + 2;
	if (!proc_shm_prepare(cap_pos, n_slots)) return 0;
	const ProcLayout L = proc_layout(cap_pos, n_slots);
	if (!proc_init(g_proc.base, L, cap_pos, n_slots, sh.k, (uint32_t)sh.verbose, sh.paths, sh.n_paths, nt)) return 0;   // more path text than there is room for: threads
	ProcHeader *h = reinterpret_cast<ProcHeader *>(g_proc.base);

	// the workers get the mapping as descriptor kWorkerFd (dup2 drops the close-on-exec flag for them only)
	constexpr int kWorkerFd = 213;
	std::vector<pid_t> pids;
	posix_spawn_file_actions_t fa;
	if (posix_spawn_file_actions_init(&fa) != 0) return 0;
	posix_spawn_file_actions_adddup2(&fa, g_proc.fd, kWorkerFd);
	char fdbuf[16];
	snprintf(fdbuf, sizeof fdbuf, "%d", kWorkerFd);
	for (uint32_t w = 0; w < nt; ++w) {
		char wbuf[16];
		snprintf(wbuf, sizeof wbuf, "%u", w);
		char *argv[] = { const_cast<char *>(exe.c_str()), fdbuf, wbuf, nullptr };
		pid_t pid = 0;
		if (posix_spawn(&pid, exe.c_str(), &fa, nullptr, argv, environ) != 0) break;
		pids.push_back(pid);
	}
	posix_spawn_file_actions_destroy(&fa);
	if (pids.empty()) return 0;

	ProcSlot *slots = reinterpret_cast<ProcSlot *>(g_proc.base + h->slots_off);
	std::vector<ntsm_batch *> inflight(n_slots, nullptr);
	auto set_error = [&](int code, const std::string &text) {
		std::lock_guard<std::mutex> g(sh.err_mu);
		if (!sh.error.load()) { sh.error.store(code); sh.err_text = text; }
		h->stop.store(1, std::memory_order_release);
	};
	size_t alive = pids.size();
	uint64_t ticket = 0;
	for (;;) {
		bool progressed = false, busy = false;
		for (uint32_t i = 0; i < n_slots; ++i) {
			const uint32_t st = slots[i].state.load(std::memory_order_acquire);
			if (st == kSlotReady && !sh.error.load()) {
				ntsm_ctx *c = sh.ctxs[ticket++ % sh.n_ctx];
				const uint8_t *b = g_proc.base + h->data_off + (uint64_t)i * h->slot_stride;
				ntsm_batch *nb = nullptr;
				const int rc = ntsm_submit_foreign(c, b, b + h->bases_bytes, slots[i].n_pos, slots[i].n_bases, slots[i].n_reads, &nb);
				if (rc) {
					set_error(rc, ntsm_last_error(c));
					slots[i].state.store(kSlotFree, std::memory_order_release);
				} else {
					inflight[i] = nb;
					slots[i].state.store(nb ? kSlotInflight : kSlotFree, std::memory_order_release);
				}
				progressed = true;
			} else if (st == kSlotReady) {                 // an error elsewhere: drop what the workers still hand over
				slots[i].state.store(kSlotFree, std::memory_order_release);
			} else if (st == kSlotInflight) {
				const int d = ntsm_batch_copy_done(inflight[i]);
				if (d != 0) {
					if (d < 0) set_error(d, ntsm_last_error(sh.ctxs[0]));
					inflight[i] = nullptr;
					slots[i].state.store(kSlotFree, std::memory_order_release);
					progressed = true;
				} else busy = true;
			} else if (st == kSlotFilling) busy = true;
		}
		// reap workers that have left
		for (size_t w = 0; w < pids.size(); ++w) {
			if (pids[w] <= 0) continue;
			int status = 0;
			const pid_t r = waitpid(pids[w], &status, WNOHANG);
			if (r == pids[w]) {
				pids[w] = 0;
				--alive;
				if (!(WIFEXITED(status) && WEXITSTATUS(status) == 0) && !h->error.load()) {
					char msg[96];
					snprintf(msg, sizeof msg, "parser worker %zu ended abnormally (status 0x%x)", w, status);
					set_error(NTSM_ERR_IO, msg);
				}
				progressed = true;
			}
		}
		if (h->error.load(std::memory_order_acquire) && !sh.error.load()) set_error(h->error.load(), h->error_text);
		if (alive == 0) {
			// a worker that died while filling a slot leaves it in kSlotFilling: nothing will come of it
			bool pending = false;
			for (uint32_t i = 0; i < n_slots; ++i) {
				const uint32_t st = slots[i].state.load(std::memory_order_acquire);
				if (st == kSlotReady || st == kSlotInflight) pending = true;
			}
			if (!pending) break;
		}
		if (!progressed) {
			(void)busy;
			struct timespec ts = { 0, 20000 };
			nanosleep(&ts, nullptr);
		}
	}
	return 1;
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
	sh.k = ntsm_ctx_k(ctxs[0]);
	// many parsers over plain files: worker PROCESSES (procpipe.h); else -- or if they cannot be started -- threads
	if (count_files_procs(sh, nt) == 0) {
		std::vector<std::thread> pool;
		for (uint32_t t = 1; t < nt; ++t) pool.emplace_back(worker, std::ref(sh), t);
		worker(sh, 0);
		for (auto &t : pool) t.join();
	}
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
