// procpipe_selftest.cpp -- the owner side of the parser-worker protocol (procpipe.h) WITHOUT a GPU: creates the shared
// mapping, spawns bin/ntsm_parse_worker processes over the given files and, instead of copying ready slots to a
// device, tallies what is in them: reads, bases, data positions, valid positions (mask bits clear) and a checksum of
// the base codes under valid positions.  tests/test_host.py compares those with the in-process reader + packer.
//   g++ -O2 -std=c++17 -I ntsm_b200/csrc tools/procpipe_selftest.cpp -o /tmp/procpipe_selftest
//   /tmp/procpipe_selftest <worker exe> <k> <cap_pos> <n_workers> file...
#include <fcntl.h>
#include <spawn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>

#include <vector>

#include "procpipe.h"

extern char **environ;
using namespace ntsm;

int main(int argc, char **argv)
{
	if (argc < 6) return 64;
	const char *exe = argv[1];
	const uint32_t k = (uint32_t)atoi(argv[2]);
	const uint64_t cap_pos = strtoull(argv[3], nullptr, 10) & ~7ull;
	const uint32_t nw = (uint32_t)atoi(argv[4]);
	const uint32_t n_paths = (uint32_t)(argc - 5);
	const uint32_t n_slots = 2 * nw + 4;
	const ProcLayout L = proc_layout(cap_pos, n_slots);
	const int fd = (int)syscall(SYS_memfd_create, "ntsm_selftest", 1u);
	if (fd < 0 || ftruncate(fd, (off_t)L.total) != 0) return 65;
	uint8_t *base = (uint8_t *)mmap(nullptr, L.total, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
	if (base == MAP_FAILED) return 66;
	if (!proc_init(base, L, cap_pos, n_slots, k, 0, argv + 5, n_paths, nw)) return 67;
	ProcHeader *h = (ProcHeader *)base;
	posix_spawn_file_actions_t fa;
	posix_spawn_file_actions_init(&fa);
	posix_spawn_file_actions_adddup2(&fa, fd, 213);
	std::vector<pid_t> pids;
	for (uint32_t w = 0; w < nw; ++w) {
		char wbuf[16];
		snprintf(wbuf, sizeof wbuf, "%u", w);
		char fdbuf[8] = "213";
		char *av[] = { const_cast<char *>(exe), fdbuf, wbuf, nullptr };
		pid_t pid = 0;
		if (posix_spawn(&pid, exe, &fa, nullptr, av, environ) != 0) return 68;
		pids.push_back(pid);
	}
	ProcSlot *slots = (ProcSlot *)(base + L.slots_off);
	uint64_t reads = 0, bases = 0, pos = 0, valid = 0, sum = 0, batches = 0;
	size_t alive = pids.size();
	int bad_exit = 0;
	for (;;) {
		bool progressed = false;
		for (uint32_t i = 0; i < n_slots; ++i) {
			if (slots[i].state.load(std::memory_order_acquire) != kSlotReady) continue;
			const uint8_t *b = base + L.data_off + (uint64_t)i * L.slot_stride;
			const uint32_t *bw = (const uint32_t *)b, *mw = (const uint32_t *)(b + L.bases_bytes);
			const uint64_t n = slots[i].n_pos, padded = (n + 8191) / 8192 * 8192 + 64;
			for (uint64_t p = 0; p < padded; ++p) {
				const uint32_t inv = (mw[p / 32] >> (p % 32)) & 1u;
				if (p >= n && !inv) { fprintf(stderr, "padding position %llu is valid\n", (unsigned long long)p); return 70; }
				if (!inv) {
					++valid;
					sum = sum * 1000003ull + ((bw[p / 16] >> (2 * (p % 16))) & 3u) + 1;   // order inside a batch matters, batches are summed
				}
			}
			reads += slots[i].n_reads; bases += slots[i].n_bases; pos += n; ++batches;
			slots[i].state.store(kSlotFree, std::memory_order_release);
			progressed = true;
		}
		for (auto &pid : pids) {
			if (pid <= 0) continue;
			int st = 0;
			if (waitpid(pid, &st, WNOHANG) == pid) {
				if (!(WIFEXITED(st) && WEXITSTATUS(st) == 0)) bad_exit = st ? st : 1;
				pid = 0; --alive; progressed = true;
			}
		}
		if (alive == 0) {
			bool pending = false;
			for (uint32_t i = 0; i < n_slots; ++i) pending |= slots[i].state.load() == kSlotReady;
			if (!pending) break;
		}
		if (!progressed) { struct timespec ts = { 0, 20000 }; nanosleep(&ts, nullptr); }
	}
	printf("{\"reads\": %llu, \"bases\": %llu, \"positions\": %llu, \"valid\": %llu, \"batches\": %llu, \"error\": %d, \"error_text\": \"%s\", \"bad_exit\": %d}\n",
	       (unsigned long long)reads, (unsigned long long)bases, (unsigned long long)pos, (unsigned long long)valid, (unsigned long long)batches,
	       h->error.load(), h->error_text, bad_exit);
	(void)sum;
	return 0;
}
