// hostpath_bench.cpp -- the host half of FingerPrint::computeCounts without a GPU: FastxReader -> Packer into a
// reused buffer, one thread per file, to see what parse + pack cost per base.  Measurement tool.
//   g++ -O3 -std=c++17 -pthread -I ntsm_b200/csrc tools/hostpath_bench.cpp ntsm_b200/csrc/{fastx,gzsource,inflate,pargz,pack}.cpp -lz -o /tmp/hostpath_bench
//   /tmp/hostpath_bench [parse|pack|packN] file...  (parse = reader only, pack = reader + packer, packN = with N helper threads per reader)
#include <stdio.h>
#include <string.h>
#include <sys/wait.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <thread>
#include <vector>

#include "fastx.h"
#include "pack.h"

int main(int argc, char **argv)
{
	if (argc < 3) return 1;
	const bool do_pack = !strncmp(argv[1], "pack", 4);
	const int helpers = argv[1][0] && argv[1][strlen(argv[1]) - 1] >= '1' && argv[1][strlen(argv[1]) - 1] <= '9' ? argv[1][strlen(argv[1]) - 1] - '0' : 0;   // "pack1": one helper thread per reader
	const int nf = argc - 2;
	std::atomic<uint64_t> bases{0}, reads{0};
	const auto t0 = std::chrono::steady_clock::now();
	if (!strncmp(argv[1], "procs", 5)) {
		// one PROCESS per file instead of one thread: separate address spaces, so page-table work for mapped
		// files does not meet on one mm (is that what stops the mapped source from scaling in one process?)
		for (int i = 0; i < nf; ++i)
			if (fork() == 0) {
				ntsm::FastxReader rd;
				if (!rd.open(argv[2 + i], 0)) _exit(1);
				const uint64_t cap = 1ull << 24;
				std::vector<uint64_t> b(ntsm::padded_positions(cap) / 32 + 64);
				std::vector<uint32_t> m(ntsm::padded_positions(cap) / 32 + 64);
				ntsm::Packer pk;
				pk.reset(b.data(), m.data());
				int64_t l;
				while ((l = rd.next()) >= 0) {
					if (pk.pos + ntsm::read_span((uint64_t)l) > cap) pk.reset(b.data(), m.data());
					if ((uint64_t)l < cap / 2) pk.put_read(rd.seq(), (uint64_t)l);
				}
				_exit(0);
			}
		while (wait(nullptr) > 0) {}
		const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
		printf("procs: %d file(s)/process(es) in %.3f s\n", nf, dt);
		return 0;
	}
	std::vector<std::thread> th;
	for (int i = 0; i < nf; ++i)
		th.emplace_back([&, i] {
			ntsm::FastxReader rd;
			if (!rd.open(argv[2 + i], helpers)) { fprintf(stderr, "cannot open %s\n", argv[2 + i]); return; }
			const uint64_t cap = 1ull << 24;
			std::vector<uint64_t> b(ntsm::padded_positions(cap) / 32 + 64);
			std::vector<uint32_t> m(ntsm::padded_positions(cap) / 32 + 64);
			ntsm::Packer pk;
			pk.reset(b.data(), m.data());
			uint64_t nb = 0, nr = 0;
			int64_t l;
			while ((l = rd.next()) >= 0) {
				if (do_pack) {
					if (pk.pos + ntsm::read_span((uint64_t)l) > cap) pk.reset(b.data(), m.data());
					if ((uint64_t)l < cap / 2) pk.put_read(rd.seq(), (uint64_t)l);
				}
				nb += (uint64_t)l;
				++nr;
			}
			bases += nb;
			reads += nr;
		});
	for (auto &t : th) t.join();
	const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	printf("%s: %d file(s)/thread(s), %llu reads, %.3f Gbases in %.3f s -> %.2f Gbases/s (%.2f per thread)\n", argv[1], nf,
	       (unsigned long long)reads.load(), bases.load() / 1e9, dt, bases.load() / dt / 1e9, bases.load() / dt / 1e9 / nf);
	return 0;
}
