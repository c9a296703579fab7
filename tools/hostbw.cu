// hostbw.cu -- what the GPU box's host can feed: DRAM read / copy / non-temporal-store bandwidth over T
// threads, and pinned host -> device DMA bandwidth over G GPUs at once (per NUMA placement when there is
// more than one node).  Measurement tool for DESIGN 5 (host ingest), not part of the library.
//   nvcc -O2 -o hostbw hostbw.cu -lpthread ; ./hostbw [threads] [gib_per_thread_x16]
#include <cuda_runtime.h>
#include <immintrin.h>
#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <chrono>
#include <thread>
#include <vector>

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static void pin_to(int cpu)
{
	cpu_set_t s; CPU_ZERO(&s); CPU_SET(cpu, &s); sched_setaffinity(0, sizeof s, &s);
}

int main(int argc, char **argv)
{
	int ncpu = std::thread::hardware_concurrency();
	int T = argc > 1 ? atoi(argv[1]) : ncpu;
	size_t per = (size_t)(argc > 2 ? atoi(argv[2]) : 8) << 26;   // bytes per thread (default 512 MiB)
	int G = 0; cudaGetDeviceCount(&G);
	printf("cpus %d threads %d bytes/thread %zu gpus %d\n", ncpu, T, per, G);
	std::vector<char *> src(T), dst(T);
	{
		std::vector<std::thread> th;
		for (int t = 0; t < T; ++t) th.emplace_back([&, t] { pin_to(t % ncpu); src[t] = (char *)aligned_alloc(4096, per); dst[t] = (char *)aligned_alloc(4096, per); memset(src[t], t + 1, per); memset(dst[t], 0, per); });
		for (auto &x : th) x.join();
	}
	auto run = [&](const char *name, int nt, double bytes_per_byte, void (*fn)(char *, char *, size_t)) {
		double best = 1e9;
		for (int rep = 0; rep < 3; ++rep) {
			std::atomic<int> ready{0};
			std::atomic<bool> go{false};
			std::vector<std::thread> th;
			std::vector<double> t1(nt);
			for (int t = 0; t < nt; ++t) th.emplace_back([&, t] { pin_to(t % ncpu); ready++; while (!go.load()) {} fn(src[t], dst[t], per); t1[t] = now(); });
			while (ready.load() < nt) {}
			double t0 = now(); go.store(true);
			for (auto &x : th) x.join();
			double e = 0; for (double v : t1) e = v > e ? v : e;
			best = e - t0 < best ? e - t0 : best;
		}
		printf("%-28s threads %3d  %.1f GB/s (traffic model %.2fx = %.1f GB/s)\n", name, nt, nt * (double)per / best / 1e9, bytes_per_byte, nt * (double)per * bytes_per_byte / best / 1e9);
	};
	auto rd = [](char *s, char *, size_t n) { __m256i a = _mm256_setzero_si256(); for (size_t i = 0; i < n; i += 32) a = _mm256_xor_si256(a, _mm256_load_si256((const __m256i *)(s + i))); volatile long long sink = _mm256_extract_epi64(a, 0); (void)sink; };
	auto cp = [](char *s, char *d, size_t n) { memcpy(d, s, n); };
	auto nt_cp = [](char *s, char *d, size_t n) { for (size_t i = 0; i < n; i += 32) _mm256_stream_si256((__m256i *)(d + i), _mm256_load_si256((const __m256i *)(s + i))); _mm_sfence(); };
	for (int nt : {1, 4, 8, 16, 32, 64}) { if (nt > T) break; run("read", nt, 1, rd); }
	for (int nt : {1, 8, 16, 32, 64}) { if (nt > T) break; run("memcpy (rd+rfo+wr)", nt, 3, cp); }
	for (int nt : {1, 8, 16, 32, 64}) { if (nt > T) break; run("nt copy (rd+wr)", nt, 2, nt_cp); }

	// pinned H2D: G GPUs at once, 1 GiB each, repeated
	if (G) {
		size_t hb = (size_t)1 << 30;
		std::vector<char *> hp(G), dp(G);
		std::vector<cudaStream_t> st(G);
		for (int g = 0; g < G; ++g) { cudaSetDevice(g); cudaMallocHost(&hp[g], hb); memset(hp[g], 1, hb); cudaMalloc(&dp[g], hb); cudaStreamCreate(&st[g]); }
		for (int ng : {1, 2, 4, 8}) {
			if (ng > G) break;
			for (int dir = 0; dir < 2; ++dir) {
				for (int g = 0; g < ng; ++g) { cudaSetDevice(g); cudaDeviceSynchronize(); }
				double t0 = now();
				for (int rep = 0; rep < 4; ++rep)
					for (int g = 0; g < ng; ++g) { cudaSetDevice(g); if (dir == 0) cudaMemcpyAsync(dp[g], hp[g], hb, cudaMemcpyHostToDevice, st[g]); else cudaMemcpyAsync(hp[g], dp[g], hb, cudaMemcpyDeviceToHost, st[g]); }
				for (int g = 0; g < ng; ++g) { cudaSetDevice(g); cudaStreamSynchronize(st[g]); }
				double dt = now() - t0;
				printf("%s pinned, %d GPU(s) at once: %.1f GB/s total (%.1f per GPU)\n", dir ? "D2H" : "H2D", ng, 4.0 * ng * hb / dt / 1e9, 4.0 * hb / dt / 1e9);
			}
		}
		// H2D while T host threads stream-read (the packer's situation)
		for (int ng : {1, G}) {
			std::atomic<bool> stop{false};
			std::vector<std::thread> th;
			std::atomic<long long> bytes{0};
			for (int t = 0; t < T; ++t) th.emplace_back([&, t] { pin_to(t % ncpu); while (!stop.load()) { rd(src[t], dst[t], per); bytes += (long long)per; } });
			double t0 = now();
			for (int rep = 0; rep < 4; ++rep)
				for (int g = 0; g < ng; ++g) { cudaSetDevice(g); cudaMemcpyAsync(dp[g], hp[g], hb, cudaMemcpyHostToDevice, st[g]); }
			for (int g = 0; g < ng; ++g) { cudaSetDevice(g); cudaStreamSynchronize(st[g]); }
			double dt = now() - t0;
			stop.store(true);
			for (auto &x : th) x.join();
			double dt2 = now() - t0;
			printf("H2D %d GPU(s) with %d reader threads: %.1f GB/s DMA, host reads %.1f GB/s meanwhile\n", ng, T, 4.0 * ng * hb / dt / 1e9, bytes.load() / dt2 / 1e9);
			if (G == 1) break;
		}
		// cudaHostRegister cost
		{
			size_t rb = (size_t)2 << 30;
			char *p = (char *)aligned_alloc(4096, rb); memset(p, 1, rb);
			double t0 = now(); cudaError_t e = cudaHostRegister(p, rb, cudaHostRegisterDefault); double dt = now() - t0;
			printf("cudaHostRegister 2 GiB touched: %s %.3f s (%.1f GB/s)\n", cudaGetErrorString(e), dt, rb / dt / 1e9);
			t0 = now(); cudaHostUnregister(p); printf("unregister %.3f s\n", now() - t0);
		}
	}
	return 0;
}
