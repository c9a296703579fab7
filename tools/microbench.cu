// tools/microbench.cu -- measurements that drive the kernel design (not product code).
//   1. random 4-byte probes into a table of T bytes (global/L2/HBM)  -> probes/s vs T
//   2. random probes into shared memory                                -> probes/s
//   3. the count kernel on a random packed stream for several pre-filter sizes
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ntsm_b200/bin/microbench tools/microbench.cu -Lntsm_b200/lib -lntsm_b200
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../include/ntsm_b200.h"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__global__ void probe_global(const uint32_t *__restrict__ t, uint32_t mask, int iters, uint32_t *out)
{
	uint32_t x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
	uint32_t acc = 0;
	for (int i = 0; i < iters; i += 8) {
		uint32_t v[8];
#pragma unroll
		for (int j = 0; j < 8; ++j) { x = x * 1664525u + 1013904223u; v[j] = __ldg(t + ((x >> 4) & mask)); }
#pragma unroll
		for (int j = 0; j < 8; ++j) acc ^= v[j];
	}
	if (acc == 0x12345678u) out[0] = acc;
}

// sector-local: 32 lanes probe the same 32-byte sector groups (n_sectors distinct sectors per warp instruction)
__global__ void probe_global_grouped(const uint32_t *__restrict__ t, uint32_t mask, int iters, int group, uint32_t *out)
{
	const int lane = threadIdx.x & 31;
	uint32_t x = ((blockIdx.x * blockDim.x + threadIdx.x) / group) * 2654435761u + 12345u;   // lanes in a group share x
	uint32_t acc = 0;
	for (int i = 0; i < iters; i += 8) {
		uint32_t v[8];
#pragma unroll
		for (int j = 0; j < 8; ++j) { x = x * 1664525u + 1013904223u; v[j] = __ldg(t + ((((x >> 4) & mask) & ~7u) | (lane & 7))); }
#pragma unroll
		for (int j = 0; j < 8; ++j) acc ^= v[j];
	}
	if (acc == 0x12345678u) out[0] = acc;
}

// line-local: groups of g lanes probe g DIFFERENT 32-byte sectors of the same 128-byte line (32 sectors but only
// 32/g lines per warp instruction) -> is the L1->L2 wall counted in sectors or in line requests?
__global__ void probe_global_line(const uint32_t *__restrict__ t, uint32_t mask, int iters, int group, uint32_t *out)
{
	const int lane = threadIdx.x & 31;
	uint32_t x = ((blockIdx.x * blockDim.x + threadIdx.x) / group) * 2654435761u + 12345u;
	uint32_t acc = 0;
	for (int i = 0; i < iters; i += 8) {
		uint32_t v[8];
#pragma unroll
		for (int j = 0; j < 8; ++j) {
			x = x * 1664525u + 1013904223u;
			v[j] = __ldg(t + ((((x >> 4) & mask) & ~31u) | ((uint32_t)(lane % group) << 3) | ((x >> 1) & 7u)));
		}
#pragma unroll
		for (int j = 0; j < 8; ++j) acc ^= v[j];
	}
	if (acc == 0x12345678u) out[0] = acc;
}

__global__ void probe_shared(int iters, uint32_t *out, int words)
{
	extern __shared__ uint32_t s[];
	for (int i = threadIdx.x; i < words; i += blockDim.x) s[i] = i * 2654435761u;
	__syncthreads();
	uint32_t x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
	uint32_t acc = 0;
	const uint32_t mask = words - 1;
	for (int i = 0; i < iters; i += 8) {
#pragma unroll
		for (int j = 0; j < 8; ++j) { x = x * 1664525u + 1013904223u; acc ^= s[(x >> 4) & mask]; }
	}
	if (acc == 0x12345678u) out[0] = acc;
}

// 4. random 4-byte probes into the distributed shared memory of a thread-block cluster: every CTA
//    holds `words` words, a probe picks (rank, word) at random -> is a cluster-wide bitmap in
//    DSMEM cheaper per probe than one in L2?
__global__ void probe_dsmem(int iters, uint32_t *out, uint32_t words, uint32_t cluster_size)
{
	extern __shared__ uint32_t s[];
	for (uint32_t i = threadIdx.x; i < words; i += blockDim.x) s[i] = i * 2654435761u;
	asm volatile("barrier.cluster.arrive.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
	const uint32_t s_base = (uint32_t)__cvta_generic_to_shared(s);
	uint32_t x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
	uint32_t acc = 0;
	for (int i = 0; i < iters; i += 8) {
		uint32_t v[8];
#pragma unroll
		for (int j = 0; j < 8; ++j) {
			x = x * 1664525u + 1013904223u;
			const uint32_t rank = (x >> 24) % cluster_size;
			const uint32_t local = s_base + 4 * (uint32_t)(((uint64_t)((x >> 4) & 0x0FFFFFFFu) * words) >> 28);
			uint32_t remote;
			asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(rank));
			asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v[j]) : "r"(remote));
		}
#pragma unroll
		for (int j = 0; j < 8; ++j) acc ^= v[j];
	}
	if (acc == 0x12345678u) out[0] = acc;
	asm volatile("barrier.cluster.arrive.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms; cudaEventElapsedTime(&ms, a, b); return ms; }

int main(int argc, char **argv)
{
	int which = argc > 1 ? atoi(argv[1]) : 7;
	cudaEvent_t e0, e1;
	CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
	uint32_t *d_out; CK(cudaMalloc(&d_out, 64));
	cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
	printf("device %s, %d SMs, L2 %d MB, clock %d kHz\n", prop.name, prop.multiProcessorCount, prop.l2CacheSize >> 20, prop.clockRate);

	if (which & 1) {
		const int iters = 512, threads = 256, blocks = prop.multiProcessorCount * 8;
		for (size_t mb : {1, 4, 8, 16, 32, 64, 96, 128, 256, 1024}) {
			size_t bytes = mb << 20;
			uint32_t *t; CK(cudaMalloc(&t, bytes)); CK(cudaMemset(t, 1, bytes));
			uint32_t mask = (uint32_t)(bytes / 4 - 1);
			if ((bytes / 4) & (bytes / 4 - 1)) { // not pow2: use largest pow2 below
				size_t w = 1; while (w * 2 <= bytes / 4) w *= 2; mask = (uint32_t)(w - 1);
			}
			for (int rep = 0; rep < 2; ++rep) {
				CK(cudaEventRecord(e0));
				probe_global<<<blocks, threads>>>(t, mask, iters, d_out);
				CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
			}
			double n = (double)blocks * threads * iters;
			printf("global random probe  T=%5zu MB : %7.1f Gprobe/s\n", mb, n / time_ms(e0, e1) / 1e6);
			for (int group : {4, 8}) {
				CK(cudaEventRecord(e0));
				probe_global_grouped<<<blocks, threads>>>(t, mask, iters, group, d_out);
				CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
				printf("   grouped x%d (32/%d sectors per warp-load)  : %7.1f Gprobe/s\n", group, group, n / time_ms(e0, e1) / 1e6);
			}
			CK(cudaFree(t));
		}
	}
	if (which & 2) {
		const int iters = 4096, threads = 256;
		for (int kb : {32, 64, 128, 200}) {
			int words = 1; while (words * 2 * 4 <= kb * 1024) words *= 2;
			CK(cudaFuncSetAttribute(probe_shared, cudaFuncAttributeMaxDynamicSharedMemorySize, words * 4));
			int per_sm = (227 * 1024) / (words * 4); if (per_sm > 8) per_sm = 8; if (per_sm < 1) per_sm = 1;
			int blocks = prop.multiProcessorCount * per_sm;
			for (int rep = 0; rep < 2; ++rep) {
				CK(cudaEventRecord(e0));
				probe_shared<<<blocks, threads, words * 4>>>(iters, d_out, words);
				CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
			}
			CK(cudaGetLastError());
			printf("shared random probe %4d KB x%d CTA/SM : %7.1f Gprobe/s\n", words * 4 / 1024, per_sm, (double)blocks * threads * iters / time_ms(e0, e1) / 1e6);
		}
	}
	if (which & 16) {
		const int iters = 512, threads = 256, blocks = prop.multiProcessorCount * 8;
		for (size_t mb : {32, 64}) {
			size_t bytes = mb << 20;
			uint32_t *t; CK(cudaMalloc(&t, bytes)); CK(cudaMemset(t, 1, bytes));
			const uint32_t mask = (uint32_t)(bytes / 4 - 1);
			const double n = (double)blocks * threads * iters;
			for (int group : {1, 2, 4}) {
				for (int rep = 0; rep < 2; ++rep) {
					CK(cudaEventRecord(e0));
					probe_global_line<<<blocks, threads>>>(t, mask, iters, group, d_out);
					CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
				}
				printf("global line-grouped probe T=%3zu MB, %d sectors of one line per %d lanes (%2d lines, 32 sectors per warp-load): %7.1f Gsector/s\n",
				       mb, group, group, 32 / group, n / time_ms(e0, e1) / 1e6);
			}
			CK(cudaFree(t));
		}
	}
	if (which & 8) {
		const int iters = 2048;
		CK(cudaFuncSetAttribute(probe_dsmem, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
		for (int kb : {64, 200}) {
			const uint32_t words = kb * 256;
			CK(cudaFuncSetAttribute(probe_dsmem, cudaFuncAttributeMaxDynamicSharedMemorySize, words * 4));
			for (int threads : {256, 1024}) for (int cs : {1, 2, 4, 8, 16}) {
				cudaLaunchConfig_t lc = {};
				cudaLaunchAttribute at[1];
				at[0].id = cudaLaunchAttributeClusterDimension;
				at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
				lc.attrs = at; lc.numAttrs = 1;
				lc.blockDim = dim3(threads); lc.dynamicSmemBytes = words * 4; lc.stream = 0;
				int nclusters = 0;
				lc.gridDim = dim3(cs);
				if (cudaOccupancyMaxActiveClusters(&nclusters, probe_dsmem, &lc) != cudaSuccess || nclusters == 0) {
					printf("dsmem probe %3d KB/CTA cluster %2d x %4d threads: not launchable (%s)\n", kb, cs, threads, cudaGetErrorString(cudaGetLastError()));
					continue;
				}
				const int blocks = nclusters * cs;
				lc.gridDim = dim3(blocks);
				cudaError_t le = cudaSuccess;
				for (int rep = 0; rep < 2 && le == cudaSuccess; ++rep) {
					CK(cudaEventRecord(e0));
					le = cudaLaunchKernelEx(&lc, probe_dsmem, iters, d_out, words, (uint32_t)cs);
					CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
				}
				if (le != cudaSuccess) { printf("dsmem launch failed: %s\n", cudaGetErrorString(le)); cudaGetLastError(); continue; }
				printf("dsmem random probe %3d KB/CTA cluster %2d x %4d threads (%3d CTAs): %7.1f Gprobe/s\n", kb, cs, threads, blocks,
				       (double)blocks * threads * iters / time_ms(e0, e1) / 1e6);
			}
		}
	}
	if (which & 4) {
		// count kernel on a random stream: 2^30 positions, panel-sized random table
		const uint32_t n_kmers = 1270317, n_sites = 96287;
		std::vector<uint64_t> hashes(n_kmers);
		uint64_t x = 88172645463325252ull;
		for (uint32_t i = 0; i < n_kmers; ++i) hashes[i] = ntsm_hash64((uint64_t)i * 977 + 5, 19);   // distinct: hash64 is a bijection
		std::vector<uint32_t> off(2 * n_sites + 1);
		for (uint32_t i = 0; i <= 2 * n_sites; ++i) off[i] = (uint32_t)((uint64_t)i * n_kmers / (2 * n_sites));
		const uint64_t n_pos = 1ull << 30;
		const uint64_t padded = ntsm_padded_positions(n_pos);
		uint32_t *d_b, *d_m;
		CK(cudaMalloc(&d_b, padded / 16 * 4)); CK(cudaMalloc(&d_m, padded / 32 * 4));
		{
			std::vector<uint32_t> hb(padded / 16), hm(padded / 32, 0);
			for (auto &w : hb) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; w = (uint32_t)x; }
			for (uint64_t i = n_pos / 32; i < padded / 32; ++i) hm[i] = 0xFFFFFFFFu;
			for (uint64_t p = 150; p < n_pos; p += 151) hm[p / 32] |= 1u << (p % 32);    // read separators
			CK(cudaMemcpy(d_b, hb.data(), hb.size() * 4, cudaMemcpyHostToDevice));
			CK(cudaMemcpy(d_m, hm.data(), hm.size() * 4, cudaMemcpyHostToDevice));
		}
		// (round 1 swept its kernel generations here through NTSM_KERNEL; the library now takes options)
		for (const char *cfgs : {"0:26", "1:26", "1:27", "2:26"}) {
			char kv[2] = { cfgs[0], 0 };
			const char *fb = cfgs + 2;
			ntsm_ctx *ctx; ntsm_cfg cfg; memset(&cfg, 0, sizeof cfg); cfg.k = 19;
			if (ntsm_ctx_create(&ctx, &cfg)) { printf("ctx: %s\n", ntsm_last_error(NULL)); return 1; }
			ntsm_ctx_set_option(ctx, "kernel", atoi(kv));          // 0 generic, 1 paired seeds, 2 wide paired seeds
			ntsm_ctx_set_option(ctx, "filter_bits", atoi(fb));
			if (ntsm_load_sites(ctx, hashes.data(), NULL, n_kmers, off.data(), n_sites)) { printf("load: %s\n", ntsm_last_error(ctx)); return 1; }
			float best = 1e9;
			for (int rep = 0; rep < 4; ++rep) {
				CK(cudaEventRecord(e0));
				ntsm_count_packed_device(ctx, d_b, d_m, n_pos, n_pos, NULL);
				// ctx compute stream is non-blocking: synchronize via ntsm_sync and time with wall events on that stream instead
				ntsm_sync(ctx);
				CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
				if (rep && time_ms(e0, e1) < best) best = time_ms(e0, e1);
			}
			uint64_t tot[3];
			ntsm_finalize(ctx, NULL, NULL, NULL, NULL, tot);
			printf("count kernel variant %s filter 2^%s bits: %8.3f ms per 2^30 positions = %7.1f Gpos/s  (TK=%llu hits=%llu)\n", kv, fb, best,
			       (double)n_pos / best / 1e6, (unsigned long long)tot[0], (unsigned long long)tot[1]);
			ntsm_ctx_destroy(ctx);
		}
	}
	return 0;
}
