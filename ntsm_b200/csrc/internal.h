// internal.h -- helpers shared between the library's translation units (not exported in the header)
#pragma once
#include <stdint.h>

#include "../../include/ntsm_b200.h"

uint64_t ntsm_ctx_max_counts(const ntsm_ctx *c);
uint64_t ntsm_ctx_reads(const ntsm_ctx *c);      // reads submitted so far (after an exact -m stop: reads counted)
uint64_t ntsm_ctx_batch_bases(const ntsm_ctx *c);
void ntsm_set_thread_error(const char *text);
// exact -m stop after the batch just completed on ctx (ctx.cu); hits_elsewhere = hits on the other GPUs
int ntsm_trim_to_cap(ntsm_ctx *c, uint64_t hits_elsewhere, uint64_t cap);
// reads that arrive as ASCII in pinned host memory and are decoded + packed on the device (ctx.cu, devpack.cuh)
int ntsm_ctx_numa_node(const ntsm_ctx *c);       // NUMA node of the ctx's GPU, -1 = unknown / one-node host
int ntsm_host_is_pinned(const void *p);
int ntsm_ctx_device_pack(const ntsm_ctx *c, uint32_t host_packers_per_ctx);
uint64_t ntsm_ascii_capacity_fixed(const ntsm_ctx *c, uint64_t read_len, uint64_t stride);   // rows one batch takes
int ntsm_submit_ascii_fixed(ntsm_ctx *c, const char *rows, uint64_t read_len, uint64_t stride, uint64_t n_reads);
int ntsm_submit_ascii_var(ntsm_ctx *c, const char *buf, const uint64_t *off, uint64_t n_reads, uint64_t *taken);
// parser worker processes (pipeline.cpp, procpipe.h): a packed batch in foreign page-locked memory -> copy + count
int ntsm_submit_foreign(ntsm_ctx *c, const void *bases, const void *mask, uint64_t n_pos, uint64_t n_bases, uint64_t n_reads,
                        ntsm_batch **out);
int ntsm_batch_copy_done(ntsm_batch *b);
int ntsm_ctx_parser_procs(const ntsm_ctx *c);
uint32_t ntsm_ctx_k(const ntsm_ctx *c);
// the multi-sample matrix path (multi.cu) works on a ctx's exact table and site lists
struct ntsm_ctx_view {
	int device;
	uint32_t k, n_kmers, n_sites, table_mask;
	const void *d_table;              // ntsm::TableSlot[table_mask + 1]
	const uint32_t *d_allele_off;     // [2 * n_sites + 1]
	void *stream;                     // the ctx's compute stream (cudaStream_t)
};
int ntsm_ctx_view_get(ntsm_ctx *c, ntsm_ctx_view *v);
void ntsm_ctx_add_launches(ntsm_ctx *c, uint64_t n);
void ntsm_ctx_add_pcie(ntsm_ctx *c, uint64_t h2d, uint64_t d2h);
void ntsm_ctx_set_error(ntsm_ctx *c, const char *text);
// ntsm_multi_insert_windows with the genotypes already packed 2 bits per sample (16 per uint32, (n_samples + 15) / 16 words per line)
int ntsm_multi_insert_windows_packed(ntsm_multi *m, const char *windows, uint32_t wstride, const uint16_t *lens, const uint32_t *geno2,
                                     uint32_t n_lines, uint32_t multi);
// printNormMatrix's numbers in blocks of rows (multi.cu): kernels once, rows fetched as the caller formats
int ntsm_multi_norm_begin(ntsm_multi *m, uint64_t *first_undef);
int ntsm_multi_norm_fetch(ntsm_multi *m, uint32_t row0, uint32_t n_rows, double *values, double *sums);
void ntsm_multi_norm_end(ntsm_multi *m);
