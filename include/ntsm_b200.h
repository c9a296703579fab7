/*
 * ntsm_b200.h -- C ABI of libntsm_b200.so: the B200-native counting hot path of ntsmCount.
 *
 * The reference (JustinChu/ntsm @663f9a5) has no FFI: the seam this library sits behind is
 * the five-call sequence of src/ntSeqMatchCount.cpp:177-181
 *     FingerPrint fp; fp.computeCounts(files); fp.printOptionalHeader();
 *     fp.printCountsMax(); fp.printInfoSummary();
 * plus the per-read consumer contract FingerPrint::insertCount(seq,len)
 * (src/FingerPrint.hpp:89-103; `T::consume(kseq_t&)` in vendor/ProdConKseqRunner.hpp:93).
 * Every entry point below names the reference member it replaces.  INTEGRATION.md shows
 * the binding a maintainer of the reference would add.
 *
 * Conventions: plain C types, opaque handles, caller-owned outputs, no exceptions and no
 * exit() across the boundary.  Every function that can fail returns 0 (NTSM_OK) or a negative
 * NTSM_ERR_* code; ntsm_last_error() gives the text.  There is NO CPU fallback: without a
 * CUDA device ntsm_ctx_create fails with NTSM_ERR_CUDA.
 *
 * Packed batch layout ("2-bit + N-mask"), the unit streamed to the GPU:
 *   A batch is ONE stream of positions.  Position p holds a base code in bits
 *   [2*(p%16), 2*(p%16)+2) of the little-endian uint32 word bases2[p/16]
 *   (A=0 C=1 G=2 T/U=3, decoded with the reference table vendor/KseqHashIterator.hpp:114-127)
 *   and an invalid flag in bit (p%32) of nmask[p/32] (1 = the byte decoded to 4, i.e. "N").
 *   Reads are laid end to end with at least one invalid position after each read, so no k-mer
 *   window can span two reads.  The library's own packers start every read at a multiple of 8
 *   positions (a read of n bases occupies (n + 8) & ~7 positions: bases, separator, padding) so
 *   that both planes are byte-granular per read; streams handed to ntsm_count_packed_* only
 *   need the one separator.  read_off[r] is the position of read r's first base
 *   (read_off[n_reads] = end).  The tail is padded with invalid positions up to
 *   ntsm_padded_positions(n_pos).
 */
#ifndef NTSM_B200_H
#define NTSM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NTSM_OK 0
#define NTSM_ERR_ARG (-1)    /* bad argument / bad state                                  */
#define NTSM_ERR_CUDA (-2)   /* CUDA runtime failure (incl. no device)                    */
#define NTSM_ERR_NCCL (-3)   /* NCCL failure                                              */
#define NTSM_ERR_IO (-5)     /* file cannot be opened / read                              */
#define NTSM_ERR_NOKEY (-134) /* the reference would abort here with std::out_of_range:
                                a duplicate k-mer erased from the table (no -d) is still in a
                                site list, or the last site has no var record
                                (src/FingerPrint.hpp:276,282)                             */

typedef struct ntsm_ctx ntsm_ctx;     /* one per GPU: table, counts, streams, pinned buffers */
typedef struct ntsm_batch ntsm_batch; /* one pinned packed batch, owned by its ctx           */
typedef struct ntsm_sites ntsm_sites; /* host result of FingerPrint::initCountsHash          */
typedef struct ntsm_reader ntsm_reader; /* FASTA/FASTQ(.gz) record reader (kseq semantics)   */

typedef struct ntsm_cfg {
	uint32_t k;           /* opt::k (src/Options.h:23); 1..31                              */
	int32_t device;       /* CUDA device ordinal                                           */
	uint32_t n_buffers;   /* pinned+device batch buffers, >=2 (0 = default 3)              */
	uint32_t reserved;
	uint64_t batch_bases; /* capacity of one batch in stream positions (0 = default 2^25)  */
	uint64_t max_counts;  /* m_maxCounts (src/FingerPrint.hpp:41-43); 0 = no -m cap        */
} ntsm_cfg;

const char *ntsm_version(void);
int ntsm_device_count(void); /* visible CUDA devices; 0 when there is none (nothing can be counted then) */
int ntsm_device_warmup(int device); /* create the device's primary context now (thread-safe; lets a caller overlap it with reading sites.fa) */
/* text of the last error on this ctx, or (ctx == NULL) of the calling thread */
const char *ntsm_last_error(const ntsm_ctx *ctx);

/* ---------------- host arithmetic shared with the reference ---------------- */
/* vendor/KseqHashIterator.hpp:114-127 */
uint32_t ntsm_nt4(uint8_t c);
/* vendor/KseqHashIterator.hpp:129-139 (mask = 4^k - 1) and its inverse (the hash is a bijection) */
uint64_t ntsm_hash64(uint64_t key, uint32_t k);
uint64_t ntsm_hash64_inv(uint64_t hash, uint32_t k);

/* ---------------- site set: FingerPrint::initCountsHash (src/FingerPrint.hpp:490-564) ------ */
/* Reads the interleaved ref/var FASTA (plain or gz), first occurrence of a k-mer wins, later
 * occurrences produce the reference's warning text and, unless allow_dupes (-d), are erased
 * from the table while staying in the first owner's list. */
int ntsm_sites_load(ntsm_sites **out, const char *path, uint32_t k, int allow_dupes);
void ntsm_sites_free(ntsm_sites *s);
uint32_t ntsm_sites_k(const ntsm_sites *s);
uint32_t ntsm_sites_n_sites(const ntsm_sites *s);        /* m_alleleIDs.size()                 */
uint32_t ntsm_sites_n_kmers(const ntsm_sites *s);        /* listed k-mers = dense index space  */
uint64_t ntsm_sites_table_size(const ntsm_sites *s);     /* m_counts.size() after dupe removal */
const uint64_t *ntsm_sites_hashes(const ntsm_sites *s);  /* [n_kmers] hash64, list order: ref_0,var_0,ref_1,... */
const uint32_t *ntsm_sites_allele_off(const ntsm_sites *s); /* [2*n_sites+1] CSR into the above   */
const uint8_t *ntsm_sites_erased(const ntsm_sites *s);   /* [n_kmers] 1 = erased duplicate      */
const char *ntsm_sites_name(const ntsm_sites *s, uint32_t i); /* m_alleleIDs[i]                 */
uint32_t ntsm_sites_n_warnings(const ntsm_sites *s);
const char *ntsm_sites_warning(const ntsm_sites *s, uint32_t i); /* "Warning: <name> of REF file has a k-mer collision at pos: <p>" */
/* 0 if printCountsMax would complete, NTSM_ERR_NOKEY if the reference would abort */
int ntsm_sites_printable(const ntsm_sites *s);
/* m_maxCounts for -m <cov>: uint64((double)table_size * cov / 2); cov <= 0 -> 0 (off) */
uint64_t ntsm_sites_max_counts(const ntsm_sites *s, double cov);

/* ---------------- device context: FingerPrint object ---------------- */
int ntsm_ctx_create(ntsm_ctx **out, const ntsm_cfg *cfg);          /* FingerPrint() :35-44 */
void ntsm_ctx_destroy(ntsm_ctx *ctx);
/* Builds the device table (open addressing, keyed by the reference hash64 value) and the
 * orientation-free k-mer pre-filter; zeroes counts.  kmer_hash[i] is the hash of dense k-mer i;
 * erased (nullable) marks entries that are listed but not in the table. */
int ntsm_load_sites(ntsm_ctx *ctx, const uint64_t *kmer_hash, const uint8_t *erased, uint32_t n_kmers,
                    const uint32_t *allele_off, uint32_t n_sites);  /* initCountsHash :490-564 */
int ntsm_load_siteset(ntsm_ctx *ctx, const ntsm_sites *s);
/* Measurement / test knobs, to be set before ntsm_load_sites; every setting gives the same counts (the
 * pre-filters in front of the exact table only ever let too many windows through).  Names:
 *   "kernel"       0 = generic kernel (one k-mer-bitmap probe per position), 1 = paired seeds (k >= 17), 2 = wide paired
 *                  seeds (k >= 19; a 1 GiB level-1 table sized for HBM, for panels of millions of k-mers), -1 = automatic
 *   "pair_fold"    paired-seed table folded 2^n : 1 (0..4), -1 = automatic (1; 0 above 5 M site k-mers)
 *   "filter_bits"  log2 of the k-mer bitmap's bits (10..32), 0 = automatic (~80 bits per site k-mer)
 *   "launch_shape" pair kernel CTA shape: 0 = 1024x1, 1 = 1024x2 (default), 2 = 512x4, 3 = 256x8 per SM
 *   "l2_persist"   1 = launch with an L2 access-policy window that marks the probe tables persisting (default 0: measured, no effect)
 *   "parser_procs" ntsm_count_files runs its parsers as worker processes (bin/ntsm_parse_worker; a mapped plain file is parsed
 *                  in place without read()'s copy, which scales across processes but not across the threads of one):
 *                  1 on, 0 off, -1 (default) when more than 6 parser threads read plain files and there is no -m cap
 *   "device_pack"  who packs ntsm_insert_reads* input that lies in page-locked memory: 0 = the host packer threads, 1 = they
 *                  and the GPUs' own packer together, -1 (default) = by the packer threads available per GPU: 14 or
 *                  more -- the host packers; 10 to 13 -- both; fewer -- the GPUs alone (may be set any time) */
int ntsm_ctx_set_option(ntsm_ctx *ctx, const char *name, int value);

/* ---------------- packed batches: the ProdCon bulk buffers (vendor/ProdConKseqRunner.hpp:34-46) */
uint64_t ntsm_padded_positions(uint64_t n_pos); /* allocation/padding contract of a packed stream */
/* blocks until one of the ctx's pinned buffers is free; thread-safe */
int ntsm_acquire_batch(ntsm_ctx *ctx, ntsm_batch **b);
/* Packs read bases seq[*pos..len) into the batch (insertCount's input, :89).  *pos = bases of
 * this read already consumed by earlier batches (0 for a new read); a read that does not fit is
 * split and the k-1 overlapping bases are re-packed by the library so every k-mer is counted
 * once.  Returns 1 = read complete, 0 = batch full (submit, acquire, call again), <0 error. */
int ntsm_batch_append(ntsm_batch *b, const char *seq, uint64_t len, uint64_t *pos);
uint64_t ntsm_batch_positions(const ntsm_batch *b);
uint64_t ntsm_batch_bases(const ntsm_batch *b);
uint64_t ntsm_batch_reads(const ntsm_batch *b);
/* async: cudaMemcpyAsync of the packed arrays + count kernel; the buffer recycles itself */
int ntsm_submit_batch(ntsm_ctx *ctx, ntsm_batch *b);
/* gives an acquired batch back without counting it */
int ntsm_release_batch(ntsm_ctx *ctx, ntsm_batch *b);

/* standalone packer with the same layout (no ctx): packs n_reads reads buf[off[r]..off[r+1])
 * into caller memory sized for ntsm_padded_positions(sum((len + 8) & ~7)). Returns n_pos. */
/* which decode+pack implementation the host uses: "avx512vbmi", "avx2" or "scalar" (picked from the
 * CPU at load time).  force: NULL = just ask; "" = back to automatic; a name = use it if the CPU can. */
const char *ntsm_pack_isa(const char *force);
uint64_t ntsm_pack_reads(const char *buf, const uint64_t *off, uint64_t n_reads, uint32_t *bases2,
                         uint32_t *nmask, uint64_t *read_off /*nullable, n_reads+1*/);
/* same; streaming != 0 (and both arrays 64-byte aligned) packs the way the library fills its pinned batches:
 * through a cache-resident staging area, whole 64-byte lines leaving it with non-temporal stores */
uint64_t ntsm_pack_reads2(const char *buf, const uint64_t *off, uint64_t n_reads, uint32_t *bases2,
                          uint32_t *nmask, uint64_t *read_off /*nullable, n_reads+1*/, int streaming);

/* Count a packed stream that is ALREADY in device memory (padding contract as above) on
 * `cuda_stream` (a cudaStream_t; NULL = the ctx's compute stream).  Adds n_bases to the base tally. */
int ntsm_count_packed_device(ntsm_ctx *ctx, const uint32_t *d_bases2, const uint32_t *d_nmask,
                             uint64_t n_pos, uint64_t n_bases, void *cuda_stream);

/* Same, for a packed stream in HOST memory (pinned for full PCIe rate): sliced into the ctx's
 * device buffers, each slice copied with cudaMemcpyAsync and counted while the next one copies. */
int ntsm_count_packed_host(ntsm_ctx *ctx, const uint32_t *h_bases2, const uint32_t *h_nmask, uint64_t n_pos,
                           uint64_t n_bases);

/* insertCount drop-in (src/FingerPrint.hpp:89): packs into the ctx's current batch and submits
 * it when full.  Single producer; ntsm_flush submits the partial batch. */
int ntsm_insert_count(ntsm_ctx *ctx, const char *seq, uint64_t len);
int ntsm_flush(ntsm_ctx *ctx);

/* insertCount for a whole bulk of reads already in host memory -- what a consumer of the
 * ProdConKseqRunner bulk queue receives (vendor/ProdConKseqRunner.hpp:34-46: records travel 256
 * at a time).  Read r is buf[off[r] .. off[r+1]) (ASCII, decoded with the reference table).
 * `threads` producer threads take contiguous ranges of reads, pack them into pinned batches and
 * submit batch i to ctxs[i % n_ctx]; returns when every read has been submitted (the GPUs may
 * still be counting: ntsm_finalize / ntsm_sync drain).  The -m cap is not looked at inside one
 * call; callers poll ntsm_poll_totals between bulks. */
int ntsm_insert_reads(ntsm_ctx *const *ctxs, uint32_t n_ctx, const char *buf, const uint64_t *off, uint64_t n_reads,
                      uint32_t threads);
/* same for a dense matrix: read r = buf[r*stride .. r*stride + read_len) */
int ntsm_insert_reads_fixed(ntsm_ctx *const *ctxs, uint32_t n_ctx, const char *buf, uint64_t read_len, uint64_t stride,
                            uint64_t n_reads, uint32_t threads);
/* When `buf` of the two calls above lies in page-locked host memory (cudaMallocHost / cudaHostAlloc, or
 * registered with ntsm_host_register), one extra feeder thread per ctx takes blocks of reads from the
 * same queue as the `threads` host packers, DMAs their ASCII bytes to the GPU as they are and has the
 * GPU decode + pack them (pack_ascii_kernel): no host core touches those bases.  `threads` == 0 then
 * means "device packing only".  With pageable memory only the host packers run (threads >= 1).  By default
 * the split follows the packer threads available per GPU: 14 or more -- host packers only; 10 to 13 -- both; fewer --
 * the GPUs alone (measurements in ctx.cu at ntsm_ctx_device_pack; option "device_pack" forces either or both).
 * ntsm_host_register pins a caller's buffer (cudaHostRegister, portable); it costs ~0.2 s per GiB, so
 * it pays for buffers that are filled more than once. */
int ntsm_host_register(void *buf, uint64_t bytes);
int ntsm_host_unregister(void *buf);

/* m_totalKmers / m_totalCounts / m_totalBases / m_earlyTerm (:458-463) over COMPLETED batches;
 * non-blocking.  cap_reached = hits > max_counts at a batch boundary. */
int ntsm_poll_totals(ntsm_ctx *ctx, uint64_t *total_kmers, uint64_t *total_hits, uint64_t *total_bases,
                     int *cap_reached);
int ntsm_sync(ntsm_ctx *ctx);          /* drain every submitted batch */
int ntsm_reset_counts(ntsm_ctx *ctx);  /* zero counts and tallies, keep the table */
int ntsm_reset_counts_async(ntsm_ctx *ctx); /* same, enqueued on the compute stream; needs no batch in flight */
/* run every kernel / all-reduce of this ctx on the caller's cudaStream_t (NULL = the ctx's own) */
int ntsm_set_stream(ntsm_ctx *ctx, void *cuda_stream);

/* ---------------- multi-GPU: one ctx (process or thread) per GPU ---------------- */
#define NTSM_NCCL_ID_BYTES 128
/* Side effect, once per process before the first NCCL call: NCCL_DEBUG=VERSION is raised to WARN and
 * NCCL_DEBUG_FILE defaults to /dev/stderr, so that NCCL's banner cannot land in a counts file on stdout. */
int ntsm_nccl_unique_id(void *id_out);                         /* ncclGetUniqueId */
int ntsm_comm_init(ntsm_ctx *ctx, const void *id, int rank, int n_ranks);
/* sum counts (u32) and tallies (u64) over all ranks: ONE ncclAllReduce each, before the per-site max */
int ntsm_allreduce(ntsm_ctx *ctx);
/* enqueue (no host sync) on the compute stream: the all-reduces above if a communicator is attached,
 * then the per-site max/sum kernel.  ntsm_finalize calls it if it has not run yet. */
int ntsm_reduce_async(ntsm_ctx *ctx);

/* Several ctxs of ONE process (one per GPU, the CLI's --gpus N): drains them all, then ctx 0's GPU sums
 * every ctx's private counts out of peer memory over NVLink inside the per-site reduce kernel (sums
 * first, max afterwards) and sums the tallies -- no NCCL communicator.  Outputs as ntsm_finalize; the
 * ctxs must hold the same site table; afterwards they need ntsm_reset_counts before counting again
 * and ntsm_get_counts(ctxs[0]) returns the combined k-mer counts.  At most 16 ctxs. */
int ntsm_group_finalize(ntsm_ctx *const *ctxs, uint32_t n_ctx, uint32_t *max_ref, uint32_t *max_var,
                        uint32_t *sum_ref, uint32_t *sum_var, uint64_t totals[3]);

/* ---------------- results: printOptionalHeader + printCountsMax (:261-311) ---------------- */
/* drains, runs the per-site reduce kernel, copies out.  Arrays have n_sites entries;
 * totals = {total_kmers (#@TK), total_hits, total_bases}.  Any pointer may be NULL. */
int ntsm_finalize(ntsm_ctx *ctx, uint32_t *max_ref, uint32_t *max_var, uint32_t *sum_ref, uint32_t *sum_var,
                  uint64_t totals[3]);
int ntsm_get_counts(ntsm_ctx *ctx, uint32_t *counts /* [n_kmers] */);
/* The three tallies {TK, hits, bases} as they stand on this ctx (drains; does not combine anything). */
int ntsm_get_totals(ntsm_ctx *ctx, uint64_t totals[3]);
/* Exact shard merge (replaces `ntsmEval --merge`'s sum of per-site maxima, src/CompareCounts.hpp:648-657):
 * counts[i] += add[i] for every listed k-mer, tallies += totals; the per-site max is taken afterwards by
 * ntsm_finalize, so merging shards equals counting the concatenated input. */
int ntsm_add_counts(ntsm_ctx *ctx, const uint32_t *counts /* [n_kmers] */, const uint64_t totals[3]);
/* k-mer count files ("NTSMKC1": k, n_kmers, digest of the site set, tallies, u32 counts[n_kmers]):
 * save this ctx's k-mer-level result / add a saved one into this ctx (same site set and k required). */
int ntsm_counts_save(ntsm_ctx *ctx, const ntsm_sites *s, const char *path);
int ntsm_counts_load_add(ntsm_ctx *ctx, const ntsm_sites *s, const char *path);
/* getSitesCoveredInSample (:389-413) from finalize()'s maxima */
uint32_t ntsm_sites_covered(const uint32_t *max_ref, const uint32_t *max_var, uint32_t n_sites);
/* writes "#@TK..#@KS..\n#locusID...\n" + rows exactly as the reference does; returns bytes written
 * (call with buf == NULL to size), or NTSM_ERR_NOKEY */
int64_t ntsm_format_counts(const ntsm_sites *s, const uint32_t *max_ref, const uint32_t *max_var,
                           const uint32_t *sum_ref, const uint32_t *sum_var, uint64_t total_kmers, char *buf,
                           size_t cap);
/* printInfoSummary text (:313-333) */
int64_t ntsm_format_summary(const ntsm_sites *s, const uint64_t totals[3], uint32_t covered, char *buf, size_t cap);

/* introspection (bench/tests): kernels launched so far, pre-filter size (log2 bits), table slots */
uint64_t ntsm_ctx_launches(const ntsm_ctx *ctx);
const char *ntsm_ctx_kernel_name(const ntsm_ctx *ctx);   /* the count kernel this ctx launches, as a profiler lists it */
uint32_t ntsm_ctx_filter_bits(const ntsm_ctx *ctx);
uint32_t ntsm_ctx_table_capacity(const ntsm_ctx *ctx);
void ntsm_ctx_pcie_bytes(const ntsm_ctx *ctx, uint64_t *h2d, uint64_t *d2h); /* bytes this ctx's data path has copied host->device / device->host so far */
int ntsm_ctx_l2_window(const ntsm_ctx *ctx);             /* 0 = no L2 access-policy window, else its hit ratio in percent */
uint64_t ntsm_ctx_probe_bytes(const ntsm_ctx *ctx);      /* bytes of the probe tables (paired-seed table + k-mer bitmap) */

/* ---------------- record reader: kseq_read (vendor/kseq.h:178-219) ---------------- */
int ntsm_reader_open(ntsm_reader **out, const char *path);
/* returns sequence length >= 0 and sets *seq (valid until the next call), or -1 end of file,
 * -2 truncated/mismatched quality, -3 stream error -- the codes of kseq.h:171-176 */
int64_t ntsm_reader_next(ntsm_reader *r, const char **seq);
const char *ntsm_reader_name(const ntsm_reader *r);
void ntsm_reader_close(ntsm_reader *r);

/* ---------------- byte source: gzopen/gzread as kseq uses them (src/FingerPrint.hpp:50,
 * vendor/kseq.h:68-79 over KSEQ_INIT(gzFile, gzread)) ----------------
 * Delivers exactly the bytes zlib's gzread would for the file (plain files pass through, gzip members
 * are inflated one after the other, trailing garbage is ignored, an error ends the stream after the
 * bytes that preceded it), but inflates with the library's own decoder over a memory-mapped file
 * and, for BGZF files, block-parallel on up to `helpers` extra threads.  ntsm_gz_mode: "zlib" |
 * "fast" | "bgzf" (what is producing bytes now); ntsm_gz_fell_back: 1 once an irregular member made
 * the source hand the rest of the file to zlib.  ntsm_reader_open2 is ntsm_reader_open with helpers. */
typedef struct ntsm_gz ntsm_gz;
int ntsm_gz_open(ntsm_gz **out, const char *path, int helpers);
int ntsm_gz_read(ntsm_gz *g, void *dst, unsigned n);   /* gzread's contract: short only at the end, 0 = end, -1 = error */
const char *ntsm_gz_mode(const ntsm_gz *g);
int ntsm_gz_fell_back(const ntsm_gz *g);
uint64_t ntsm_gz_parallel_chunks(const ntsm_gz *g);   /* chunks of single gzip members inflated by helper threads and accepted so far */
void ntsm_gz_close(ntsm_gz *g);
uint32_t ntsm_crc32(uint32_t crc, const void *buf, uint64_t len);   /* the CRC-32 of gzip trailers (carry-less multiply when available) */
int ntsm_reader_open2(ntsm_reader **out, const char *path, int helpers);
/* which newline scanner the reader's FASTQ fast path uses: "avx512", "avx2" or "memchr" (picked from the
 * CPU at load time).  force: NULL = just ask; "" = back to automatic; a name = use it if the CPU can. */
const char *ntsm_scan_isa(const char *force);

/* ---------------- whole path: FingerPrint::computeCounts (src/FingerPrint.hpp:46-87) -------- */
/* Reads every file with `threads` parser threads (one file per thread at a time, like the
 * reference's omp parallel for over files), packs reads into pinned batches and streams batch i
 * to ctxs[i % n_ctx].  With a -m cap (cfg.max_counts of ctxs[0]) the summed hit tally is checked
 * after every submitted batch; parsing stops once it is exceeded and *early (nullable) is set =
 * m_earlyTerm.  With one parser thread (threads == 1 or one file: the reference's deterministic
 * -t 1) the last batch is then cut back to the read that crossed the cap, so the counters are
 * exactly the reference's (src/FingerPrint.hpp:473-488); with more threads the stop is
 * batch-granular and every read parsed before it is counted.  Returns NTSM_ERR_IO (text in ntsm_last_error)
 * if a file cannot be opened. */
int ntsm_count_files(ntsm_ctx *const *ctxs, uint32_t n_ctx, const char *const *paths, uint32_t n_paths,
                     uint32_t threads, int verbose, int *early);

/* ---------------- multi-sample matrix: MultiCount driven by VCFConvert ----------------
 * src/MultiCount.hpp:36-289, src/VCFConvert.hpp:40-218, src/ntSeqMatchVCF.cpp.  The matrix
 * m_matCounts[sample][dense k-mer index] (uint8) lives in device memory on top of a ctx whose site table is
 * loaded (ntsm_load_siteset = MultiCount::initCountsHash, which is FingerPrint's).  Results equal the
 * reference's classes run with one thread (tests/golden/vcf, made by tools/ref_vcf_harness.cpp).  Calls return
 * NTSM_ERR_NOKEY where the reference process dies (uncaught exception / failed assert, exit 134). */
/* One caller at a time per ntsm_multi / ntsm_vcf (the library brings its own threads: `threads` arguments); the
 * object works on the ctx's compute stream as it was when ntsm_multi_create ran. */
typedef struct ntsm_multi ntsm_multi; /* MultiCount */
typedef struct ntsm_vcf ntsm_vcf;     /* VCFConvert */
int ntsm_multi_create(ntsm_multi **out, ntsm_ctx *ctx, uint32_t n_samples); /* MultiCount(sampleIDs) :43-49; the ctx must outlive it */
void ntsm_multi_destroy(ntsm_multi *m);
uint32_t ntsm_multi_n_samples(const ntsm_multi *m);
/* insertCount(sampleIndex, hashVal, multi) :52-70, one call (a one-thread kernel: for drop-in completeness) */
int ntsm_multi_insert_count(ntsm_multi *m, uint32_t sample, uint64_t hash, uint32_t multi);
/* The inner loops of VCFConvert::count (:148-169) for n_lines SNP lines at once.  Line l has two windows of
 * the reference genome, windows[(2l + a) * wstride .. + lens[2l + a]): a = 0 with the reference allele, a = 1
 * with the alternative allele in the middle (ASCII, decoded with the reference's table); and one genotype code
 * per sample, genotypes[l * n_samples + s]: 0 = "0|0" (and anything unrecognised), 1 = "0|1" / "1|0", 2 = "1|1".
 * Every k-mer of window a that is a site k-mer is inserted for every sample: 2 * multi when the sample is
 * homozygous for allele a, multi when heterozygous.  The first non-zero value written to a cell stays (the
 * order is the reference's with one thread: line, a, offset in the window); later different values only warn. */
int ntsm_multi_insert_windows(ntsm_multi *m, const char *windows, uint32_t wstride, const uint16_t *lens,
                              const uint8_t *genotypes, uint32_t n_lines, uint32_t multi);
/* "Warning: Inconsistent k-mer counts, ..." lines (:60-61) raised so far, in the reference's serial order */
uint64_t ntsm_multi_n_warnings(const ntsm_multi *m);
int64_t ntsm_multi_warnings_text(const ntsm_multi *m, char *buf, size_t cap); /* returns the full length */
int ntsm_multi_get_matrix(ntsm_multi *m, uint8_t *out /* [n_samples][n_kmers] */);
/* introspection (bench): device time so far by CUDA events -- ms[0] k-merize + lookup + occurrence lists, ms[1] the
 * fill kernel's passes, ms[2] the norm-matrix kernels; cells = (occurring k-mer, sample) pairs walked per fill pass */
void ntsm_multi_kernel_ms(const ntsm_multi *m, double ms[3], uint64_t *cells);
/* printCountsMax(index) :93-138 as arrays (n_sites each; any may be NULL) and as text ("\n#locusID..." + rows:
 * no #@TK / #@KS lines); the text call returns the length, or NTSM_ERR_NOKEY */
int ntsm_multi_counts_max(ntsm_multi *m, uint32_t sample, uint32_t *max_ref, uint32_t *max_var, uint32_t *sum_ref,
                          uint32_t *sum_var);
int64_t ntsm_multi_format_counts(ntsm_multi *m, const ntsm_sites *s, uint32_t sample, char *buf, size_t cap);
/* printNormMatrix :148-203, the numbers: values[site * n_samples + sample] = maxREF / (maxREF + maxVAR), or
 * DBL_MAX (MultiCount::UNDEF) where both are zero; sums[site] = the sum of the site's defined values taken in
 * sample order (the reference divides it by n_samples in long double to get the centre).  Either may be NULL. */
int ntsm_multi_norm_matrix(ntsm_multi *m, double *values, double *sums);
/* printNormMatrix, the two files (matrix with missing values replaced by the centre; one centre per line),
 * digit for digit including the stream precision that switches to 19 at the first missing value */
int ntsm_multi_write_norm_matrix(ntsm_multi *m, const ntsm_sites *s, const char *const *sample_ids, const char *matrix_path,
                                 const char *center_path, uint32_t threads /* host threads that format the text */);

/* VCFConvert() + count(vcf) :42-174: reads the reference genome (plain or gz FASTA) and the multi-sample VCF
 * (plain text as upstream; gzip / bgzip'ed files are inflated on the way, an extension), cuts the window around
 * every SNP line (getSeqFromSite :202-215) and inserts batches of lines on the GPU.  NTSM_ERR_ARG for a site whose window would start before its chromosome (undefined upstream). */
int ntsm_vcf_convert(ntsm_vcf **out, ntsm_ctx *ctx, const ntsm_sites *s, const char *ref_path, const char *vcf_path,
                     uint32_t multi /* opt::multi, 20 */, uint32_t window /* opt::window, 31 */,
                     uint32_t threads /* opt::threads: host threads that parse the VCF (and later format the matrix);
                                         the result does not depend on it */, int verbose);
void ntsm_vcf_destroy(ntsm_vcf *v);
/* The host half of the above on its own, no device needed: the SNP lines of the VCF as ntsm_multi_insert_windows takes
 * them (windows [2 * count][wstride] + lens [2 * count], genotypes [count][n_samples]).  Returns the code
 * ntsm_vcf_convert would; *out is set even then and holds the lines in front of the fatal one.  out == NULL: parse
 * only, keep nothing (what the host half costs). */
typedef struct ntsm_vcf_lines ntsm_vcf_lines;
int ntsm_vcf_parse(ntsm_vcf_lines **out, const char *ref_path, const char *vcf_path, uint32_t window, uint32_t threads, int verbose);
void ntsm_vcf_lines_free(ntsm_vcf_lines *l);
uint32_t ntsm_vcf_lines_n_samples(const ntsm_vcf_lines *l);
const char *ntsm_vcf_lines_sample_id(const ntsm_vcf_lines *l, uint32_t i);
uint64_t ntsm_vcf_lines_count(const ntsm_vcf_lines *l);
uint32_t ntsm_vcf_lines_wstride(const ntsm_vcf_lines *l);
const char *ntsm_vcf_lines_windows(const ntsm_vcf_lines *l);
const uint16_t *ntsm_vcf_lines_lens(const ntsm_vcf_lines *l);
const uint8_t *ntsm_vcf_lines_genotypes(const ntsm_vcf_lines *l);
/* test knob: bytes per region in which a VCF that is not a plain regular file (gzip, pipe) is read; 0 = just ask */
uint64_t ntsm_vcf_stream_chunk(uint64_t bytes);
/* which genotype-column decoder the VCF parser uses: 0 byte-wise, 1 AVX2 (8 columns a step), 2 AVX-512 (16), picked from the
 * CPU; force >= 0 sets it (capped at what the CPU has), -1 = back to automatic, -2 = just ask.  Same codes either way. */
int ntsm_vcf_genotype_isa(int force);
ntsm_multi *ntsm_vcf_multi(ntsm_vcf *v);
uint32_t ntsm_vcf_n_samples(const ntsm_vcf *v);
const char *ntsm_vcf_sample_id(const ntsm_vcf *v, uint32_t i);
uint64_t ntsm_vcf_lines_counted(const ntsm_vcf *v);
int ntsm_vcf_output_matrix(ntsm_vcf *v, const char *prefix); /* outputMatrix :186-193: <prefix>_matrix.tsv, <prefix>_center.txt */
int ntsm_vcf_output_counts(ntsm_vcf *v, const char *dir);    /* outputCounts :173-184: <dir>/<sampleID>.counts.txt (NULL = cwd) */
/* the ntsmVCF command line (src/ntSeqMatchVCF.cpp:53-217); returns the process exit code */
int ntsm_vcf_main(int argc, char **argv);

/* the ntsmCount command line (src/ntSeqMatchCount.cpp:53-184); returns the process exit code */
int ntsm_main(int argc, char **argv);

#ifdef __cplusplus
}
#endif
#endif /* NTSM_B200_H */
