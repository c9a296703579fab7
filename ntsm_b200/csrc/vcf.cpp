// vcf.cpp -- host side of the multi-sample matrix path: VCFConvert (src/VCFConvert.hpp:40-218) and the
// ntsmVCF command line (src/ntSeqMatchVCF.cpp:53-217) over the device MultiCount of multi.cu.
//
// The host only reads text: the reference genome (kseq grammar, fastx.h), the VCF's header and data lines, the
// window around each SNP (getSeqFromSite, :202-215) and one genotype code per sample.  Batches of lines go to
// ntsm_multi_insert_windows, where every k-mer x sample insert of VCFConvert::count's inner loops (:148-169) runs on
// the GPU.  The printers format what the device reduced (ntsm_multi_norm_matrix, ntsm_multi_counts_max) with the
// ostream state the reference carries (printNormMatrix's sticky setprecision(19), src/MultiCount.hpp:148-203).
//
// Where the reference process dies (uncaught exception / failed assert, rc 134) the calls return NTSM_ERR_NOKEY;
// where it has undefined behaviour (a site closer than window/2 to the start of its chromosome, or past its end:
// getSeqFromSite reads outside the sequence) they return NTSM_ERR_ARG.
#include <errno.h>
#include <immintrin.h>
#include <getopt.h>
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <fstream>
#include <iostream>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/ntsm_b200.h"
#include "fastx.h"
#include "internal.h"

struct ntsm_vcf {
	ntsm_ctx *ctx = nullptr;
	const ntsm_sites *sites = nullptr;
	ntsm_multi *multi = nullptr;
	std::vector<std::string> sample_ids;        // m_sampleIDs (:192)
	uint64_t lines_counted = 0;                 // SNP lines whose windows were inserted
	uint32_t threads = 1;                       // opt::threads: parser threads of count(), formatter threads of outputMatrix()
};

namespace {

int vfail(ntsm_ctx *ctx, int code, const std::string &text)
{
	ntsm_ctx_set_error(ctx, text.c_str());
	return code;
}

// The input as a sequence of regions of whole lines.  The reference reads the VCF through an ifstream (:69): plain text.
// A plain regular file is ONE region, scanned in place through GzSource's mapping; anything else arrives in regions of
// g_stream_chunk bytes cut at line ends -- and a gzip / bgzip'ed VCF (what cohort VCFs are shipped as: tens of GB of text
// that could not be held whole; upstream would read the compressed bytes as text and find no header) is inflated on the
// way by the library's own decoder, BGZF blocks on `helpers` threads (gzsource.h).
size_t g_stream_chunk = 64u << 20;

struct Regions {
	ntsm::GzSource src;
	bool mapped = false, done = false;
	const uint8_t *base = nullptr;
	size_t size = 0;
	std::vector<char> buf;
	size_t carry = 0;                                              // bytes of an unfinished line at the front of buf
	bool open(const char *path, int helpers)
	{
		if (!src.open(path, helpers, true)) return false;
		mapped = src.mapped(&base, &size);
		return true;
	}
	// the next region [*p, *p + *n): whole lines, except that the LAST region ends where the input ends; false when there is none
	bool next(const char **p, size_t *n, bool *last)
	{
		if (done) return false;
		if (mapped) {
			done = true;
			*p = (const char *)base;
			*n = size;
			*last = true;
			return size > 0;
		}
		if (buf.size() < g_stream_chunk) buf.resize(g_stream_chunk);
		size_t have = carry;
		for (;;) {
			while (have < buf.size()) {
				const int got = src.read(buf.data() + have, (unsigned)std::min<size_t>(buf.size() - have, 1u << 30));
				if (got <= 0) { done = true; break; }
				have += (size_t)got;
			}
			if (done) {
				*p = buf.data();
				*n = have;
				*last = true;
				carry = 0;
				return have > 0;
			}
			// cut at the last line end; a line longer than the buffer makes the buffer grow
			size_t cut = have;
			while (cut > 0 && buf[cut - 1] != '\n') --cut;
			if (cut == 0) {
				buf.resize(buf.size() * 2);
				continue;
			}
			region_end = cut;
			*p = buf.data();
			*n = cut;
			*last = false;
			carry = have - cut;
			return true;
		}
	}
	// called before the following next(): moves the unfinished line to the front
	void recycle()
	{
		if (!mapped && carry && region_end) memmove(buf.data(), buf.data() + region_end, carry);
		region_end = 0;
	}
	size_t region_end = 0;
};

struct Field { const char *p; size_t n; };

// tab-separated fields of one line; getline(ss, item, '\t') semantics: a trailing tab yields a last empty field.
// At most max_fields are split off; *rest (nullable) = where the next field starts, or nullptr if the line ended.
void split(const char *p, size_t n, std::vector<Field> &f, size_t max_fields = (size_t)-1, const char **rest = nullptr)
{
	f.clear();
	const char *end = p + n, *at = p;
	if (rest) *rest = nullptr;
	for (;;) {
		const char *t = (const char *)memchr(at, '\t', (size_t)(end - at));
		if (!t) { f.push_back({ at, (size_t)(end - at) }); break; }
		f.push_back({ at, (size_t)(t - at) });
		at = t + 1;
		if (f.size() == max_fields) {
			if (rest) *rest = at;
			break;
		}
	}
}

struct Chrom { std::string seq; };

inline uint8_t genotype_code(const char *g, size_t n)
{   // :138-144; anything else keeps the vector's value-initialised hom1
	if (n != 3 || g[1] != '|') return 0;
	if (g[0] == '0') return g[2] == '1' ? 1 : 0;           // "0|0" hom1, "0|1" het ("0|x" otherwise: hom1)
	if (g[0] == '1') return g[2] == '0' ? 1 : g[2] == '1' ? 2 : 0;
	return 0;
}

// The usual cohort VCF column is "a|b\t" with a, b in {0, 1}: four bytes whose code is simply a + b.  16 (AVX-512) or 8
// (AVX2) columns are checked and decoded at once: as little-endian uint32 lanes, (lane & 0xFFFEFFFE) must equal
// '0' '|' '0' '\t', and the low bits of bytes 0 and 2 are the two alleles.  Returns false (nothing consumed) when any of
// the columns is something else -- "./.", "0/1", a multi-digit allele, the line's last column (it ends in '\n') -- and the
// caller's byte-wise loop takes over, so the codes are genotype_code's either way.
constexpr uint32_t kGtPattern = (uint32_t)'0' | (uint32_t)'|' << 8 | (uint32_t)'0' << 16 | (uint32_t)'\t' << 24;

__attribute__((target("avx512f,avx512bw,bmi2"))) bool decode16_avx512(const char *q, uint32_t *codes32)
{
	const __m512i v = _mm512_loadu_si512((const void *)q);
	if (_mm512_cmpeq_epi32_mask(_mm512_and_si512(v, _mm512_set1_epi32((int)0xFFFEFFFEu)), _mm512_set1_epi32((int)kGtPattern)) != 0xFFFF) return false;
	const uint32_t a = _mm512_test_epi32_mask(v, _mm512_set1_epi32(1)), b = _mm512_test_epi32_mask(v, _mm512_set1_epi32(1 << 16));
	*codes32 = (uint32_t)_pdep_u32(a ^ b, 0x55555555u) | (uint32_t)_pdep_u32(a & b, 0xAAAAAAAAu);   // a + b, two bits per column
	return true;
}

__attribute__((target("avx2,bmi2"))) bool decode8_avx2(const char *q, uint32_t *codes16)
{
	const __m256i v = _mm256_loadu_si256((const __m256i *)q);
	const __m256i ok = _mm256_cmpeq_epi32(_mm256_and_si256(v, _mm256_set1_epi32((int)0xFFFEFFFEu)), _mm256_set1_epi32((int)kGtPattern));
	if (_mm256_movemask_ps(_mm256_castsi256_ps(ok)) != 0xFF) return false;
	const uint32_t a = (uint32_t)_mm256_movemask_ps(_mm256_castsi256_ps(_mm256_slli_epi32(v, 31)));
	const uint32_t b = (uint32_t)_mm256_movemask_ps(_mm256_castsi256_ps(_mm256_slli_epi32(v, 15)));
	*codes16 = (uint32_t)_pdep_u32(a ^ b, 0x5555u) | (uint32_t)_pdep_u32(a & b, 0xAAAAu);
	return true;
}

const int g_gt_isa = [] {
	__builtin_cpu_init();                                          // this runs as a static initialiser of a shared library
	if (__builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("bmi2")) return 2;
	return __builtin_cpu_supports("avx2") && __builtin_cpu_supports("bmi2") ? 1 : 0;
}();
int g_gt_isa_forced = -1;                                          // tests: 0 scalar, 1 AVX2, 2 AVX-512 (never above what the CPU has)
inline int gt_isa() { return g_gt_isa_forced >= 0 && g_gt_isa_forced < g_gt_isa ? g_gt_isa_forced : g_gt_isa; }

// n items over `threads` workers (an atomic counter hands them out); threads <= 1 runs inline
template <class F> void parallel_for(size_t n, uint32_t threads, F fn)
{
	if (threads <= 1 || n <= 1) {
		for (size_t i = 0; i < n; ++i) fn(i);
		return;
	}
	std::atomic<size_t> next{ 0 };
	std::vector<std::thread> pool;
	const size_t nt = std::min<size_t>(threads, n);
	for (size_t t = 0; t < nt; ++t)
		pool.emplace_back([&] {
			for (size_t i; (i = next.fetch_add(1)) < n;) fn(i);
		});
	for (std::thread &t : pool) t.join();
}

struct ParseEnv {
	const std::vector<Chrom> *chroms;
	const std::unordered_map<std::string, uint32_t> *chr_ids;
	uint32_t S, window, wstride;
	int verbose;
};

// what one worker makes of a run of data lines: the windows and packed genotypes of the lines that are SNPs
struct LineBatch {
	std::vector<char> windows;
	std::vector<uint16_t> lens;
	std::vector<uint32_t> geno2;                // 2 bits per sample, (S + 15) / 16 words per line
	uint32_t n = 0;
	int err = 0;                                // first line that ends the run, as upstream: NTSM_ERR_NOKEY (dies) / NTSM_ERR_ARG (undefined)
	std::string msg;
};

std::mutex g_verbose_mu;

// VCFConvert::count's per-line work up to the inserts (:103-146) for the whole lines of [at, end)
void parse_lines(const ParseEnv &E, const char *at, const char *end, LineBatch &B)
{
	const uint32_t S = E.S, window = E.window, wstride = E.wstride, half = window / 2, gwords = (S + 15) / 16;
	size_t n_lines = 0;
	for (const char *p = at; p < end; ++n_lines) p = (const char *)memchr(p, '\n', (size_t)(end - p)) + 1;
	B.windows.assign(n_lines * 2 * wstride, 0);
	B.lens.assign(n_lines * 2, 0);
	B.geno2.assign(n_lines * gwords + 1, 0u);
	B.n = 0;
	B.err = 0;
	B.msg.clear();
	std::vector<Field> f;
	auto die = [&](int code, const std::string &msg) {
		B.err = code;
		B.msg = msg;
	};
	while (at < end) {
		const char *nl = (const char *)memchr(at, '\n', (size_t)(end - at));
		const char *cols;                                          // the sample columns, if the line has any
		split(at, (size_t)(nl - at), f, 9, &cols);
		at = nl + 1;
		// getline on an exhausted stringstream leaves `item` as it was: a missing field reads as the last present one
		auto field = [&](size_t i) -> const Field & { return f[i < f.size() ? i : f.size() - 1]; };
		const std::string chr(field(0).p, field(0).n);
		const std::string pos_text(field(1).p, field(1).n);
		char *pe;
		errno = 0;
		const long loc_l = strtol(pos_text.c_str(), &pe, 10);       // stoi (:113)
		if (pe == pos_text.c_str() || errno == ERANGE || loc_l > INT_MAX || loc_l < INT_MIN)
			return die(NTSM_ERR_NOKEY, "POS '" + pos_text + "' is not an int: the reference dies in stoi (src/VCFConvert.hpp:113)");
		if (E.verbose > 2) {
			std::lock_guard<std::mutex> lk(g_verbose_mu);          // `omp critical` (:117-118)
			std::cerr << "Processing site: " << std::string(field(2).p, field(2).n) << std::endl;
		}
		if (field(3).n == 1 && field(3).p[0] == '.') continue;     // :121-123
		if (field(4).n != 1) continue;                             // :125-127: ALT must be one character (REF's length is not looked at)
		const char alt = field(4).p[0];
		// getSeqFromSite (:202-215)
		const auto it = E.chr_ids->find(chr);
		if (it == E.chr_ids->end())
			return die(NTSM_ERR_NOKEY, "chromosome '" + chr + "' is not in the reference: the reference dies in robin_map::at (src/VCFConvert.hpp:204)");
		const std::string &seq = (*E.chroms)[it->second].seq;
		if (loc_l < (long)half + 1 || (size_t)(loc_l - half - 1) > seq.size())
			return die(NTSM_ERR_ARG, "site " + chr + ":" + pos_text + " lies outside what getSeqFromSite can cut (src/VCFConvert.hpp:207-209 reads outside the sequence there)");
		const size_t offset = (size_t)loc_l - half - 1;
		size_t avail = std::min<size_t>(seq.size() - offset, window);            // strncpy stops at the sequence's NUL (or one inside it) and pads with NULs
		avail = strnlen(seq.data() + offset, avail);
		char *wr = B.windows.data() + (size_t)B.n * 2 * wstride, *wv = wr + wstride;
		memset(wr, 0, 2 * (size_t)wstride);                        // a line given up below leaves its slot to the next one
		memcpy(wr, seq.data() + offset, avail);
		memcpy(wv, seq.data() + offset, avail);
		if (half < wstride) wv[half] = alt;                        // :211
		// std::string(refStr): up to the first NUL -- which may also be a NUL byte inside the genome or the ALT column
		B.lens[2 * (size_t)B.n] = (uint16_t)strnlen(wr, window);
		B.lens[2 * (size_t)B.n + 1] = (uint16_t)strnlen(wv, window);
		// sample columns (:129-146): "0|0" hom1, "0|1" / "1|0" het, "1|1" hom2, anything else stays hom1
		uint32_t *g = B.geno2.data() + (size_t)B.n * gwords;
		for (uint32_t w = 0; w < gwords; ++w) g[w] = 0;
		size_t n_cols = 0;
		const int isa = gt_isa();
		for (const char *q = cols; q;) {
			// 16 / 8 regular columns at a time where the CPU can (appended at bit 2 * n_cols of the line's words)
			uint32_t codes;
			if (isa == 2 && n_cols + 16 <= S && q + 64 <= nl && decode16_avx512(q, &codes)) {
				const uint64_t v = (uint64_t)codes << (2 * (n_cols & 15));
				g[n_cols >> 4] |= (uint32_t)v;
				if (v >> 32) g[(n_cols >> 4) + 1] |= (uint32_t)(v >> 32);
				n_cols += 16;
				q += 64;
				continue;
			}
			if (isa == 1 && n_cols + 8 <= S && q + 32 <= nl && decode8_avx2(q, &codes)) {
				const uint64_t v = (uint64_t)codes << (2 * (n_cols & 15));
				g[n_cols >> 4] |= (uint32_t)v;
				if (v >> 32) g[(n_cols >> 4) + 1] |= (uint32_t)(v >> 32);
				n_cols += 8;
				q += 32;
				continue;
			}
			// almost every other column is three characters and a tab
			const char *t = q + 3 < nl && q[3] == '\t' && q[0] != '\t' && q[1] != '\t' && q[2] != '\t' ? q + 3 : (const char *)memchr(q, '\t', (size_t)(nl - q));
			const char *fe = t ? t : nl;
			if (!t && fe == q) break;                              // the same `while (getline(...))` (:136): nothing after the line's last tab, no column
			if (n_cols < S) g[n_cols >> 4] |= (uint32_t)genotype_code(q, (size_t)(fe - q)) << (2 * (n_cols & 15));
			++n_cols;
			q = t ? t + 1 : nullptr;
		}
		if (n_cols != S)
			return die(NTSM_ERR_NOKEY, "line for " + chr + ":" + pos_text + " has " + std::to_string(n_cols) + " sample columns, the header names " +
			                               std::to_string(S) + ": the reference dies on its assert (src/VCFConvert.hpp:146)");
		++B.n;
	}
}

}  // namespace

extern "C" void ntsm_vcf_destroy(ntsm_vcf *v)
{
	if (!v) return;
	ntsm_multi_destroy(v->multi);
	delete v;
}

extern "C" ntsm_multi *ntsm_vcf_multi(ntsm_vcf *v) { return v ? v->multi : nullptr; }
extern "C" uint32_t ntsm_vcf_n_samples(const ntsm_vcf *v) { return v ? (uint32_t)v->sample_ids.size() : 0; }
extern "C" const char *ntsm_vcf_sample_id(const ntsm_vcf *v, uint32_t i) { return v && i < v->sample_ids.size() ? v->sample_ids[i].c_str() : nullptr; }
extern "C" uint64_t ntsm_vcf_lines_counted(const ntsm_vcf *v) { return v ? v->lines_counted : 0; }

namespace {

// The host half of VCFConvert::VCFConvert (:42-59) + VCFConvert::count (:62-146): genome, header, and every SNP line
// turned into two windows + packed genotypes, handed on in file order.  on_header(sample IDs) is called once, then
// on_batch(batch, wstride) for every run of lines; either may return non-zero to stop.  No device is touched here.
template <class H, class B>
int vcf_stream(const char *ref_path, const char *vcf_path, uint32_t window, uint32_t threads, int verbose, std::string &err, H on_header, B on_batch)
{
	if (window == 0 || window > 65535) { err = "window must be in 1..65535"; return NTSM_ERR_ARG; }
	// the reference genome: every record whole; a later record of the same name replaces the earlier one (:47-58)
	if (verbose > 1) std::cerr << "Loading Reference " << ref_path << std::endl;
	std::vector<Chrom> chroms;
	std::unordered_map<std::string, uint32_t> chr_ids;
	{
		ntsm::FastxReader rd;
		if (!rd.open(ref_path, 0, false)) { err = std::string("file ") + ref_path + " cannot be opened"; return NTSM_ERR_IO; }
		int64_t l;
		while ((l = rd.next()) >= 0) {
			chr_ids[rd.name()] = (uint32_t)chroms.size();
			chroms.emplace_back();
			chroms.back().seq.assign(rd.seq(), (size_t)l);
		}
	}

	if (verbose > 1) std::cerr << "Reading VCF file: " << vcf_path << std::endl;
	Regions regions;
	if (!regions.open(vcf_path, (int)std::max(1u, threads) - 1)) { err = std::string("file ") + vcf_path + " cannot be opened"; return NTSM_ERR_IO; }

	std::vector<std::string> sample_ids;
	std::vector<Field> f;
	bool in_header = true;
	uint32_t S = 0;
	int rc = NTSM_OK;
	const uint32_t wstride = (window + 15u) & ~15u;
	const uint32_t batch_lines = 2048;
	const uint32_t T = std::max(1u, threads);
	const size_t round = std::max<size_t>(4 * (size_t)T, 16);
	std::vector<std::pair<const char *, const char *>> spans;       // [begin, end) of each batch: whole lines, '\n' included
	std::vector<LineBatch> batches;
	auto header_done = [&]() -> int {
		in_header = false;
		S = (uint32_t)sample_ids.size();
		if (verbose > 1) std::cerr << "Starting multicount of each rsID for " << S << " samples." << std::endl;
		return on_header(sample_ids);
	};

	const char *rp;
	size_t rn;
	bool last;
	while (regions.next(&rp, &rn, &last)) {
		const char *at = rp, *const end = rp + rn;
		// header: lines are looked at until the one whose first field is "#CHROM"; 8 more fields are skipped, the rest
		// are the sample IDs (:71-93).  A last line without '\n' is still a line here (the stream only goes bad after it).
		while (in_header && at < end) {
			const char *nl = (const char *)memchr(at, '\n', (size_t)(end - at));
			const char *le = nl ? nl : end;
			const size_t len = (size_t)(le - at);
			if (len == 0) {                                        // line.at(0) throws
				err = "empty line before the #CHROM line: the reference dies in string::at (src/VCFConvert.hpp:74)";
				return NTSM_ERR_NOKEY;
			}
			const bool is_header = at[0] == '#';
			if (is_header) split(at, len, f);
			at = nl ? nl + 1 : end;
			if (is_header && f[0].n == 6 && memcmp(f[0].p, "#CHROM", 6) == 0) {
				// `while (getline(ss, item, '\t'))` (:88): a tab at the very end of the line leaves nothing to extract and ends
				// the loop -- no empty last ID (an empty field between two tabs does count)
				if (f.size() > 9 && f.back().n == 0) f.pop_back();
				for (size_t i = 9; i < f.size(); ++i) sample_ids.emplace_back(f[i].p, f[i].n);
				if ((rc = header_done())) return rc;
			}
		}
		if (in_header) {
			regions.recycle();
			continue;
		}

		// The data lines.  Parsing is per line and independent, so `threads` workers (opt::threads: the reference runs this
		// loop under `omp parallel`, :99) parse batches of lines side by side; the batches are then handed on in file
		// order, which -- unlike upstream with more than one thread -- keeps "the first writer of a cell wins" the
		// one-thread result whatever `threads` is.
		spans.clear();
		{
			const char *b = at;
			uint32_t in_batch = 0;
			while (at < end) {
				const char *nl = (const char *)memchr(at, '\n', (size_t)(end - at));
				if (!nl) break;                                    // :101-108: the getline that hits end of file leaves the stream not good(): that line is dropped
				at = nl + 1;
				if (++in_batch == batch_lines) {
					spans.emplace_back(b, at);
					b = at;
					in_batch = 0;
				}
			}
			if (in_batch) spans.emplace_back(b, at);
		}
		const ParseEnv env{ &chroms, &chr_ids, S, window, wstride, verbose };
		if (batches.size() < std::min(round, spans.size())) batches.resize(std::min(round, spans.size()));
		for (size_t r0 = 0; r0 < spans.size(); r0 += round) {
			const size_t nb = std::min(round, spans.size() - r0);
			parallel_for(nb, T, [&](size_t i) { parse_lines(env, spans[r0 + i].first, spans[r0 + i].second, batches[i]); });
			for (size_t i = 0; i < nb; ++i) {
				LineBatch &b = batches[i];
				if (b.n && (rc = on_batch(b, wstride))) return rc; // the lines before a fatal one are handed on, as upstream inserts them before it dies
				if (b.err) {
					err = b.msg;
					return b.err;
				}
			}
		}
		regions.recycle();
	}
	if (in_header && (rc = header_done())) return rc;              // no #CHROM line at all: no samples, no data lines (:71-98)
	return NTSM_OK;
}

}  // namespace

// VCFConvert::VCFConvert (:42-59) + VCFConvert::count (:62-174)
extern "C" int ntsm_vcf_convert(ntsm_vcf **out, ntsm_ctx *ctx, const ntsm_sites *sites, const char *ref_path, const char *vcf_path,
                                uint32_t multi, uint32_t window, uint32_t threads, int verbose)
{
	if (!out || !ctx || !sites || !ref_path || !vcf_path) return vfail(ctx, NTSM_ERR_ARG, "ntsm_vcf_convert: null argument");
	*out = nullptr;
	ntsm_vcf *v = new ntsm_vcf();
	v->ctx = ctx;
	v->sites = sites;
	v->threads = std::max(1u, threads);
	std::string err;
	bool own_error = false;                                        // the failing call has already left its text on the ctx
	const int rc = vcf_stream(
	    ref_path, vcf_path, window, threads, verbose, err,
	    [&](const std::vector<std::string> &ids) {
		    v->sample_ids = ids;
		    const int r = ntsm_multi_create(&v->multi, ctx, (uint32_t)ids.size());
		    own_error = r != 0;
		    return r;
	    },
	    [&](const LineBatch &b, uint32_t wstride) {
		    const int r = ntsm_multi_insert_windows_packed(v->multi, b.windows.data(), wstride, b.lens.data(), b.geno2.data(), b.n, multi);
		    own_error = r != 0;
		    v->lines_counted += b.n;
		    return r;
	    });
	if (rc) {
		ntsm_vcf_destroy(v);
		return own_error ? rc : vfail(ctx, rc, err);
	}
	*out = v;
	return NTSM_OK;
}

// The host half alone, kept whole in memory: what ntsm_vcf_convert would insert.  No device needed (tests of the parser;
// callers that feed ntsm_multi_insert_windows themselves).
struct ntsm_vcf_lines {
	std::vector<std::string> sample_ids;
	uint32_t wstride = 0;
	uint64_t n = 0;
	std::vector<char> windows;
	std::vector<uint16_t> lens;
	std::vector<uint8_t> genotypes;             // [n][n_samples] 0 hom1, 1 het, 2 hom2
};

extern "C" int ntsm_vcf_parse(ntsm_vcf_lines **out, const char *ref_path, const char *vcf_path, uint32_t window, uint32_t threads, int verbose)
{
	if (!ref_path || !vcf_path) return vfail(nullptr, NTSM_ERR_ARG, "ntsm_vcf_parse: null argument");
	if (!out) {                                                    // parse only, nothing kept: what the host half costs (tools, timing)
		std::string e;
		const int r = vcf_stream(ref_path, vcf_path, window, threads, verbose, e, [](const std::vector<std::string> &) { return 0; },
		                         [](const LineBatch &, uint32_t) { return 0; });
		if (r) vfail(nullptr, r, e);
		return r;
	}
	*out = nullptr;
	ntsm_vcf_lines *L = new ntsm_vcf_lines();
	std::string err;
	const int rc = vcf_stream(
	    ref_path, vcf_path, window, threads, verbose, err,
	    [&](const std::vector<std::string> &ids) {
		    L->sample_ids = ids;
		    return 0;
	    },
	    [&](const LineBatch &b, uint32_t wstride) {
		    const size_t S = L->sample_ids.size(), gwords = (S + 15) / 16;
		    L->wstride = wstride;
		    L->windows.insert(L->windows.end(), b.windows.begin(), b.windows.begin() + (size_t)b.n * 2 * wstride);
		    L->lens.insert(L->lens.end(), b.lens.begin(), b.lens.begin() + (size_t)b.n * 2);
		    for (uint32_t l = 0; l < b.n; ++l)
			    for (size_t s = 0; s < S; ++s) L->genotypes.push_back((uint8_t)((b.geno2[l * gwords + (s >> 4)] >> (2 * (s & 15))) & 3u));
		    L->n += b.n;
		    return 0;
	    });
	// the lines in front of a fatal one stay available (upstream has inserted them by then); the code says how it ended
	*out = L;
	if (rc) vfail(nullptr, rc, err);
	return rc;
}

extern "C" void ntsm_vcf_lines_free(ntsm_vcf_lines *L) { delete L; }
extern "C" uint32_t ntsm_vcf_lines_n_samples(const ntsm_vcf_lines *L) { return L ? (uint32_t)L->sample_ids.size() : 0; }
extern "C" const char *ntsm_vcf_lines_sample_id(const ntsm_vcf_lines *L, uint32_t i) { return L && i < L->sample_ids.size() ? L->sample_ids[i].c_str() : nullptr; }
extern "C" uint64_t ntsm_vcf_lines_count(const ntsm_vcf_lines *L) { return L ? L->n : 0; }
extern "C" uint32_t ntsm_vcf_lines_wstride(const ntsm_vcf_lines *L) { return L ? L->wstride : 0; }
extern "C" const char *ntsm_vcf_lines_windows(const ntsm_vcf_lines *L) { return L ? L->windows.data() : nullptr; }
extern "C" const uint16_t *ntsm_vcf_lines_lens(const ntsm_vcf_lines *L) { return L ? L->lens.data() : nullptr; }
extern "C" const uint8_t *ntsm_vcf_lines_genotypes(const ntsm_vcf_lines *L) { return L ? L->genotypes.data() : nullptr; }

// ---- printers ----
namespace {

// ostream << double at the stream's precision, with the few distinct values a matrix holds formatted once
struct DoubleText {
	struct Slot { uint64_t bits; int prec; uint8_t len; char text[32]; };
	std::vector<Slot> slots = std::vector<Slot>(1024, Slot{ 0, 0, 0, { 0 } });
	void put(std::string &o, double v, int prec)
	{
		uint64_t b;
		memcpy(&b, &v, 8);
		Slot &s = slots[((b * 0x9E3779B97F4A7C15ull) >> 54) ^ (prec == 6 ? 0 : 512)];
		if (s.len == 0 || s.bits != b || s.prec != prec) {
			s.bits = b;
			s.prec = prec;
			s.len = (uint8_t)snprintf(s.text, sizeof s.text, "%.*g", prec, v);
		}
		o.append(s.text, s.len);
	}
};

}  // namespace

// MultiCount::printNormMatrix (src/MultiCount.hpp:148-203) into two files.  The numbers come from the device
// (ntsm_multi_norm_matrix); the text is made by `threads` workers, a block of rows each, and written in row order.
// The one piece of state the reference's stream carries from row to row -- setprecision(19) from the first missing
// value on (:192) -- is known up front: it is the position of the first UNDEF in the matrix.
extern "C" int ntsm_multi_write_norm_matrix(ntsm_multi *m, const ntsm_sites *sites, const char *const *sample_ids, const char *matrix_path,
                                            const char *center_path, uint32_t threads)
{
	if (!m || !sites || !matrix_path || !center_path) return NTSM_ERR_ARG;
	const uint32_t S = ntsm_sites_n_sites(sites), N = ntsm_multi_n_samples(m);
	if (N && !sample_ids) return NTSM_ERR_ARG;
	for (uint32_t j = 0; j < N; ++j)
		if (!sample_ids[j]) return NTSM_ERR_ARG;
	FILE *out = fopen(matrix_path, "wb");
	FILE *cf = fopen(center_path, "wb");
	if (!out || !cf) {
		if (out) fclose(out);
		if (cf) fclose(cf);
		ntsm_set_thread_error("cannot open the matrix / center file for writing");
		return NTSM_ERR_IO;
	}
	std::string o = "alleleID";                                    // :149-154
	for (uint32_t j = 0; j < N; ++j) {
		o += "\t";
		o += sample_ids[j];
	}
	o += "\n";
	fwrite(o.data(), 1, o.size(), out);
	int rc = NTSM_OK;
	if (ntsm_sites_printable(sites) != NTSM_OK) rc = NTSM_ERR_NOKEY;   // the first `at` on an erased k-mer / a missing var list throws (:158,:169)
	uint64_t sw = UINT64_MAX;                                      // row-major position of the first missing value: the precision switch
	if (rc == NTSM_OK && S) rc = ntsm_multi_norm_begin(m, &sw);
	if (rc == NTSM_OK && S) {
		// Three stages side by side: a fetcher copies the next block of rows off the device, `threads` workers turn the
		// current block into text (64 rows each), a writer puts the finished text into the two files in row order.
		const uint32_t T = std::max(1u, threads), rows_per_item = 64;
		const uint32_t block_rows = std::max<uint32_t>(rows_per_item, std::min<uint32_t>(S, (uint32_t)std::max<size_t>(rows_per_item * 2 * (size_t)T, (32u << 20) / ((size_t)N * 8 + 8))));
		const uint32_t n_blocks = (S + block_rows - 1) / block_rows;
		struct Block {
			std::vector<double> values, sums;
			std::vector<std::string> mtext, ctext;
			int state = 0;                                         // 0 free, 1 fetched, 2 formatted
		} blk[2];
		for (Block &b : blk) {
			b.values.resize((size_t)block_rows * N + 1);
			b.sums.resize(block_rows);
			ntsm_host_register(b.values.data(), b.values.size() * sizeof(double));   // page-locked: the copies run at the link's rate; best effort
		}
		std::mutex mu;
		std::condition_variable cv;
		int failed = NTSM_OK;
		std::thread fetcher([&] {
			for (uint32_t b = 0; b < n_blocks; ++b) {
				Block &B = blk[b & 1];
				{
					std::unique_lock<std::mutex> lk(mu);
					cv.wait(lk, [&] { return B.state == 0 || failed; });
					if (failed) return;
				}
				const uint32_t r0 = b * block_rows, nr = std::min(block_rows, S - r0);
				const int r = ntsm_multi_norm_fetch(m, r0, nr, B.values.data(), B.sums.data());
				std::lock_guard<std::mutex> lk(mu);
				if (r) failed = r;
				B.state = 1;
				cv.notify_all();
			}
		});
		std::thread writer([&] {
			for (uint32_t b = 0; b < n_blocks; ++b) {
				Block &B = blk[b & 1];
				{
					std::unique_lock<std::mutex> lk(mu);
					cv.wait(lk, [&] { return B.state == 2 || failed; });
					if (failed) return;
				}
				for (size_t i = 0; i < B.mtext.size(); ++i) {
					fwrite(B.mtext[i].data(), 1, B.mtext[i].size(), out);
					fwrite(B.ctext[i].data(), 1, B.ctext[i].size(), cf);
				}
				std::lock_guard<std::mutex> lk(mu);
				B.state = 0;
				cv.notify_all();
			}
		});
		for (uint32_t b = 0; b < n_blocks; ++b) {
			Block &B = blk[b & 1];
			{
				std::unique_lock<std::mutex> lk(mu);
				cv.wait(lk, [&] { return B.state == 1 || failed; });
				if (failed) break;
			}
			const uint32_t r0 = b * block_rows, nr = std::min(block_rows, S - r0);
			const size_t items = (nr + rows_per_item - 1) / rows_per_item;
			B.mtext.resize(items);
			B.ctext.resize(items);
			parallel_for(items, T, [&](size_t it) {
				std::string &mo = B.mtext[it], &co = B.ctext[it];
				mo.clear();
				co.clear();
				DoubleText cache;
				char num[64];
				const uint32_t l0 = (uint32_t)it * rows_per_item, l1 = std::min(nr, l0 + rows_per_item);
				for (uint32_t l = l0; l < l1; ++l) {
					const uint32_t i = r0 + l;
					mo += ntsm_sites_name(sites, i);               // :187
					const long double center = (long double)B.sums[l] / (long double)(uint64_t)N;   // :188-189: size counts every sample
					const int cl = snprintf(num, sizeof num, "%.19Lg", center);
					const double *row = B.values.data() + (size_t)l * N;
					for (uint32_t j = 0; j < N; ++j) {
						mo.push_back('\t');
						if (row[j] == 1.7976931348623157e308) mo.append(num, (size_t)cl);        // UNDEF (:41,190): the centre, 19 digits
						else cache.put(mo, row[j], (uint64_t)i * N + j > sw ? 19 : 6);          // :194 at the stream's precision
					}
					mo.push_back('\n');
					co.append(num, (size_t)cl);                    // :198
					co.push_back('\n');
				}
			});
			std::lock_guard<std::mutex> lk(mu);
			B.state = 2;
			cv.notify_all();
		}
		fetcher.join();
		writer.join();
		for (Block &b : blk) ntsm_host_unregister(b.values.data());
		if (failed) rc = failed;
	}
	ntsm_multi_norm_end(m);
	fclose(out);
	fclose(cf);
	return rc;
}

// MultiCount::printCountsMax(index) (:93-138): the counts file of one sample, without the #@TK / #@KS lines
extern "C" int64_t ntsm_multi_format_counts(ntsm_multi *m, const ntsm_sites *sites, uint32_t sample, char *buf, size_t cap)
{
	if (!m || !sites) return NTSM_ERR_ARG;
	const uint32_t S = ntsm_sites_n_sites(sites);
	std::vector<uint32_t> mr(S + 1), mv(S + 1), sr(S + 1), sv(S + 1);
	const int rc = ntsm_multi_counts_max(m, sample, mr.data(), mv.data(), sr.data(), sv.data());
	if (rc) return rc;
	// same rows as FingerPrint::printCountsMax: ntsm_format_counts's text minus its two header lines
	const int64_t need = ntsm_format_counts(sites, mr.data(), mv.data(), sr.data(), sv.data(), 0, nullptr, 0);
	if (need < 0) return need;
	std::string t((size_t)need, '\0');
	ntsm_format_counts(sites, mr.data(), mv.data(), sr.data(), sv.data(), 0, &t[0], t.size());
	const size_t cut = t.find("\n#locusID");                       // printCountsMax starts with that "\n" (:94)
	const size_t len = t.size() - cut;
	if (buf) memcpy(buf, t.data() + cut, std::min(cap, len));
	return (int64_t)len;
}

// VCFConvert::outputMatrix (:186-193)
extern "C" int ntsm_vcf_output_matrix(ntsm_vcf *v, const char *prefix)
{
	if (!v || !prefix) return NTSM_ERR_ARG;
	std::vector<const char *> ids;
	for (const std::string &s : v->sample_ids) ids.push_back(s.c_str());
	ids.push_back(nullptr);
	const std::string p(prefix);
	return ntsm_multi_write_norm_matrix(v->multi, v->sites, ids.data(), (p + "_matrix.tsv").c_str(), (p + "_center.txt").c_str(), v->threads);
}

// VCFConvert::outputCounts (:173-184): <dir>/<sampleID>.counts.txt for every sample (the reference writes into the
// working directory; dir == NULL or "" does the same)
extern "C" int ntsm_vcf_output_counts(ntsm_vcf *v, const char *dir)
{
	if (!v) return NTSM_ERR_ARG;
	const std::string d = dir && *dir ? std::string(dir) + "/" : std::string();
	for (uint32_t i = 0; i < v->sample_ids.size(); ++i) {
		const int64_t need = ntsm_multi_format_counts(v->multi, v->sites, i, nullptr, 0);
		if (need < 0) return (int)need;
		std::string t((size_t)need, '\0');
		ntsm_multi_format_counts(v->multi, v->sites, i, &t[0], t.size());
		FILE *fh = fopen((d + v->sample_ids[i] + ".counts.txt").c_str(), "wb");
		if (!fh) {
			ntsm_set_thread_error(("cannot write " + d + v->sample_ids[i] + ".counts.txt").c_str());
			return NTSM_ERR_IO;
		}
		fwrite(t.data(), 1, t.size(), fh);
		fclose(fh);
	}
	return NTSM_OK;
}

// ---- the ntsmVCF command line (src/ntSeqMatchVCF.cpp:53-217) ----
#define PROGRAM "ntsmVCF"

namespace {

size_t vcf_rss_kb()
{   // src/Util.h:31-50
	std::ifstream f("/proc/self/status");
	std::string line;
	while (std::getline(f, line))
		if (line.compare(0, 6, "VmRSS:") == 0) return (size_t)strtoull(line.c_str() + 6, nullptr, 10);
	return 0;
}

bool vcf_fexists(const std::string &p) { return std::ifstream(p.c_str()).good(); }   // src/Util.h:22

template <class T> bool vcf_parse(const char *arg, T &out)
{
	std::stringstream ss(arg ? arg : "");
	return (bool)(ss >> out);
}

const char kVcfHelp[] =
    "Usage: " PROGRAM " -s [FASTA] -r [FASTA] [VCF]\n"
    "Converts a multi vcf file to a set of counts files.\n"
    "Alternatively, you may also create a matrix to be used for PCA.\n"
    "  -t, --threads = INT    Number of threads to run.[1]\n"
    "  -d, --dupes            Allow shared k-mers between sites to\n"
    "                         be counted.\n"
    "  -s, --snp = STR        Interleaved fasta of SNP sites to\n"
    "                         k-merize. [required]\n"
    "  -p, --pca = STR        With multivcf generate rotation and\n"
    "                         centering files with this prefix.\n"
    "  -k, --kmer = INT       k-mer size used. [19]\n"
    "  -m, --multi = INT      Value to multiply base counts [20]\n"
    "  -w, --window = INT     Window size used. [31]\n"
    "  -r, --ref = STR        Reference fasta. [required]\n"
    "  -h, --help             Display this dialog.\n"
    "  -v, --verbose          Display verbose output.\n"
    "      --version          Print version information.\n";

}  // namespace

extern "C" int ntsm_vcf_main(int argc, char **argv)
{
	bool die = false;
	int opt_version = 0, opt_counts = 0, verbose = 0, device = 0;
	unsigned threads = 1, k = 19, multi = 20, window = 31;
	bool dupes = false;
	std::string snp, ref, pca;
	static struct option long_options[] = { { "threads", required_argument, nullptr, 't' }, { "dupes", no_argument, nullptr, 'd' },
		                                    { "snp", required_argument, nullptr, 's' },     { "pca", required_argument, nullptr, 'p' },
		                                    { "kmer", required_argument, nullptr, 'k' },    { "multi", required_argument, nullptr, 'm' },
		                                    { "window", required_argument, nullptr, 'w' },  { "ref", required_argument, nullptr, 'r' },
		                                    { "help", no_argument, nullptr, 'h' },          { "version", no_argument, &opt_version, 1 },
		                                    { "verbose", no_argument, nullptr, 'v' },
		                                    // additions (long options only): per-sample counts files (VCFConvert::outputCounts, which the
		                                    // reference's main never calls), and the GPU to use
		                                    { "counts", no_argument, &opt_counts, 1 },      { "device", required_argument, nullptr, 1000 },
		                                    { nullptr, 0, nullptr, 0 } };
	optind = 1;
	int c;
	while ((c = getopt_long(argc, argv, "s:t:vhk:dr:w:m:p:", long_options, nullptr)) != -1) {
		switch (c) {
		case 'h': std::cerr << kVcfHelp << std::endl; return 0;
		case 'd': dupes = true; break;
		case 'v': verbose++; break;
		case '?': die = true; break;
		case 0: break;
		default: {
			bool ok = true;
			if (c == 's') ok = vcf_parse(optarg, snp);
			else if (c == 'p') ok = vcf_parse(optarg, pca);
			else if (c == 'k') ok = vcf_parse(optarg, k);
			else if (c == 'w') ok = vcf_parse(optarg, window);
			else if (c == 'm') ok = vcf_parse(optarg, multi);
			else if (c == 't') ok = vcf_parse(optarg, threads);
			else if (c == 'r') ok = vcf_parse(optarg, ref);
			else if (c == 1000) ok = vcf_parse(optarg, device);
			if (!ok) {
				std::cerr << "Error - Invalid parameter " << (char)(c == 1000 ? 'D' : c) << ": " << optarg << std::endl;
				return 0;                                          // ntSeqMatchVCF.cpp:95,103,...: `return 0`
			}
		}
		}
	}
	if (opt_version) {
		std::cerr << PROGRAM " (ntsm) 663f9a5 -- ntsm_b200 matrix path\n"
		             "Written by Justin Chu <cjustin@ds.dfci.harvard.edu>\n\nCopyright 2020 Dana-Farber Cancer Institute\n"
		          << std::endl;
		return 0;
	}
	if (k > 32) {
		die = true;
		std::cerr << "k cannot be greater than 32" << std::endl;
	} else if (k == 32 || k == 0) {
		die = true;
		std::cerr << "k must be in 1..31 here (k = 32 is undefined behaviour upstream, vendor/KseqHashIterator.hpp:29)" << std::endl;
	}
	std::vector<std::string> inputs;
	while (optind < argc) {
		inputs.emplace_back(argv[optind++]);
		if (!vcf_fexists(inputs.back())) {                         // assert(Util::fexists(...)) (:172)
			std::cerr << PROGRAM ": " << inputs.back() << " does not exist" << std::endl;
			return 134;
		}
	}
	if (inputs.empty()) {
		std::cerr << "Error: Need Input File" << std::endl;
		die = true;
	}
	if (!vcf_fexists(ref)) {
		std::cerr << "Error: Unable to load reference file" << std::endl;
		die = true;
	}
	if (die) {
		std::cerr << "Try '--help' for more information.\n";
		return EXIT_FAILURE;
	}
	if (inputs.size() != 1) {                                       // assert(inputFiles.size() == 1) (:199)
		std::cerr << PROGRAM ": exactly one VCF file is expected" << std::endl;
		return 134;
	}
	const auto t0 = std::chrono::steady_clock::now();
	if (ntsm_device_count() == 0) {
		std::cerr << PROGRAM ": no CUDA device; this build has no CPU path" << std::endl;
		return 1;
	}
	// the GPU's context comes up while the site file is read
	std::thread warm([device] { ntsm_device_warmup(device); });
	ntsm_sites *sites = nullptr;
	int rc = ntsm_sites_load(&sites, snp.c_str(), k, dupes ? 1 : 0);
	warm.join();
	if (rc) {
		std::cerr << "file " << snp << " cannot be opened" << std::endl;   // MultiCount.hpp:219-222
		return 1;
	}
	for (uint32_t i = 0; i < ntsm_sites_n_warnings(sites); ++i) std::cerr << ntsm_sites_warning(sites, i) << std::endl;
	ntsm_cfg cfg{};
	cfg.k = k;
	cfg.device = device;
	cfg.n_buffers = 2;
	cfg.batch_bases = 4096;
	ntsm_ctx *ctx = nullptr;
	if ((rc = ntsm_ctx_create(&ctx, &cfg)) || (rc = ntsm_load_siteset(ctx, sites))) {
		std::cerr << PROGRAM ": " << ntsm_last_error(ctx) << std::endl;
		return 1;
	}
	ntsm_vcf *v = nullptr;
	rc = ntsm_vcf_convert(&v, ctx, sites, ref.c_str(), inputs[0].c_str(), multi, window, threads, verbose);
	const auto fail_exit = [&](int code) {
		if (code == NTSM_ERR_NOKEY) {
			std::cerr << "terminate called after throwing an instance of 'std::out_of_range'\n  what():  " << ntsm_last_error(ctx) << std::endl;
			return 134;
		}
		std::cerr << PROGRAM ": " << ntsm_last_error(ctx) << std::endl;
		return 1;
	};
	if (rc) return fail_exit(rc);
	{
		const int64_t wl = ntsm_multi_warnings_text(ntsm_vcf_multi(v), nullptr, 0);
		if (wl > 0) {
			std::string w((size_t)wl, '\0');
			ntsm_multi_warnings_text(ntsm_vcf_multi(v), &w[0], w.size());
			fwrite(w.data(), 1, w.size(), stderr);
		}
	}
	if (pca.empty()) {
		if (verbose > 1) std::cerr << "Outputting counts" << std::endl;   // :201-204 (and nothing else: the reference's main never calls outputCounts)
	} else {
		if (verbose > 1) std::cerr << "Outputting matrix and normalization values for PCA" << std::endl;
		if ((rc = ntsm_vcf_output_matrix(v, pca.c_str()))) {
			if (rc == NTSM_ERR_NOKEY) ntsm_ctx_set_error(ctx, "Couldn't find key.");
			return fail_exit(rc);
		}
	}
	if (opt_counts && (rc = ntsm_vcf_output_counts(v, nullptr))) {
		if (rc == NTSM_ERR_NOKEY) ntsm_ctx_set_error(ctx, "Couldn't find key.");
		return fail_exit(rc);
	}
	const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	std::cerr << "Time: " << secs << " s Memory: " << vcf_rss_kb() << " kbytes" << std::endl;   // :213-214
	ntsm_vcf_destroy(v);
	ntsm_ctx_destroy(ctx);
	ntsm_sites_free(sites);
	return 0;
}

// test knob: bytes per region when the VCF is not a plain regular file (default 64 MiB); returns the previous value
extern "C" uint64_t ntsm_vcf_stream_chunk(uint64_t bytes)
{
	const uint64_t old = g_stream_chunk;
	if (bytes) g_stream_chunk = (size_t)std::max<uint64_t>(bytes, 64);
	return old;
}

// which genotype decoder the VCF parser uses: 0 byte-wise, 1 AVX2 (8 columns a step), 2 AVX-512 (16); force >= 0 sets it
// (capped at what the CPU has), -1 = back to automatic, -2 = just ask
extern "C" int ntsm_vcf_genotype_isa(int force)
{
	if (force >= -1) g_gt_isa_forced = force;
	return gt_isa();
}
