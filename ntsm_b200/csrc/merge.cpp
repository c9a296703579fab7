// merge.cpp -- k-mer-level count files: the EXACT way to combine shards of one sample.
//
// The reference combines shard results with `ntsmEval --merge`, which adds up the shards' counts FILES
// (src/CompareCounts.hpp:626-674): countAT = sum over shards of max over the site's k-mers (:648-657).
// That is not what counting the concatenated input gives (max over k-mers of the summed counts):
// max-of-sums <= sum-of-maxes, with equality only when the same k-mer leads in every shard.
// A k-mer count file keeps every k-mer's counter (m_counts' values in site-list order, SURVEY A13's
// flat layout) plus the three tallies; merging adds the files per k-mer on the GPU and only then
// runs the per-site reduce (site_reduce_kernel), so merge(shards) == count(all reads), byte for byte.
//
// Format (little endian): "NTSMKC1\n", u32 k, u32 n_kmers, u64 digest of the site set (FNV-1a over the
// listed hash64 values and the CSR offsets: a file only merges into the panel it was counted
// against), u64 totals[3] = {TK, hits, bases}, u32 counts[n_kmers].
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/ntsm_b200.h"
#include "internal.h"

namespace {

const char kMagic[8] = { 'N', 'T', 'S', 'M', 'K', 'C', '1', '\n' };

uint64_t fnv1a(uint64_t h, const void *p, size_t n)
{
	const unsigned char *b = static_cast<const unsigned char *>(p);
	for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 0x100000001B3ull; }
	return h;
}

uint64_t site_digest(const ntsm_sites *s)
{
	uint64_t h = 0xCBF29CE484222325ull;
	const uint32_t k = ntsm_sites_k(s), nk = ntsm_sites_n_kmers(s), ns = ntsm_sites_n_sites(s);
	h = fnv1a(h, &k, 4);
	h = fnv1a(h, ntsm_sites_hashes(s), (size_t)nk * 8);
	h = fnv1a(h, ntsm_sites_allele_off(s), (2 * (size_t)ns + 1) * 4);
	return h;
}

int io_fail(const char *what, const char *path)
{
	ntsm_set_thread_error((std::string(what) + " " + path).c_str());
	return NTSM_ERR_IO;
}

}  // namespace

extern "C" int ntsm_counts_save(ntsm_ctx *ctx, const ntsm_sites *s, const char *path)
{
	if (!ctx || !s || !path) return NTSM_ERR_ARG;
	const uint32_t k = ntsm_sites_k(s), nk = ntsm_sites_n_kmers(s);
	std::vector<uint32_t> counts(nk ? nk : 1);
	uint64_t totals[3] = { 0, 0, 0 };
	int rc = ntsm_get_counts(ctx, counts.data());
	if (rc) return rc;
	if ((rc = ntsm_get_totals(ctx, totals))) return rc;
	FILE *f = fopen(path, "wb");
	if (!f) return io_fail("cannot write", path);
	const uint64_t dg = site_digest(s);
	bool ok = fwrite(kMagic, 1, 8, f) == 8 && fwrite(&k, 4, 1, f) == 1 && fwrite(&nk, 4, 1, f) == 1 && fwrite(&dg, 8, 1, f) == 1 &&
	          fwrite(totals, 8, 3, f) == 3 && (nk == 0 || fwrite(counts.data(), 4, nk, f) == nk);
	ok = (fclose(f) == 0) && ok;
	return ok ? NTSM_OK : io_fail("short write to", path);
}

extern "C" int ntsm_counts_load_add(ntsm_ctx *ctx, const ntsm_sites *s, const char *path)
{
	if (!ctx || !s || !path) return NTSM_ERR_ARG;
	FILE *f = fopen(path, "rb");
	if (!f) return io_fail("cannot open", path);
	char magic[8];
	uint32_t k = 0, nk = 0;
	uint64_t dg = 0, totals[3] = { 0, 0, 0 };
	bool ok = fread(magic, 1, 8, f) == 8 && !memcmp(magic, kMagic, 8) && fread(&k, 4, 1, f) == 1 && fread(&nk, 4, 1, f) == 1 &&
	          fread(&dg, 8, 1, f) == 1 && fread(totals, 8, 3, f) == 3;
	if (!ok) { fclose(f); return io_fail("not a k-mer count file:", path); }
	if (k != ntsm_sites_k(s) || nk != ntsm_sites_n_kmers(s) || dg != site_digest(s)) {
		fclose(f);
		return io_fail("k-mer count file was made with another site set or k:", path);
	}
	std::vector<uint32_t> counts(nk ? nk : 1);
	ok = nk == 0 || fread(counts.data(), 4, nk, f) == nk;
	ok = ok && fgetc(f) == EOF;
	fclose(f);
	if (!ok) return io_fail("truncated or oversized k-mer count file:", path);
	return ntsm_add_counts(ctx, counts.data(), totals);
}
