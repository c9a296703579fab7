// sites.cpp -- host side of FingerPrint::initCountsHash (src/FingerPrint.hpp:490-564), laid out
// the way MultiCount keeps it (src/MultiCount.hpp:208-209,247,268): every listed k-mer gets a
// dense index in file order and the per-site lists become CSR offsets into one flat array.
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/ntsm_b200.h"
#include "fastx.h"
#include "kmer_math.h"

namespace {

// exact set of 64-bit keys -> dense index (open addressing, grows by doubling)
class KeyIndex {
public:
	KeyIndex() { resize(1u << 16); }
	// returns the existing index, or inserts `idx` and returns UINT32_MAX
	uint32_t find_or_insert(uint64_t key, uint32_t idx)
	{
		if ((size_ + 1) * 2 > keys_.size()) resize(keys_.size() * 2);
		size_t i = slot(key);
		while (vals_[i] != kEmpty) {
			if (keys_[i] == key) return vals_[i];
			i = (i + 1) & (keys_.size() - 1);
		}
		keys_[i] = key;
		vals_[i] = idx;
		++size_;
		return kEmpty;
	}
	size_t size() const { return size_; }

private:
	static constexpr uint32_t kEmpty = 0xFFFFFFFFu;
	size_t slot(uint64_t k) const
	{
		k ^= k >> 31;
		k *= 0x9E3779B97F4A7C15ULL;
		return (size_t)(k >> 20) & (keys_.size() - 1);
	}
	void resize(size_t cap)
	{
		std::vector<uint64_t> ok;
		std::vector<uint32_t> ov;
		ok.swap(keys_);
		ov.swap(vals_);
		keys_.assign(cap, 0);
		vals_.assign(cap, kEmpty);
		size_ = 0;
		for (size_t i = 0; i < ok.size(); ++i)
			if (ov[i] != kEmpty) find_or_insert(ok[i], ov[i]);
	}
	std::vector<uint64_t> keys_;
	std::vector<uint32_t> vals_;
	size_t size_ = 0;
};

}  // namespace

struct ntsm_sites {
	uint32_t k = 19;
	bool allow_dupes = false;
	std::vector<uint64_t> hashes;        // dense k-mer index -> hash64 value
	std::vector<uint8_t> erased;         // duplicate erased from the table (no -d)
	std::vector<uint32_t> allele_off;    // 2*n_sites+1
	std::vector<std::string> names;      // m_alleleIDs
	std::vector<std::string> warnings;
	uint32_t n_records = 0;
	uint64_t table_size = 0;
};

extern "C" int ntsm_sites_load(ntsm_sites **out, const char *path, uint32_t k, int allow_dupes)
{
	if (!out || !path || k < 1 || k > 31) return NTSM_ERR_ARG;
	ntsm::FastxReader rd;
	if (!rd.open(path)) return NTSM_ERR_IO;
	ntsm_sites *s = new ntsm_sites();
	s->k = k;
	s->allow_dupes = allow_dupes != 0;
	KeyIndex index;
	std::vector<uint32_t> dupes;
	const uint64_t m = ntsm::kmer_mask(k);
	const unsigned shift = 2 * (k - 1);
	int64_t l;
	while ((l = rd.next()) >= 0) {                                   // :508
		const bool is_ref = (s->n_records % 2) == 0;                 // :510
		s->allele_off.push_back((uint32_t)s->hashes.size());
		const char *seq = rd.seq();
		uint64_t fw = 0, rv = 0;
		unsigned run = 0;
		for (int64_t p = 0; p < l; ++p) {                            // KseqHashIterator :95-112
			const unsigned c = ntsm::nt4((unsigned char)seq[p]);
			if (c < 4) {
				fw = ((fw << 2) | c) & m;
				rv = (rv >> 2) | ((uint64_t)(3 - c) << shift);
				if (++run >= k) {
					const uint64_t hv = ntsm::hash64(fw < rv ? fw : rv, m);
					const uint32_t prev = index.find_or_insert(hv, (uint32_t)s->hashes.size());
					if (prev != 0xFFFFFFFFu) {                       // :520-524 / :541-545
						char w[512];
						snprintf(w, sizeof w, "Warning: %s of %s file has a k-mer collision at pos: %llu", rd.name(),
						         is_ref ? "REF" : "VAR", (unsigned long long)(p + 1));
						s->warnings.emplace_back(w);
						dupes.push_back(prev);
					} else {                                         // :526-527 / :547-548
						s->hashes.push_back(hv);
					}
				}
			} else {
				fw = rv = 0;
				run = 0;
			}
		}
		if (is_ref) s->names.emplace_back(rd.name());                // :530
		s->n_records++;
	}
	s->allele_off.push_back((uint32_t)s->hashes.size());
	if (s->n_records % 2) s->allele_off.push_back((uint32_t)s->hashes.size());   // absent var list = empty
	s->erased.assign(s->hashes.size(), 0);
	s->table_size = s->hashes.size();
	if (!s->allow_dupes)                                             // :557-563
		for (uint32_t d : dupes)
			if (!s->erased[d]) {
				s->erased[d] = 1;
				s->table_size--;
			}
	*out = s;
	return NTSM_OK;
}

extern "C" void ntsm_sites_free(ntsm_sites *s) { delete s; }
extern "C" uint32_t ntsm_sites_k(const ntsm_sites *s) { return s->k; }
extern "C" uint32_t ntsm_sites_n_sites(const ntsm_sites *s) { return (uint32_t)s->names.size(); }
extern "C" uint32_t ntsm_sites_n_kmers(const ntsm_sites *s) { return (uint32_t)s->hashes.size(); }
extern "C" uint64_t ntsm_sites_table_size(const ntsm_sites *s) { return s->table_size; }
extern "C" const uint64_t *ntsm_sites_hashes(const ntsm_sites *s) { return s->hashes.data(); }
extern "C" const uint32_t *ntsm_sites_allele_off(const ntsm_sites *s) { return s->allele_off.data(); }
extern "C" const uint8_t *ntsm_sites_erased(const ntsm_sites *s) { return s->erased.data(); }
extern "C" const char *ntsm_sites_name(const ntsm_sites *s, uint32_t i) { return s->names[i].c_str(); }
extern "C" uint32_t ntsm_sites_n_warnings(const ntsm_sites *s) { return (uint32_t)s->warnings.size(); }
extern "C" const char *ntsm_sites_warning(const ntsm_sites *s, uint32_t i) { return s->warnings[i].c_str(); }

extern "C" int ntsm_sites_printable(const ntsm_sites *s)
{
	if (s->n_records % 2) return NTSM_ERR_NOKEY;                     // m_alleleIDToKmerVar.at(i) throws, :276
	for (uint8_t e : s->erased)
		if (e) return NTSM_ERR_NOKEY;                                // m_counts.at(hv) throws, :282/:289
	return NTSM_OK;
}

extern "C" uint64_t ntsm_sites_max_counts(const ntsm_sites *s, double cov)
{
	if (!(cov > 0)) return 0;                                        // -m 0 disables (:41); default never triggers
	const double v = ((double)s->table_size * cov) / 2;              // :42
	if (v >= 18446744073709551615.0) return 0;
	return (uint64_t)v;
}

extern "C" uint32_t ntsm_sites_covered(const uint32_t *max_ref, const uint32_t *max_var, uint32_t n_sites)
{
	uint32_t n = 0;                                                  // :389-413
	for (uint32_t i = 0; i < n_sites; ++i) n += (max_ref[i] > 0 || max_var[i] > 0);
	return n;
}

extern "C" int64_t ntsm_format_counts(const ntsm_sites *s, const uint32_t *max_ref, const uint32_t *max_var,
                                      const uint32_t *sum_ref, const uint32_t *sum_var, uint64_t total_kmers,
                                      char *buf, size_t cap)
{
	// printOptionalHeader :261-268 + printCountsMax :270-311
	std::string o;
	o.reserve(64 + (size_t)s->names.size() * 48);
	o += "#@TK\t" + std::to_string(total_kmers) + "\n#@KS\t" + std::to_string(s->k);
	o += "\n#locusID\tcountAT\tcountCG\tsumAT\tsumCG\tdistinctAT\tdistinctCG\n";
	int rc = 0;
	char row[160];
	for (size_t i = 0; i < s->names.size(); ++i) {
		const uint32_t r0 = s->allele_off[2 * i], r1 = s->allele_off[2 * i + 1], v1 = s->allele_off[2 * i + 2];
		bool bad = (s->n_records % 2) && i + 1 == s->names.size();
		for (uint32_t j = r0; j < v1 && !bad; ++j) bad = s->erased[j];
		if (bad) { rc = NTSM_ERR_NOKEY; break; }                     // the reference dies here, rows so far were written
		const int n = snprintf(row, sizeof row, "\t%u\t%u\t%u\t%u\t%u\t%u\n", max_ref[i], max_var[i], sum_ref[i],
		                       sum_var[i], r1 - r0, v1 - r1);
		o += s->names[i];
		o.append(row, (size_t)n);
	}
	if (buf) {
		const size_t n = o.size() < cap ? o.size() : cap;
		memcpy(buf, o.data(), n);
	}
	return rc ? rc : (int64_t)o.size();
}

extern "C" int64_t ntsm_format_summary(const ntsm_sites *s, const uint64_t totals[3], uint32_t covered, char *buf,
                                       size_t cap)
{
	// printInfoSummary :313-333
	std::string o;
	o += "Total Bases Considered: " + std::to_string(totals[2]) + "\n";
	o += "Total k-mers Considered: " + std::to_string(totals[0]) + "\n";
	o += "Total k-mers Recorded: " + std::to_string(totals[1]) + "\n";
	o += "Distinct k-mers in initial set: " + std::to_string(s->table_size) + "\n";
	o += "Total Sites: " + std::to_string(s->names.size()) + "\n";
	o += "Sites Covered by at least one k-mer: " + std::to_string(covered) + "\n";
	if (buf) {
		const size_t n = o.size() < cap ? o.size() : cap;
		memcpy(buf, o.data(), n);
	}
	return (int64_t)o.size();
}

extern "C" uint64_t ntsm_hash64(uint64_t key, uint32_t k) { return ntsm::hash64(key, ntsm::kmer_mask(k)); }
extern "C" uint64_t ntsm_hash64_inv(uint64_t h, uint32_t k) { return ntsm::hash64_inv(h, ntsm::kmer_mask(k)); }
