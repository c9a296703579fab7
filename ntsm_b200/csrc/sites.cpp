// sites.cpp -- host side of FingerPrint::initCountsHash (src/FingerPrint.hpp:490-564), laid out
// the way MultiCount keeps it (src/MultiCount.hpp:208-209,247,268): every listed k-mer gets a
// dense index in file order and the per-site lists become CSR offsets into one flat array.
#include <sched.h>
#include <stdio.h>
#include <string.h>

#include <stdlib.h>

#include <algorithm>
#include <atomic>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "../../include/ntsm_b200.h"
#include "fastx.h"
#include "kmer_math.h"

namespace {

// Exact "first occurrence wins" over all k-mer occurrences of the site file, in parallel.
//
// The reference inserts k-mers one by one, in file order, into a hash map (src/FingerPrint.hpp:
// 516-548): the first occurrence of a hash value owns it, every later one is a collision.  The same
// partition falls out of an order-free formulation: every occurrence g (numbered in file order)
// writes min(g) into the slot of its key; afterwards occurrence g is a first occurrence iff the
// slot holds g.  That is two embarrassingly parallel passes over a lock-free open-addressing table
// (64-bit CAS to claim a key, CAS-min on the value), which matters for panels like BASELINE's
// config 5 (26 M k-mers: the serial map took most of the CLI's wall time).
class MinIndexTable {
public:
	explicit MinIndexTable(uint64_t n_keys)
	{
		cap_ = 1024;
		while (cap_ < 2 * n_keys) cap_ <<= 1;
		keys_.reset(new std::atomic<uint64_t>[cap_]);
		vals_.reset(new std::atomic<uint32_t>[cap_]);
	}
	void clear_range(uint64_t lo, uint64_t hi)
	{
		for (uint64_t i = lo; i < hi; ++i) {
			keys_[i].store(kEmptyKey, std::memory_order_relaxed);
			vals_[i].store(0xFFFFFFFFu, std::memory_order_relaxed);
		}
	}
	uint64_t capacity() const { return cap_; }
	void put_min(uint64_t key, uint32_t g)
	{
		uint64_t i = slot(key);
		for (;;) {
			uint64_t k = keys_[i].load(std::memory_order_acquire);
			if (k == kEmptyKey && keys_[i].compare_exchange_strong(k, key, std::memory_order_acq_rel)) k = key;
			if (k == key) {
				uint32_t cur = vals_[i].load(std::memory_order_relaxed);
				while (g < cur && !vals_[i].compare_exchange_weak(cur, g, std::memory_order_relaxed)) {}
				return;
			}
			i = (i + 1) & (cap_ - 1);
		}
	}
	uint32_t get(uint64_t key) const
	{
		uint64_t i = slot(key);
		while (keys_[i].load(std::memory_order_relaxed) != key) i = (i + 1) & (cap_ - 1);
		return vals_[i].load(std::memory_order_relaxed);
	}

private:
	static constexpr uint64_t kEmptyKey = ~0ull;          // hash64 values have at most 62 bits
	uint64_t slot(uint64_t k) const
	{
		k ^= k >> 31;
		k *= 0x9E3779B97F4A7C15ULL;
		return (k >> 20) & (cap_ - 1);
	}
	uint64_t cap_ = 0;
	std::unique_ptr<std::atomic<uint64_t>[]> keys_;
	std::unique_ptr<std::atomic<uint32_t>[]> vals_;
};

template <class F> void parallel_for(unsigned n_threads, uint64_t n, F &&body)   // body(thread, lo, hi)
{
	if (n_threads <= 1 || n < 2) {
		body(0u, (uint64_t)0, n);
		return;
	}
	std::vector<std::thread> th;
	for (unsigned t = 0; t < n_threads; ++t)
		th.emplace_back([&, t] { body(t, n * t / n_threads, n * (t + 1) / n_threads); });
	for (auto &x : th) x.join();
}

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
	const uint64_t m = ntsm::kmer_mask(k);
	const unsigned shift = 2 * (k - 1);

	// 1. the records (kseq grammar), sequences and names kept back to back                       :508
	std::string seqs, names;
	std::vector<uint64_t> seq_off(1, 0), name_off(1, 0);
	int64_t l;
	while ((l = rd.next()) >= 0) {
		seqs.append(rd.seq(), (size_t)l);
		seq_off.push_back(seqs.size());
		names.append(rd.name());
		name_off.push_back(names.size());
	}
	rd.close();
	const uint64_t n_rec = seq_off.size() - 1;
	s->n_records = (uint32_t)n_rec;

	// the cores this process may actually use (a cgroup / taskset can leave far fewer than the machine has)
	unsigned avail = std::max(1u, std::thread::hardware_concurrency());
	{
		cpu_set_t set;
		CPU_ZERO(&set);
		if (sched_getaffinity(0, sizeof set, &set) == 0 && CPU_COUNT(&set) > 0) avail = std::min<unsigned>(avail, (unsigned)CPU_COUNT(&set));
	}
	unsigned n_threads = std::min(16u, avail);
	if (seqs.size() < (1u << 20)) n_threads = 1;           // small panels: not worth starting threads
	if (const char *e = getenv("NTSM_SITES_THREADS")) n_threads = (unsigned)std::min(256, std::max(1, atoi(e)));

	// 2. every k-mer occurrence of every record (KseqHashIterator :95-112 + hash64), records split
	//    over the threads in contiguous ranges: (hash, 1-based end position) in file order
	std::vector<std::vector<uint64_t>> t_hash(n_threads);
	std::vector<std::vector<uint32_t>> t_pos(n_threads);
	std::vector<uint32_t> rec_occ(n_rec + 1, 0);            // occurrences per record, then their prefix sum
	parallel_for(n_threads, n_rec, [&](unsigned t, uint64_t lo, uint64_t hi) {
		std::vector<uint64_t> &H = t_hash[t];
		std::vector<uint32_t> &P = t_pos[t];
		for (uint64_t r = lo; r < hi; ++r) {
			const char *seq = seqs.data() + seq_off[r];
			const uint64_t len = seq_off[r + 1] - seq_off[r];
			uint64_t fw = 0, rv = 0;
			unsigned run = 0;
			uint32_t cnt = 0;
			for (uint64_t p = 0; p < len; ++p) {
				const unsigned c = ntsm::nt4((unsigned char)seq[p]);
				if (c < 4) {
					fw = ((fw << 2) | c) & m;
					rv = (rv >> 2) | ((uint64_t)(3 - c) << shift);
					if (++run >= k) {
						H.push_back(ntsm::hash64(fw < rv ? fw : rv, m));
						P.push_back((uint32_t)(p + 1));
						++cnt;
					}
				} else {
					fw = rv = 0;
					run = 0;
				}
			}
			rec_occ[r + 1] = cnt;
		}
	});
	for (uint64_t r = 0; r < n_rec; ++r) rec_occ[r + 1] += rec_occ[r];
	const uint64_t n_occ = rec_occ[n_rec];
	if (n_occ >= 0xFFFFFFFFull) {
		delete s;
		return NTSM_ERR_ARG;
	}
	std::vector<uint64_t> occ_hash(n_occ);
	std::vector<uint32_t> occ_pos(n_occ);
	{
		uint64_t at = 0;
		for (unsigned t = 0; t < n_threads; ++t) {         // thread t's records precede thread t+1's
			if (!t_hash[t].empty()) {
				memcpy(occ_hash.data() + at, t_hash[t].data(), t_hash[t].size() * 8);
				memcpy(occ_pos.data() + at, t_pos[t].data(), t_pos[t].size() * 4);
			}
			at += t_hash[t].size();
			std::vector<uint64_t>().swap(t_hash[t]);
			std::vector<uint32_t>().swap(t_pos[t]);
		}
	}

	// 3. first occurrence of every hash value: min over occurrence numbers, then read back
	std::vector<uint32_t> first_of(n_occ);
	{
		MinIndexTable tab(n_occ);
		parallel_for(n_threads, tab.capacity(), [&](unsigned, uint64_t lo, uint64_t hi) { tab.clear_range(lo, hi); });
		parallel_for(n_threads, n_occ, [&](unsigned, uint64_t lo, uint64_t hi) {
			for (uint64_t g = lo; g < hi; ++g) tab.put_min(occ_hash[g], (uint32_t)g);
		});
		parallel_for(n_threads, n_occ, [&](unsigned, uint64_t lo, uint64_t hi) {
			for (uint64_t g = lo; g < hi; ++g) first_of[g] = tab.get(occ_hash[g]);
		});
	}

	// 4. file order again: dense indices for first occurrences, a warning and a dupe entry for the rest
	std::vector<uint32_t> dupes;
	std::vector<uint32_t> dense_of(n_occ);
	uint64_t n_first = 0;
	for (uint64_t g = 0; g < n_occ; ++g) n_first += first_of[g] == g;
	s->hashes.reserve(n_first);
	s->allele_off.reserve(n_rec + 2);
	s->names.reserve(n_rec / 2 + 1);
	for (uint64_t r = 0; r < n_rec; ++r) {
		const bool is_ref = (r % 2) == 0;                            // :510
		s->allele_off.push_back((uint32_t)s->hashes.size());
		const std::string name(names.data() + name_off[r], names.data() + name_off[r + 1]);
		for (uint64_t g = rec_occ[r]; g < rec_occ[r + 1]; ++g) {
			if (first_of[g] == g) {                                  // :526-527 / :547-548
				dense_of[g] = (uint32_t)s->hashes.size();
				s->hashes.push_back(occ_hash[g]);
			} else {                                                 // :520-524 / :541-545
				char w[512];
				snprintf(w, sizeof w, "Warning: %s of %s file has a k-mer collision at pos: %llu", name.c_str(),
				         is_ref ? "REF" : "VAR", (unsigned long long)occ_pos[g]);
				s->warnings.emplace_back(w);
				dupes.push_back(dense_of[first_of[g]]);
			}
		}
		if (is_ref) s->names.push_back(name);                        // :530
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
