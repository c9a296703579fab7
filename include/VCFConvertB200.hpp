// VCFConvertB200.hpp -- the binding a maintainer of the reference adds for the multi-sample matrix path: drop-ins
// for class MultiCount (src/MultiCount.hpp:36-289) and class VCFConvert (src/VCFConvert.hpp:40-218) backed by
// libntsm_b200.so, with the member functions ntsmVCF's main() calls (src/ntSeqMatchVCF.cpp:198-211: the
// constructor, count, outputMatrix) plus outputCounts and MultiCount's insertCount / printCountsMax /
// printNormMatrix -- same names, same argument meaning, same error behaviour (uncaught std::out_of_range where the
// reference throws one; an assert-style abort where its assert fails).
//
// Unlike the class it replaces, this VCFConvert builds its MultiCount once the sample IDs have been read from the
// VCF header: upstream builds it in the constructor, before any ID is known (src/VCFConvert.hpp:42), so its matrix
// has room for zero samples (src/MultiCount.hpp:266) and the shipped ntsmVCF crashes on its first insertCount.
//
// Use: compile the reference's untouched src/ntSeqMatchVCF.cpp with  -include VCFConvertB200.hpp  (this header claims
// the include guards of VCFConvert.hpp and MultiCount.hpp, so the reference's own classes are skipped), link with
// -lntsm_b200.  oracle/Makefile (target `dropin_vcf`) builds exactly that; tests/test_dropin.py runs the golden
// fixtures of tests/golden/vcf through the result.
#ifndef SRC_VCFCONVERT_HPP_
#define SRC_VCFCONVERT_HPP_
#ifndef SRC_MULTICOUNT_HPP_
#define SRC_MULTICOUNT_HPP_
#endif

#include <omp.h>
#include <stdint.h>
#include <stdlib.h>

#include <fstream>
#include <iostream>
#include <limits>
#include <stdexcept>
#include <string>
#include <vector>

#include "Options.h"
#include "ntsm_b200.h"

class MultiCount {
public:
	typedef uint32_t CountIndex;
	typedef uint64_t HashedKmer;
	static constexpr double UNDEF = std::numeric_limits<double>::max();      // :41

	MultiCount(const std::vector<std::string> &sampleIDs) : m_sampleIDs(sampleIDs) {    // :43-49 + initCountsHash :214-288
		if (ntsm_sites_load(&m_sites, opt::snp.c_str(), opt::k, opt::dupes)) {
			std::cerr << "file " << opt::snp << " cannot be opened" << std::endl;       // :219-222
			exit(1);
		}
		if (opt::verbose) std::cerr << "Opening " << opt::snp << std::endl;
		for (uint32_t i = 0; i < ntsm_sites_n_warnings(m_sites); ++i) std::cerr << ntsm_sites_warning(m_sites, i) << std::endl;
		ntsm_cfg cfg = ntsm_cfg();
		cfg.k = opt::k;
		cfg.n_buffers = 2;
		cfg.batch_bases = 4096;
		if (ntsm_ctx_create(&m_ctx, &cfg) || ntsm_load_siteset(m_ctx, m_sites) || ntsm_multi_create(&m_multi, m_ctx, (uint32_t)sampleIDs.size())) {
			std::cerr << "ntsmVCF: " << ntsm_last_error(m_ctx) << std::endl;            // no GPU: there is no CPU path to fall back to
			exit(1);
		}
		m_owned = true;
	}
	// over a matrix VCFConvert::count has filled (ntsm_vcf_multi)
	MultiCount(const std::vector<std::string> &sampleIDs, ntsm_sites *sites, ntsm_ctx *ctx, ntsm_multi *multi)
	    : m_sampleIDs(sampleIDs), m_sites(sites), m_ctx(ctx), m_multi(multi) {}
	~MultiCount() {
		if (!m_owned) return;
		ntsm_multi_destroy(m_multi);
		ntsm_ctx_destroy(m_ctx);
		ntsm_sites_free(m_sites);
	}
	MultiCount(const MultiCount &) = delete;
	MultiCount &operator=(const MultiCount &) = delete;

	void insertCount(unsigned sampleIndex, uint64_t hashVal, unsigned multi = 1) {      // :52-70
		const uint64_t before = ntsm_multi_n_warnings(m_multi);
		if (ntsm_multi_insert_count(m_multi, sampleIndex, hashVal, multi)) throw std::runtime_error(ntsm_last_error(m_ctx));
		if (ntsm_multi_n_warnings(m_multi) != before) flushWarnings();
	}

	void printCountsMax(unsigned index, std::ostream &out = std::cout) const {          // :93-138
		const int64_t n = ntsm_multi_format_counts(m_multi, m_sites, index, NULL, 0);
		if (n == NTSM_ERR_NOKEY) throw std::out_of_range("Couldn't find key.");
		if (n < 0) throw std::runtime_error(ntsm_last_error(m_ctx));
		std::string s((size_t)n, '\0');
		ntsm_multi_format_counts(m_multi, m_sites, index, &s[0], s.size());
		out << s;
	}

	// :148-203; the reference takes two streams -- these are the two files VCFConvert::outputMatrix opens on them
	void printNormMatrix(const std::string &matrixPath, const std::string &centerPath) {
		std::vector<const char *> ids;
		for (size_t i = 0; i < m_sampleIDs.size(); ++i) ids.push_back(m_sampleIDs[i].c_str());
		ids.push_back(NULL);
		const int rc = ntsm_multi_write_norm_matrix(m_multi, m_sites, ids.data(), matrixPath.c_str(), centerPath.c_str(),
		                                            opt::threads < 1 ? 1 : opt::threads);
		if (rc == NTSM_ERR_NOKEY) throw std::out_of_range("Couldn't find key.");
		if (rc) throw std::runtime_error(ntsm_last_error(m_ctx));
	}

	// the "Inconsistent k-mer counts" warnings raised since the last call (:59-62), to stderr as upstream
	void flushWarnings() {
		const int64_t n = ntsm_multi_warnings_text(m_multi, NULL, 0);
		if (n <= (int64_t)m_warned) return;
		std::string s((size_t)n, '\0');
		ntsm_multi_warnings_text(m_multi, &s[0], s.size());
		std::cerr.write(s.data() + m_warned, n - (int64_t)m_warned);
		m_warned = (size_t)n;
	}

private:
	const std::vector<std::string> &m_sampleIDs;
	ntsm_sites *m_sites = NULL;
	ntsm_ctx *m_ctx = NULL;
	ntsm_multi *m_multi = NULL;
	bool m_owned = false;
	size_t m_warned = 0;
};

class VCFConvert {
public:
	VCFConvert() {                                                           // :42-59 (the genome is read by count())
		if (opt::verbose > 1) std::cerr << "Loading Reference " << opt::ref << std::endl;
		if (ntsm_sites_load(&m_sites, opt::snp.c_str(), opt::k, opt::dupes)) {
			std::cerr << "file " << opt::snp << " cannot be opened" << std::endl;
			exit(1);
		}
		if (opt::verbose) std::cerr << "Opening " << opt::snp << std::endl;
		for (uint32_t i = 0; i < ntsm_sites_n_warnings(m_sites); ++i) std::cerr << ntsm_sites_warning(m_sites, i) << std::endl;
		ntsm_cfg cfg = ntsm_cfg();
		cfg.k = opt::k;
		cfg.n_buffers = 2;
		cfg.batch_bases = 4096;
		if (ntsm_ctx_create(&m_ctx, &cfg) || ntsm_load_siteset(m_ctx, m_sites)) {
			std::cerr << "ntsmVCF: " << ntsm_last_error(m_ctx) << std::endl;
			exit(1);
		}
	}
	~VCFConvert() {
		delete m_counts;
		ntsm_vcf_destroy(m_vcf);
		ntsm_ctx_destroy(m_ctx);
		ntsm_sites_free(m_sites);
	}
	VCFConvert(const VCFConvert &) = delete;
	VCFConvert &operator=(const VCFConvert &) = delete;

	void count(std::string filename) {                                       // :62-174
		if (opt::verbose > 1) std::cerr << "Reading VCF file: " << filename << std::endl;
		const int rc = ntsm_vcf_convert(&m_vcf, m_ctx, m_sites, opt::ref.c_str(), filename.c_str(), opt::multi, opt::window,
		                                opt::threads < 1 ? 1 : opt::threads, opt::verbose > 2 ? opt::verbose : 0);
		if (rc == NTSM_ERR_NOKEY) throw std::out_of_range(ntsm_last_error(m_ctx));   // where upstream dies on an exception / its assert
		if (rc) {
			std::cerr << "ntsmVCF: " << ntsm_last_error(m_ctx) << std::endl;
			exit(1);
		}
		for (uint32_t i = 0; i < ntsm_vcf_n_samples(m_vcf); ++i) m_sampleIDs.push_back(ntsm_vcf_sample_id(m_vcf, i));
		if (opt::verbose > 1) std::cerr << "Starting multicount of each rsID for " << m_sampleIDs.size() << " samples." << std::endl;
		m_counts = new MultiCount(m_sampleIDs, m_sites, m_ctx, ntsm_vcf_multi(m_vcf));
		m_counts->flushWarnings();
	}

	void outputCounts() {                                                    // :173-184
		for (unsigned i = 0; i < m_sampleIDs.size(); ++i) {
			if (opt::verbose > 1) std::cerr << "Outputting counts for " << m_sampleIDs.at(i) << std::endl;
			std::ofstream out((m_sampleIDs.at(i) + ".counts.txt").c_str());
			m_counts->printCountsMax(i, out);
			out.close();
		}
	}

	void outputMatrix(const std::string &prefix) {                           // :186-193
		if (opt::verbose > 1) std::cerr << "Outputting matrix and normalization values for PCA" << std::endl;
		m_counts->printNormMatrix(prefix + "_matrix.tsv", prefix + "_center.txt");
	}

private:
	std::vector<std::string> m_sampleIDs;
	ntsm_sites *m_sites = NULL;
	ntsm_ctx *m_ctx = NULL;
	ntsm_vcf *m_vcf = NULL;
	MultiCount *m_counts = NULL;
};

#endif /* SRC_VCFCONVERT_HPP_ */
