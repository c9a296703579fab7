// tools/ref_vcf_harness.cpp -- golden-vector generator for the multi-sample matrix path
// (MultiCount::insertCount / printCountsMax / printNormMatrix driven by
// VCFConvert::count).  NOT product code, NOT shipped.
//
// Why a harness and not the reference's own ntsmVCF binary: at 663f9a5 that tool cannot run.
// VCFConvert's constructor builds its MultiCount member while m_sampleIDs is still empty
// (src/VCFConvert.hpp:42), so MultiCount::initCountsHash sizes m_matCounts for ZERO samples
// (src/MultiCount.hpp:266) and the first insertCount writes out of bounds
// (tools/ref_ntsmvcf_repro.sh shows the SIGSEGV).  This harness compiles the reference's two
// classes UNMODIFIED from where they lie (-I/root/reference), constructs VCFConvert exactly as
// ntSeqMatchVCF.cpp:198 does, and then does the one thing the constructor evidently meant to do:
// sizes m_matCounts as (listed k-mers) x (samples named on the VCF's #CHROM line), the expression
// of MultiCount.hpp:266 evaluated with the sample IDs known.  Everything after that --
// VCFConvert::count, MultiCount::insertCount, printCountsMax, printNormMatrix -- is the
// reference's own code running on its own data structures.
//
//   usage: ref_vcf_harness <sites.fa> <ref.fa> <in.vcf> <out_prefix> [k] [multi] [window] [dupes 0|1] [threads] [matrix_only 0|1]
//   writes <out_prefix>_matrix.tsv, <out_prefix>_center.txt (VCFConvert::outputMatrix),
//          <out_prefix>_counts_<j>.txt for every sample j (MultiCount::printCountsMax(j)),
//          <out_prefix>_mat.bin (raw m_matCounts) ; stderr = the reference's warnings
#include <cassert>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <iterator>
#include <limits>
#include <memory>
#include <sstream>
#include <string>
#include <vector>
#include <math.h>
#include <omp.h>
#include <zlib.h>

#define private public
#include "src/Options.h"
#include "src/VCFConvert.hpp"
#undef private

static size_t samples_on_header(const char *path)
{
	// same walk as src/VCFConvert.hpp:71-93
	std::ifstream fh(path);
	std::string line;
	while (getline(fh, line)) {
		if (line.empty() || line.at(0) != '#') continue;
		std::stringstream ss(line);
		std::string item;
		getline(ss, item, '\t');
		if (item.compare("#CHROM") == 0) {
			for (unsigned i = 0; i < 8; ++i) getline(ss, item, '\t');
			size_t n = 0;
			while (getline(ss, item, '\t')) ++n;
			return n;
		}
	}
	return 0;
}

int main(int argc, char **argv)
{
	if (argc < 5) {
		fprintf(stderr, "usage: %s sites.fa ref.fa in.vcf out_prefix [k] [multi] [window] [dupes]\n", argv[0]);
		return 2;
	}
	opt::snp = argv[1];
	opt::ref = argv[2];
	const std::string vcf = argv[3], prefix = argv[4];
	if (argc > 5) opt::k = atoi(argv[5]);
	if (argc > 6) opt::multi = atoi(argv[6]);
	if (argc > 7) opt::window = atoi(argv[7]);
	if (argc > 8) opt::dupes = atoi(argv[8]) != 0;
	opt::threads = argc > 9 ? atoi(argv[9]) : 1;   // 1 = the reference's only deterministic mode (first writer of a cell wins, MultiCount.hpp:52-69)
	omp_set_num_threads(opt::threads);             // ntSeqMatchVCF.cpp:158-161; more threads only for timing (tools/bench_matrix.py)
	const bool matrix_only = argc > 10 && atoi(argv[10]) != 0;

	VCFConvert convert;          // ntSeqMatchVCF.cpp:198
	size_t listed = 0;           // kmerCount of MultiCount.hpp:217: every k-mer that entered a site list
	for (size_t i = 0; i < convert.m_counts.m_alleleIDToKmerRef.size(); ++i) listed += convert.m_counts.m_alleleIDToKmerRef[i]->size();
	for (size_t i = 0; i < convert.m_counts.m_alleleIDToKmerVar.size(); ++i) listed += convert.m_counts.m_alleleIDToKmerVar[i]->size();
	convert.m_counts.m_matCounts = std::vector<uint8_t>(listed * samples_on_header(vcf.c_str()), 0);   // MultiCount.hpp:266 with the IDs known

	double t0 = omp_get_wtime();
	convert.count(vcf);          // ntSeqMatchVCF.cpp:200
	double t1 = omp_get_wtime();
	if (matrix_only) {           // what ntsmVCF -p does and nothing else, timed (ntSeqMatchVCF.cpp:200-211)
		convert.outputMatrix(prefix);
		fprintf(stderr, "harness_seconds count %.3f outputMatrix %.3f\n", t1 - t0, omp_get_wtime() - t1);
		return 0;
	}
	{
		std::ofstream mat((prefix + "_mat.bin").c_str(), std::ios::binary);
		mat.write((const char *)convert.m_counts.m_matCounts.data(), convert.m_counts.m_matCounts.size());
	}
	for (unsigned j = 0; j < convert.m_sampleIDs.size(); ++j) {   // VCFConvert::outputCounts, one file per sample index
		std::ofstream out((prefix + "_counts_" + std::to_string(j) + ".txt").c_str());
		convert.m_counts.printCountsMax(j, out);
	}
	convert.outputMatrix(prefix);   // ntSeqMatchVCF.cpp:210
	return 0;
}
