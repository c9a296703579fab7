// cli.cpp -- ntsm_main: the ntsmCount command line, option for option
// (src/ntSeqMatchCount.cpp:53-184, src/Options.h:20-61), driving the GPU path through the C ABI.
// Additions (long options only, nothing the reference parses changes meaning):
//   --gpus N          use N GPUs of this box (default 1; 0 = one per 16 parser threads)
//   --batch-bases B   stream positions per pinned batch (default 2^25; 2^22 with -m)
//   --dump-kmer-counts F    also write every k-mer's counter + the tallies to F (merge.cpp's format)
//   --merge-kmer-counts     FILES are such dumps of shards of ONE sample: add them per k-mer, then print
//                           the counts file -- the exact merge `ntsmEval --merge` is not
//                           (src/CompareCounts.hpp:648-657 sums per-site maxima)
#include <getopt.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "../../include/ntsm_b200.h"
#include "internal.h"

#define PROGRAM "ntsmCount"

namespace {

size_t rss_kb()
{   // src/Util.h:31-50
	std::ifstream f("/proc/self/status");
	std::string line;
	while (std::getline(f, line))
		if (line.compare(0, 6, "VmRSS:") == 0) return (size_t)strtoull(line.c_str() + 6, nullptr, 10);
	return 0;
}

bool fexists(const std::string &p) { return std::ifstream(p.c_str()).good(); }   // src/Util.h:22

template <class T> bool parse(const char *arg, T &out)
{   // the reference converts every option value through a stringstream (ntSeqMatchCount.cpp:84-125)
	std::stringstream ss(arg ? arg : "");
	return (bool)(ss >> out);
}

const char kHelp[] =
    "Usage: " PROGRAM " -s [FASTA] [OPTION]... [FILES...]\n"
    "  -t, --threads = INT    Number of threads to run.[1]\n"
    "  -m, --maxCov = INT     k-mer coverage threshold for early\n"
    "                         termination. [inf]\n"
    "  -o, --output = STR     Output for summary file.\n"
    "  -d, --dupes            Allow shared k-mers between sites to\n"
    "                         be counted.\n"
    "  -s, --snp = STR        Interleaved fasta of SNP sites to\n"
    "                         k-merize. [required]\n"
    "  -k, --kmer = INT       k-mer size used. [19]\n"
    "  -h, --help             Display this dialog.\n"
    "  -v, --verbose          Display verbose output.\n"
    "      --version          Print version information.\n"
    "      --gpus = INT       GPUs of this box to use (0 = one per 16\n"
    "                         threads, as many as -t can feed). [1]\n"
    "      --batch-bases = INT  positions per pinned batch. [2^25]\n"
    "      --dump-kmer-counts = STR  also write k-mer level counts to\n"
    "                         this file (for --merge-kmer-counts).\n"
    "      --merge-kmer-counts  FILES are k-mer count dumps of shards\n"
    "                         of one sample: merge them exactly.\n";

}  // namespace

extern "C" int ntsm_main(int argc, char **argv)
{
	// opt:: globals of src/Options.h, local here
	int verbose = 0;
	unsigned threads = 1, k = 19;
	std::string snp, summary;
	double covThresh = 0;          // 0 = no cap (reference default DBL_MAX never triggers, Options.h:29)
	bool dupes = false, die = false;
	int opt_version = 0;
	int gpus = 1;
	unsigned long long batch_bases = 0;
	std::string dump_path;
	int merge_mode = 0;

	static struct option long_options[] = {
	    {"threads", required_argument, NULL, 't'}, {"maxCov", required_argument, NULL, 'm'},
	    {"output", required_argument, NULL, 'o'},  {"dupes", required_argument, NULL, 'd'},
	    {"snp", required_argument, NULL, 's'},     {"kmer", required_argument, NULL, 'k'},
	    {"help", no_argument, NULL, 'h'},          {"version", no_argument, &opt_version, 1},
	    {"verbose", no_argument, NULL, 'v'},       {"gpus", required_argument, NULL, 1001},
	    {"batch-bases", required_argument, NULL, 1002}, {"dump-kmer-counts", required_argument, NULL, 1003},
	    {"merge-kmer-counts", no_argument, &merge_mode, 1}, {NULL, 0, NULL, 0}};
	optind = 1;
	int c, option_index = 0;
	while ((c = getopt_long(argc, argv, "s:t:vhk:m:do:", long_options, &option_index)) != -1) {
		switch (c) {
		case 'h': std::cerr << kHelp << std::endl; return EXIT_SUCCESS;
		case 'o': if (!parse(optarg, summary)) { std::cerr << "Error - Invalid parameter o: " << optarg << std::endl; return 0; } break;
		case 'd': dupes = true; break;
		case 's': if (!parse(optarg, snp)) { std::cerr << "Error - Invalid parameter s: " << optarg << std::endl; return 0; } break;
		case 'm': if (!parse(optarg, covThresh)) { std::cerr << "Error - Invalid parameter m: " << optarg << std::endl; return 0; } break;
		case 'k': if (!parse(optarg, k)) { std::cerr << "Error - Invalid parameter k: " << optarg << std::endl; return 0; } break;
		case 't': if (!parse(optarg, threads)) { std::cerr << "Error - Invalid parameter t: " << optarg << std::endl; return 0; } break;
		case 'v': verbose++; break;
		case 1001: if (!parse(optarg, gpus)) { std::cerr << "Error - Invalid parameter gpus: " << optarg << std::endl; return 0; } break;
		case 1002: if (!parse(optarg, batch_bases)) { std::cerr << "Error - Invalid parameter batch-bases: " << optarg << std::endl; return 0; } break;
		case 1003: if (!parse(optarg, dump_path)) { std::cerr << "Error - Invalid parameter dump-kmer-counts: " << optarg << std::endl; return 0; } break;
		case '?': die = true; break;
		}
	}
	if (opt_version) {
		std::cerr << PROGRAM " (ntsm-b200) " << ntsm_version() << "\n" << std::endl;
		return EXIT_SUCCESS;
	}
	if (k > 32) { die = true; std::cerr << "Error: k cannot be greater than 32" << std::endl; }
	else if (k == 32 || k < 1) { die = true; std::cerr << "Error: k must be between 1 and 31 (k=32 overflows the reference's 64-bit mask)" << std::endl; }
	if (snp.empty()) { die = true; std::cerr << "Error: Missing variants (-s) file" << std::endl; }
	std::vector<std::string> inputFiles;
	while (optind < argc) {
		inputFiles.emplace_back(argv[optind]);
		if (!fexists(inputFiles.back())) {                      // assert(Util::fexists(...)), ntSeqMatchCount.cpp:160
			std::cerr << PROGRAM ": input file " << inputFiles.back() << " does not exist" << std::endl;
			return 134;
		}
		optind++;
	}
	if (inputFiles.empty()) { std::cerr << "Error: Need input files" << std::endl; die = true; }
	if (die) { std::cerr << "Try '--help' for more information.\n"; return EXIT_FAILURE; }

	const auto t0 = std::chrono::steady_clock::now();
	// phase timing on stderr when NTSM_TIMING is set (measurement aid, not part of the reference's output)
	const bool timing = getenv("NTSM_TIMING") != nullptr;
	auto lap = [&, last = t0](const char *what) mutable {
		const auto now = std::chrono::steady_clock::now();
		if (timing) fprintf(stderr, "[ntsm timing] %-28s %.3f s\n", what, std::chrono::duration<double>(now - last).count());
		last = now;
	};

	// the CUDA primary contexts come up on their own threads while this one reads the site file
	std::vector<std::thread> warm;
	{
		const int nd = ntsm_device_count();
		// --gpus 0 = as many as the parser threads can feed: file ingest is host-bound at ~2-4 Gbases/s per parser
		// thread, one GPU takes 144 over PCIe, and every further CUDA context costs ~0.7 s to bring up (8 GPUs: 5.7 s)
		if (gpus <= 0) gpus = (int)std::min<unsigned>((unsigned)std::max(1, nd), std::max(1u, (threads + 15) / 16));
		const int ng = gpus > nd ? nd : gpus;
		for (int g = 0; g < ng; ++g) warm.emplace_back([g] { ntsm_device_warmup(g); });
	}
	struct Joiner {
		std::vector<std::thread> &t;
		~Joiner() { for (auto &x : t) if (x.joinable()) x.join(); }
	} joiner{warm};

	// FingerPrint fp;  (ntSeqMatchCount.cpp:177)
	ntsm_sites *sites = nullptr;
	int rc = ntsm_sites_load(&sites, snp.c_str(), k, dupes);
	if (rc) { std::cerr << "file " << snp << " cannot be opened" << std::endl; return 1; }   // FingerPrint.hpp:493-499
	if (verbose) std::cerr << "Opening " << snp << std::endl;
	for (uint32_t i = 0; i < ntsm_sites_n_warnings(sites); ++i) std::cerr << ntsm_sites_warning(sites, i) << std::endl;
	lap("sites.fa -> site table (host)");

	const int n_dev = ntsm_device_count();
	if (n_dev <= 0) {
		std::cerr << PROGRAM ": no CUDA device found; this build has no CPU path" << std::endl;
		return 1;
	}
	if (gpus <= 0 || gpus > n_dev) gpus = n_dev;            // (0 was resolved above; more than the box has = all)
	ntsm_cfg cfg;
	memset(&cfg, 0, sizeof cfg);
	cfg.k = k;
	cfg.max_counts = ntsm_sites_max_counts(sites, covThresh);
	cfg.batch_bases = batch_bases ? batch_bases : (cfg.max_counts ? (1ull << 22) : 0);
	cfg.n_buffers = 2 + (threads < 1 ? 1 : threads);
	for (auto &x : warm) x.join();
	std::vector<ntsm_ctx *> ctxs((size_t)gpus, nullptr);
	{
		// one thread per GPU: pinned ring, streams, table upload + build_tables_kernel
		std::vector<std::thread> th;
		std::vector<int> rcs((size_t)gpus, 0);
		std::vector<std::string> errs((size_t)gpus);
		for (int g = 0; g < gpus; ++g)
			th.emplace_back([&, g] {
				ntsm_cfg c = cfg;
				c.device = g;
				if ((rcs[g] = ntsm_ctx_create(&ctxs[g], &c)) == 0) rcs[g] = ntsm_load_siteset(ctxs[g], sites);
				if (rcs[g]) errs[g] = ntsm_last_error(ctxs[g]);      // the text of a failed create lives in this thread
			});
		for (auto &t : th) t.join();
		for (int g = 0; g < gpus; ++g)
			if (rcs[g]) {
				std::cerr << PROGRAM ": " << errs[g] << std::endl;
				return 1;
			}
	}
	lap("CUDA context + device tables");
	// fp.computeCounts(inputFiles);  (:178)
	std::vector<const char *> paths;
	for (auto &f : inputFiles) paths.push_back(f.c_str());
	int early = 0;
	if (merge_mode) {
		// FILES are k-mer count dumps of shards of one sample: summed per k-mer on GPU 0, the per-site max comes after
		for (const char *p : paths)
			if ((rc = ntsm_counts_load_add(ctxs[0], sites, p))) {
				std::cerr << PROGRAM ": " << (rc == NTSM_ERR_IO ? ntsm_last_error(nullptr) : ntsm_last_error(ctxs[0])) << std::endl;
				return 1;
			}
		lap("merge k-mer count files");
	} else {
		rc = ntsm_count_files(ctxs.data(), (uint32_t)gpus, paths.data(), (uint32_t)paths.size(), threads, verbose, &early);
		if (rc) { std::cerr << ntsm_last_error(nullptr) << std::endl; return 1; }
		if (early) {
			if (verbose > 0) {
				// FingerPrint.hpp:477-484; m_totalReads only moves under -vvv (:70-72), so -v / -vv print 0 reads there too.
				// (The "Current Total:" progress lines of -vvv, racy upstream, are not reproduced.)
				uint64_t t[3] = {0, 0, 0}, reads = 0;
				for (ntsm_ctx *x : ctxs) {
					uint64_t u[3] = {0, 0, 0};
					ntsm_get_totals(x, u);
					t[0] += u[0]; t[1] += u[1]; t[2] += u[2];
					reads += ntsm_ctx_reads(x);
				}
				// (upstream bumps its read counter after the cap check, :67-72: the deciding read is not in it yet)
				std::cerr << "max count reached at " << (verbose > 2 && reads ? reads - 1 : 0) << " reads, " << t[0] << " k-mers, " << t[1]
				          << " total counts, and " << t[2] << " total bases " << std::endl;
			}
			std::cerr << "Reached desired (-m) threshold" << std::endl;   // FingerPrint.hpp:84-86
		}
		lap("parse + pack + count");
	}

	// combine the GPUs and fetch the rows: GPU 0 sums every GPU's private k-mer counts out of peer
	// memory over NVLink inside the per-site reduce kernel (sums first, max afterwards) -- no communicator
	const uint32_t S = ntsm_sites_n_sites(sites);
	std::vector<uint32_t> mr(S), mv(S), sr(S), sv(S);
	uint64_t totals[3] = {0, 0, 0};
	rc = ntsm_group_finalize(ctxs.data(), (uint32_t)gpus, mr.data(), mv.data(), sr.data(), sv.data(), totals);
	if (rc) { std::cerr << PROGRAM ": " << ntsm_last_error(ctxs[0]) << std::endl; return 1; }
	if (!dump_path.empty() && (rc = ntsm_counts_save(ctxs[0], sites, dump_path.c_str()))) {     // GPU 0 holds the combined k-mer counts now
		std::cerr << PROGRAM ": " << (rc == NTSM_ERR_IO ? ntsm_last_error(nullptr) : ntsm_last_error(ctxs[0])) << std::endl;
		return 1;
	}

	// fp.printOptionalHeader(); fp.printCountsMax();  (:179-180)
	const int64_t need = ntsm_format_counts(sites, mr.data(), mv.data(), sr.data(), sv.data(), totals[0], nullptr, 0);
	if (need < 0) {
		// the reference throws std::out_of_range from m_counts.at() here and dies with SIGABRT
		std::cerr << "terminate called after throwing an instance of 'std::out_of_range'\n  what():  Couldn't find key." << std::endl;
		return 134;
	}
	std::string text((size_t)need, '\0');
	ntsm_format_counts(sites, mr.data(), mv.data(), sr.data(), sv.data(), totals[0], &text[0], text.size());
	fwrite(text.data(), 1, text.size(), stdout);
	fflush(stdout);
	lap("combine + rows + counts file");

	// cerr << fp.printInfoSummary() << endl;  (:181, FingerPrint.hpp:313-349)
	const uint32_t covered = ntsm_sites_covered(mr.data(), mv.data(), S);
	char sum[1024];
	const int64_t sl = ntsm_format_summary(sites, totals, covered, sum, sizeof sum);
	const std::string sumtext(sum, (size_t)sl);
	if (!summary.empty()) { std::ofstream fh(summary.c_str()); fh << sumtext; }
	const double covPer = double(covered) / double(S);
	if (covPer < 0.75f)
		std::cerr << "Warning: site coverage is : " << covPer
		          << "(<75%). Data may be sorted or sparse along the genome. Any PCA projection may be inaccurate." << std::endl;
	std::cerr << sumtext << std::endl;
	const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	std::cerr << "Time: " << secs << " s Memory: " << rss_kb() << " kbytes" << std::endl;   // ntSeqMatchCount.cpp:182

	for (ntsm_ctx *x : ctxs) ntsm_ctx_destroy(x);
	ntsm_sites_free(sites);
	return 0;
}
