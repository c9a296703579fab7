// tools/ref_iter_dump.cpp -- golden-vector generator (NOT product code, NOT shipped).
// Compiles against the reference header where it lies (-I/root/reference) and dumps,
// for every line "k<TAB>sequence" on stdin, the k-mers its iterator yields:
//     k <TAB> seq <TAB> pos:hash,pos:hash,...
// Used only by tools/make_golden.py to produce tests/golden/iter_vectors.tsv.
#include <cstdint>
#include <cstdio>
#include <iostream>
#include <sstream>
#include <string>
#include "vendor/KseqHashIterator.hpp"

int main() {
	std::string line;
	while (std::getline(std::cin, line)) {
		size_t tab = line.find('\t');
		if (tab == std::string::npos) continue;
		int k = std::stoi(line.substr(0, tab));
		std::string seq = line.substr(tab + 1);
		// sequences are given with \xNN escapes so raw bytes 0..3 can be exercised
		std::string raw;
		for (size_t i = 0; i < seq.size(); ++i) {
			if (seq[i] == '\\' && i + 3 < seq.size() && seq[i + 1] == 'x') {
				raw.push_back((char) std::stoi(seq.substr(i + 2, 2), nullptr, 16));
				i += 3;
			} else raw.push_back(seq[i]);
		}
		std::cout << k << '\t' << seq << '\t';
		bool first = true;
		for (KseqHashIterator it(raw.data(), raw.size(), k); it != it.end(); ++it) {
			if (!first) std::cout << ',';
			first = false;
			std::cout << it.getPos() << ':' << std::hex << *it << std::dec;
		}
		std::cout << '\n';
	}
	return 0;
}
