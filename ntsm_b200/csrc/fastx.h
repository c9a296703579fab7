// fastx.h -- FASTA/FASTQ(.gz) record reader with the exact record grammar of kseq_read
// (vendor/kseq.h:178-219) as FingerPrint::computeCounts drives it (src/FingerPrint.hpp:64-67).
//
// Same grammar, different machinery: a 1 MiB window scanned with memchr instead of a 16 KiB buffer
// walked byte by byte, filled by GzSource (gzsource.h: the byte stream gzread would deliver).  Behaviour that must match (all covered by tests/golden):
//   * a record starts at the next '>' or '@' (anywhere, when hunting; at a line start otherwise);
//   * name = header up to the first isspace(); rest of the header line is ignored;
//   * sequence = concatenation of the following lines until a line STARTING with '>', '+' or '@';
//     empty lines are skipped; after each appended line one trailing '\r' is dropped if the
//     sequence so far is longer than 1 byte; every other byte is kept verbatim;
//   * '+' starts the quality block: rest of that line skipped, then whole lines are appended
//     (same '\r' rule) until qual is at least as long as seq; length mismatch -> -2;
//   * end of input after the '+' line -> -2; a record cut off inside its sequence is returned
//     as a FASTA-style record;
//   * return codes: >=0 sequence length, -1 end of file, -2 bad quality, -3 read error.
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

#include "gzsource.h"

namespace ntsm {

class FastxReader {
public:
	FastxReader() = default;
	~FastxReader() { close(); }
	FastxReader(const FastxReader &) = delete;
	FastxReader &operator=(const FastxReader &) = delete;

	// helpers: idle threads the byte source may use for block-parallel inflate (gzsource.h)
	// map_plain: scan plain files in place through a mapping (faster per reader, does not scale past ~8 readers per process)
	bool open(const char *path, int helpers = 0, bool map_plain = true);
	const char *source_mode() const { return src_.mode(); }
	void close();
	// next record; sequence available through seq()/name() until the following call
	int64_t next();
	const char *seq() const { return cur_seq_; }
	uint64_t seq_len() const { return cur_len_; }
	const char *name();

private:
	// Fast path for the record shape every sequencer writes -- "@name\nSEQ\n+...\nQUAL\n" with one
	// sequence line and a quality line of the same length, wholly inside the window: four memchr
	// calls, nothing copied (seq() points into the window).  Returns false, with no state changed,
	// for anything else; the byte-exact general parser below then handles the record.
	bool next_fast(int64_t *len);
	int64_t next_general();
	bool fill();                       // refill the window; false at end of input / error
	int getc();                        // next byte, -1 end, -3 error
	// append the rest of the current line to dst (without '\n'); returns false if nothing could
	// be read because the input had already ended
	bool take_line(std::vector<char> &dst);
	bool skip_line();

	GzSource src_;
	bool open_ = false;
	std::vector<unsigned char> own_;   // the 1 MiB window when the source has to be read()
	unsigned char *buf_ = nullptr;     // window base: own_.data(), or the source's own mapping (read-only then: nothing writes to it)
	size_t cap_ = 0;                   // window capacity
	size_t beg_ = 0, end_ = 0;
	bool mapped_ = false;              // the window IS the whole input (GzSource "mapped")
	size_t populated_ = 0;             // mapped: page tables are set up for [0, populated_)
	void populate_ahead();
	bool eof_ = false, err_ = false, src_err_ = false;
	int last_ = 0;                     // header byte already consumed by the previous record
	std::string name_;
	std::vector<char> seq_, qual_;
	const char *cur_seq_ = "";         // sequence of the current record: in seq_ or inside the window
	uint64_t cur_len_ = 0;
	const unsigned char *fast_name_ = nullptr;   // header text of a fast-path record (name built on demand)
	size_t fast_name_len_ = 0;
};

void fastx_rescan_isa();               // re-read NTSM_SCAN_ISA (tests)
const char *fastx_scan_isa();          // "avx512" | "avx2" | "memchr": the line scanner of the FASTQ fast path

}  // namespace ntsm
