// gzsource.h -- the byte stream gzread() would deliver for a file, produced faster.
//
// The reference opens every input with gzopen and pulls it through gzread (src/FingerPrint.hpp:50,
// vendor/kseq.h:68-79): plain files pass through, gzip files are inflated member after member,
// trailing garbage after a member is ignored, an error ends the file.  For gzip'd FASTQ that call
// is the whole cost of ingest (cfg 4 of BASELINE.json), so GzSource
// produces the same bytes three ways:
//   zlib      gzread itself: pipes and other non-regular files, and the continuation of any
//             file the other modes gave up on;
//   mapped    a regular file that is not gzip at all (gzread's transparent mode): memory-mapped, the
//             reader scans the page cache's pages in place -- no read() copy;
//   fast      the file memory-mapped and decoded by ntsm::Inflater (inflate.h), CRC-32 and ISIZE
//             of every member checked before its last bytes are handed out;
//             With helper threads, a large single member is cut into chunks that workers inflate
//             speculatively from block starts they find themselves and that are accepted only if
//             each starts at the bit where the one before it stopped (pargz.h);
//   bgzf      a BGZF file (gzip members that carry their own size in a 'BC' extra field, as
//             bgzip / htslib and the Illumina converters write) inflated block-parallel by helper
//             threads -- the threads `-t` leaves idle when there are fewer files than threads,
//             because the reference only parallelises over files (src/FingerPrint.hpp:47-48).
// Whenever a member is not perfectly regular (header flags we do not parse, a decode error, a
// checksum mismatch, a block that is not BGZF after all) the source re-opens the file with zlib at
// that member's offset, skips the bytes of it that were already decoded, and carries on with zlib's
// inflate -- so which bytes of a damaged file count as decodable is zlib's verdict.  How many of
// them the reference then gets to see is reproduced as well: kseq pulls 16 KiB per gzread
// (vendor/kseq.h:229) and gzread drops the whole call in which a data error turns up, so only whole
// 16 KiB reads ahead of the error are released (a truncated file, by contrast, just ends).
#pragma once
#include <stdint.h>
#include <zlib.h>

#include <memory>
#include <string>

namespace ntsm {

class GzSource {
public:
	GzSource();
	~GzSource();
	GzSource(const GzSource &) = delete;
	GzSource &operator=(const GzSource &) = delete;

	// helpers: extra threads this source may start for block-parallel inflate (0 = none).
	// NTSM_INFLATE=zlib forces plain gzread (tests, comparisons).
	// map_plain: plain (non-gzip) regular files are memory-mapped rather than pulled through gzread
	bool open(const char *path, int helpers = 0, bool map_plain = true);
	void close();
	// gzread's contract: the number of bytes delivered, short only at the end of the input;
	// 0 at the end; -1 on a read/format error (after the bytes that preceded it were delivered)
	int read(void *dst, unsigned n);
	const char *mode() const;          // "zlib" | "fast" | "bgzf" | "mapped" -- what is producing bytes right now
	// "mapped": a plain (not gzip) regular file, memory-mapped; the whole input is [*base, *base + *size) and
	// a reader may scan it in place instead of calling read()
	bool mapped(const uint8_t **base, size_t *size) const;
	bool fell_back() const;            // a fast mode handed the file over to zlib
	bool bad() const;                  // a data error was met (read() returns -1 once the bytes before it are out)
	uint64_t parallel_chunks() const;  // chunks of single members that worker threads inflated and the stitcher accepted

private:
	struct Impl;
	std::unique_ptr<Impl> p_;
};

// CRC-32 (IEEE, as in gzip trailers) with carry-less multiplies when the CPU has them
uint32_t crc32_fast(uint32_t crc, const uint8_t *buf, size_t len);

}  // namespace ntsm
