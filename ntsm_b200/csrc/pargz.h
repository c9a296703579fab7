// pargz.h -- one DEFLATE stream inflated by several threads.
//
// A gzip'd FASTQ of a whole sequencing run is ONE member of tens of gigabytes, and the reference (like
// every gzread caller) inflates it on one core; with two such files and sixteen cores, fourteen sit
// idle.  DEFLATE has no index, but it can still be cut (the idea of pugz / rapidgzip):
//   * the compressed bytes are dealt out in chunks; the worker of chunk i looks for the first
//     position at or after its chunk's start where a dynamic-Huffman block header can begin (a cheap
//     bit test, then the real header parser, then the blocks really have to decode);
//   * it does not know the 32 KiB that precede that point, so it decodes into 16-bit symbols over a
//     window pre-filled with markers (256 + i = "byte i of the unknown window") and stops in front
//     of the first block header at or after the next chunk's start;
//   * the consumer takes the chunks in order.  Chunk i is accepted only if it started at the very bit
//     where chunk i-1 stopped -- then, by induction from the stream's true start, it began at a true
//     block boundary and decoded the true continuation.  Its markers are replaced from the last
//     32 KiB of real output, which also yields the window for chunk i+1.
// Nothing is guessed in the result: a chunk that does not line up (a false block start, a block
// longer than the search range, damaged data) ends the parallel phase at the last confirmed block
// boundary, and the caller carries on from that bit with the ordinary one-thread decoder (or, if the
// data really is damaged, hands the member to zlib as always).  CRC-32 and ISIZE of the member are
// checked by the caller like for any other member.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <memory>
#include <vector>

namespace ntsm {

class ParallelInflate {
public:
	enum End { kRunning, kStreamEnd, kBail };

	// [base, base + size): the mapped file; the raw DEFLATE stream starts at byte `deflate_start`
	ParallelInflate(const uint8_t *base, size_t size, size_t deflate_start, int workers, size_t chunk_bytes);
	~ParallelInflate();
	ParallelInflate(const ParallelInflate &) = delete;
	ParallelInflate &operator=(const ParallelInflate &) = delete;

	// The next run of decoded bytes, valid until the following call, and (if asked) their own CRC-32,
	// already computed by a worker.  false = no more from here: see end().
	bool next(const uint8_t **p, size_t *n, uint32_t *crc = nullptr);
	End end() const;
	// kStreamEnd: first byte after the DEFLATE stream (the gzip trailer).
	size_t end_byte() const;
	// kBail: the confirmed block boundary to resume at, and the (up to) 32 KiB of output before it
	uint64_t resume_bit() const;
	const std::vector<uint8_t> &window() const;
	uint64_t chunks_accepted() const;

private:
	struct Impl;
	std::unique_ptr<Impl> p_;
};

}  // namespace ntsm
