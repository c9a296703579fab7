// fastx.cpp -- see fastx.h.  Grammar per vendor/kseq.h:178-219; return codes per :171-176.
#include "fastx.h"

#include <immintrin.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>

#include "../../include/ntsm_b200.h"

namespace ntsm {

static constexpr size_t kWindow = 1u << 20;

bool FastxReader::open(const char *path, int helpers, bool map_plain)
{
	close();
	if (!src_.open(path, helpers, map_plain)) return false;   // plain files pass through, like the reference's gzopen (FingerPrint.hpp:50)
	open_ = true;
	beg_ = end_ = 0;
	eof_ = err_ = src_err_ = false;
	const uint8_t *mbase = nullptr;
	size_t msize = 0;
	mapped_ = src_.mapped(&mbase, &msize);
	populated_ = 0;
	if (mapped_) {
		// the whole file is the window: nothing is ever read() or moved, the scanners walk the page cache
		buf_ = const_cast<unsigned char *>(mbase);
		cap_ = end_ = msize;
		eof_ = true;
		populate_ahead();
	} else {
		own_.resize(kWindow);
		buf_ = own_.data();
		cap_ = own_.size();
	}
	last_ = 0;
	fast_name_ = nullptr;
	cur_seq_ = "";
	cur_len_ = 0;
	return true;
}

void FastxReader::close()
{
	if (open_) src_.close();
	open_ = false;
}

// Mapped input: set the page tables up for the next stretch in one call (MADV_POPULATE_READ, Linux 5.14+)
// instead of taking a fault per 64 KiB while scanning; where the kernel does not know the advice the
// faults simply happen as the scanners touch the pages.
void FastxReader::populate_ahead()
{
#ifndef MADV_POPULATE_READ
#define MADV_POPULATE_READ 22
#endif
	constexpr size_t kStretch = 32u << 20;
	if (!mapped_ || populated_ >= end_) return;
	const size_t from = populated_, len = end_ - from < kStretch ? end_ - from : kStretch;
	madvise(buf_ + from, len, MADV_POPULATE_READ);         // failure (old kernel) costs nothing: plain faults take over
	populated_ = from + len;
}

bool FastxReader::fill()
{
	if (beg_ < end_) return true;
	if (eof_) {
		if (src_err_) err_ = true;      // a read error met while topping the window up surfaces once the window is drained
		return false;
	}
	beg_ = 0;
	const int n = src_.read(buf_, (unsigned)cap_);
	if (n <= 0) {
		eof_ = true;
		err_ = n < 0;
		end_ = 0;
		return false;
	}
	end_ = (size_t)n;
	return true;
}

int FastxReader::getc()
{
	if (err_) return -3;
	if (!fill()) return err_ ? -3 : -1;
	return buf_[beg_++];
}

bool FastxReader::take_line(std::vector<char> &dst)
{
	bool got = false;
	for (;;) {
		if (!fill()) break;
		got = true;
		const unsigned char *p = buf_ + beg_;
		const size_t avail = end_ - beg_;
		const unsigned char *nl = (const unsigned char *)memchr(p, '\n', avail);
		const size_t n = nl ? (size_t)(nl - p) : avail;
		dst.insert(dst.end(), (const char *)p, (const char *)p + n);
		beg_ += n + (nl ? 1 : 0);
		if (nl) break;
	}
	if (!got) return false;
	if (dst.size() > 1 && dst.back() == '\r') dst.pop_back();   // kseq.h:141
	return true;
}

bool FastxReader::skip_line()
{
	for (;;) {
		if (!fill()) return false;
		const unsigned char *p = buf_ + beg_;
		const unsigned char *nl = (const unsigned char *)memchr(p, '\n', end_ - beg_);
		if (nl) { beg_ = (size_t)(nl - buf_) + 1; return true; }
		beg_ = end_;
	}
}

const char *FastxReader::name()
{
	if (fast_name_) {                                      // name = header up to the first white-space byte
		size_t i = 0;
		while (i < fast_name_len_ && !(fast_name_[i] == ' ' || (fast_name_[i] >= '\t' && fast_name_[i] <= '\r'))) ++i;
		name_.assign((const char *)fast_name_, i);
		fast_name_ = nullptr;
	}
	return name_.c_str();
}

// The next (up to) four '\n' positions at or after p, before e.  A FASTQ record is four lines, and four
// memchr calls per ~300-byte record were most of the fast path's time (call overhead, not bytes);
// one pass of 64-byte compares finds all four.
namespace {

__attribute__((target("avx512f,avx512bw,bmi"))) int newlines4_avx512(const unsigned char *p, const unsigned char *e, const unsigned char **out)
{
	int found = 0;
	const __m512i nl = _mm512_set1_epi8('\n');
	while (p + 64 <= e) {
		uint64_t m = _mm512_cmpeq_epi8_mask(_mm512_loadu_si512((const void *)p), nl);
		while (m) {
			out[found++] = p + __builtin_ctzll(m);
			if (found == 4) return 4;
			m &= m - 1;
		}
		p += 64;
	}
	if (p < e) {
		const __mmask64 keep = ~0ull >> (64 - (e - p));
		uint64_t m = _mm512_cmpeq_epi8_mask(_mm512_maskz_loadu_epi8(keep, (const void *)p), nl) & keep;
		while (m) {
			out[found++] = p + __builtin_ctzll(m);
			if (found == 4) return 4;
			m &= m - 1;
		}
	}
	return found;
}

__attribute__((target("avx2,bmi"))) int newlines4_avx2(const unsigned char *p, const unsigned char *e, const unsigned char **out)
{
	int found = 0;
	const __m256i nl = _mm256_set1_epi8('\n');
	while (p + 32 <= e) {
		uint32_t m = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(_mm256_loadu_si256((const __m256i *)p), nl));
		while (m) {
			out[found++] = p + __builtin_ctz(m);
			if (found == 4) return 4;
			m &= m - 1;
		}
		p += 32;
	}
	while (p < e && found < 4) {
		const unsigned char *q = (const unsigned char *)memchr(p, '\n', (size_t)(e - p));
		if (!q) break;
		out[found++] = q;
		p = q + 1;
	}
	return found;
}

int newlines4_memchr(const unsigned char *p, const unsigned char *e, const unsigned char **out)
{
	int found = 0;
	while (p < e && found < 4) {
		const unsigned char *q = (const unsigned char *)memchr(p, '\n', (size_t)(e - p));
		if (!q) break;
		out[found++] = q;
		p = q + 1;
	}
	return found;
}

typedef int (*Newlines4)(const unsigned char *, const unsigned char *, const unsigned char **);
Newlines4 pick_newlines4()
{
	const char *force = getenv("NTSM_SCAN_ISA");           // "memchr" | "avx2" | "avx512": tests walk all three
	const bool a2 = __builtin_cpu_supports("avx2"), a5 = a2 && __builtin_cpu_supports("avx512bw");
	if (force) {
		if (!strcmp(force, "memchr")) return newlines4_memchr;
		if (!strcmp(force, "avx2") && a2) return newlines4_avx2;
		if (!strcmp(force, "avx512") && a5) return newlines4_avx512;
	}
	return a5 ? newlines4_avx512 : a2 ? newlines4_avx2 : newlines4_memchr;
}
Newlines4 g_newlines4 = pick_newlines4();

}  // namespace

void fastx_rescan_isa() { g_newlines4 = pick_newlines4(); }
const char *fastx_scan_isa() { return g_newlines4 == newlines4_avx512 ? "avx512" : g_newlines4 == newlines4_avx2 ? "avx2" : "memchr"; }

bool FastxReader::next_fast(int64_t *len)
{
	if (last_ != 0 || err_) return false;                  // only when the next record starts at the next byte
	for (int attempt = 0; attempt < 2; ++attempt) {
		if (beg_ >= end_) {
			if (!fill()) return false;
		}
		const unsigned char *p = buf_ + beg_, *e = buf_ + end_;
		if (*p != '@') return false;
		const unsigned char *nl[4];
		const unsigned char *nl1, *nl2, *nl3, *nl4;
		bool complete = false;
		do {
			// the same walk the four memchr calls did: each newline is looked for from just behind the
			// previous one (the one skipped byte, nl2[1], is checked to be '+')
			const int got = g_newlines4(p + 1, e, nl);
			if (got < 1) break;
			nl1 = nl[0];
			const unsigned char *s0 = nl1 + 1;
			if (s0 >= e) break;
			if (*s0 == '\n' || *s0 == '>' || *s0 == '+' || *s0 == '@') return false;
			if (got < 2) break;
			nl2 = nl[1];
			if (nl2 + 1 >= e) break;
			if (nl2[1] != '+') return false;                 // multi-line sequence, FASTA, ...
			if (got < 3) break;
			nl3 = nl[2];
			const unsigned char *q0 = nl3 + 1;
			if (got < 4) break;
			nl4 = nl[3];
			complete = true;
			size_t sl = (size_t)(nl2 - s0), ql = (size_t)(nl4 - q0);
			if (sl > 1 && s0[sl - 1] == '\r') --sl;          // kseq.h:141, applied per appended line
			if (ql > 1 && q0[ql - 1] == '\r') --ql;
			if (ql != sl) return false;                      // short (continues on the next line) or long (-2): general parser
			cur_seq_ = (const char *)s0;
			cur_len_ = sl;
			fast_name_ = p + 1;
			fast_name_len_ = (size_t)(nl1 - p - 1);
			beg_ = (size_t)(nl4 - buf_) + 1;
			*len = (int64_t)sl;
			return true;
		} while (0);
		if (complete || eof_ || attempt) return false;
		// the record runs past the window: slide the unread tail to the front and top the window up
		const size_t tail = end_ - beg_;
		if (tail >= cap_ / 2) return false;                // a record longer than half the window is not the common case
		memmove(buf_, buf_ + beg_, tail);
		beg_ = 0;
		end_ = tail;
		const int n = src_.read(buf_ + tail, (unsigned)(cap_ - tail));
		if (n < 0) { src_err_ = true; eof_ = true; return false; }
		if ((size_t)n < cap_ - tail) {              // the source only comes back short at the end of the input,
			eof_ = true;                                   // or ahead of a data error it will report next
			if (src_.bad()) src_err_ = true;
		}
		end_ = tail + (size_t)n;
		if (n == 0) return false;
	}
	return false;
}

int64_t FastxReader::next()
{
	int64_t l;
	if (mapped_ && beg_ + (4u << 20) > populated_) populate_ahead();
	if (next_fast(&l)) return l;
	fast_name_ = nullptr;
	l = next_general();
	cur_seq_ = seq_.data();
	cur_len_ = seq_.size();
	return l;
}

int64_t FastxReader::next_general()
{
	int c;
	if (last_ == 0) {                                     // hunt for the next header byte
		for (;;) {
			if (!fill()) return err_ ? -3 : -1;
			while (beg_ < end_ && buf_[beg_] != '>' && buf_[beg_] != '@') ++beg_;
			if (beg_ < end_) break;
		}
		last_ = buf_[beg_++];
	}
	seq_.clear();
	qual_.clear();
	// name: up to the first white-space byte
	name_.clear();
	bool got = false;
	int delim = 0;
	for (;;) {
		if (err_) return -3;
		if (!fill()) { if (err_) return -3; break; }
		got = true;
		size_t i = beg_;
		while (i < end_) {
			const unsigned char ch = buf_[i];
			if (ch == ' ' || (ch >= '\t' && ch <= '\r')) break;
			++i;
		}
		name_.append((const char *)buf_ + beg_, i - beg_);
		const bool hit = i < end_;
		if (hit) delim = buf_[i];
		beg_ = i + (hit ? 1 : 0);
		if (hit) break;
	}
	if (!got) return -1;                                   // header byte was the last byte of the input
	if (delim != '\n') skip_line();                        // comment
	// sequence lines
	while ((c = getc()) >= 0 && c != '>' && c != '+' && c != '@') {
		if (c == '\n') continue;
		seq_.push_back((char)c);
		take_line(seq_);
	}
	if (c == '>' || c == '@') last_ = c;
	if (c != '+') return (int64_t)seq_.size();             // FASTA record (or input ended)
	// quality block
	while ((c = getc()) >= 0 && c != '\n') {}
	if (c == -1) return -2;
	while (take_line(qual_) && qual_.size() < seq_.size()) {}
	last_ = 0;
	if (qual_.size() != seq_.size()) return -2;
	return (int64_t)seq_.size();
}

}  // namespace ntsm

// ---------------------------------------------------------------- C ABI
struct ntsm_reader {
	ntsm::FastxReader r;
};

extern "C" int ntsm_reader_open(ntsm_reader **out, const char *path)
{
	if (!out || !path) return NTSM_ERR_ARG;
	ntsm_reader *h = new ntsm_reader();
	if (!h->r.open(path)) {
		delete h;
		return NTSM_ERR_IO;
	}
	*out = h;
	return NTSM_OK;
}

extern "C" const char *ntsm_scan_isa(const char *force)
{
	if (force) {
		if (*force) setenv("NTSM_SCAN_ISA", force, 1);
		else unsetenv("NTSM_SCAN_ISA");
		ntsm::fastx_rescan_isa();
	}
	return ntsm::fastx_scan_isa();
}

extern "C" int ntsm_reader_open2(ntsm_reader **out, const char *path, int helpers)
{
	if (!out || !path) return NTSM_ERR_ARG;
	ntsm_reader *h = new ntsm_reader();
	if (!h->r.open(path, helpers)) {
		delete h;
		return NTSM_ERR_IO;
	}
	*out = h;
	return NTSM_OK;
}

extern "C" int64_t ntsm_reader_next(ntsm_reader *r, const char **seq)
{
	const int64_t l = r->r.next();
	if (seq) *seq = r->r.seq();
	return l;
}

extern "C" const char *ntsm_reader_name(const ntsm_reader *r) { return const_cast<ntsm_reader *>(r)->r.name(); }

extern "C" void ntsm_reader_close(ntsm_reader *r) { delete r; }
