// fastx.cpp -- see fastx.h.  Grammar per vendor/kseq.h:178-219; return codes per :171-176.
#include "fastx.h"

#include <string.h>

#include "../../include/ntsm_b200.h"

namespace ntsm {

static constexpr size_t kWindow = 1u << 20;

bool FastxReader::open(const char *path)
{
	close();
	f_ = gzopen(path, "r");          // plain files pass through, like the reference (FingerPrint.hpp:50)
	if (!f_) return false;
	gzbuffer(f_, 1u << 18);
	buf_.resize(kWindow);
	beg_ = end_ = 0;
	eof_ = err_ = false;
	last_ = 0;
	return true;
}

void FastxReader::close()
{
	if (f_) gzclose(f_);
	f_ = nullptr;
}

bool FastxReader::fill()
{
	if (beg_ < end_) return true;
	if (eof_) return false;
	beg_ = 0;
	const int n = gzread(f_, buf_.data(), (unsigned)buf_.size());
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
		const unsigned char *p = buf_.data() + beg_;
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
		const unsigned char *p = buf_.data() + beg_;
		const unsigned char *nl = (const unsigned char *)memchr(p, '\n', end_ - beg_);
		if (nl) { beg_ = (size_t)(nl - buf_.data()) + 1; return true; }
		beg_ = end_;
	}
}

int64_t FastxReader::next()
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
		name_.append((const char *)buf_.data() + beg_, i - beg_);
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

extern "C" int64_t ntsm_reader_next(ntsm_reader *r, const char **seq)
{
	const int64_t l = r->r.next();
	if (seq) *seq = r->r.seq();
	return l;
}

extern "C" const char *ntsm_reader_name(const ntsm_reader *r) { return r->r.name(); }

extern "C" void ntsm_reader_close(ntsm_reader *r) { delete r; }
