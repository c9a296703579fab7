// gzsource.cpp -- see gzsource.h.
#include "gzsource.h"

#include <fcntl.h>
#include <immintrin.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

#include "inflate.h"
#include "pargz.h"

namespace ntsm {

// ---------------------------------------------------------------------------------------------
// CRC-32 by folding with carry-less multiplication (Gopal et al., "Fast CRC computation for
// generic polynomials using PCLMULQDQ"): four 128-bit lanes folded 64 bytes at a time, reduced to
// one lane, then to 32 bits by Barrett reduction.  The first call checks the routine against
// zlib's crc32 on a fixed pattern and falls back to zlib for good if they ever disagree.
namespace {

__attribute__((target("pclmul,sse4.1"))) uint32_t crc32_clmul(uint32_t crc, const uint8_t *buf, size_t len)
{
	// len >= 64 and a multiple of 16; crc is the running (already inverted) register
	const __m128i k1k2 = _mm_set_epi64x(0x01c6e41596, 0x0154442bd4);
	const __m128i k3k4 = _mm_set_epi64x(0x00ccaa009e, 0x01751997d0);
	const __m128i k5 = _mm_set_epi64x(0, 0x0163cd6124);
	const __m128i poly = _mm_set_epi64x(0x01f7011641, 0x01db710641);
	__m128i x1 = _mm_loadu_si128((const __m128i *)(buf + 0)), x2 = _mm_loadu_si128((const __m128i *)(buf + 16));
	__m128i x3 = _mm_loadu_si128((const __m128i *)(buf + 32)), x4 = _mm_loadu_si128((const __m128i *)(buf + 48));
	x1 = _mm_xor_si128(x1, _mm_cvtsi32_si128((int)crc));
	buf += 64;
	len -= 64;
	while (len >= 64) {
		const __m128i a1 = _mm_clmulepi64_si128(x1, k1k2, 0x00), a2 = _mm_clmulepi64_si128(x2, k1k2, 0x00);
		const __m128i a3 = _mm_clmulepi64_si128(x3, k1k2, 0x00), a4 = _mm_clmulepi64_si128(x4, k1k2, 0x00);
		x1 = _mm_clmulepi64_si128(x1, k1k2, 0x11);
		x2 = _mm_clmulepi64_si128(x2, k1k2, 0x11);
		x3 = _mm_clmulepi64_si128(x3, k1k2, 0x11);
		x4 = _mm_clmulepi64_si128(x4, k1k2, 0x11);
		x1 = _mm_xor_si128(_mm_xor_si128(x1, a1), _mm_loadu_si128((const __m128i *)(buf + 0)));
		x2 = _mm_xor_si128(_mm_xor_si128(x2, a2), _mm_loadu_si128((const __m128i *)(buf + 16)));
		x3 = _mm_xor_si128(_mm_xor_si128(x3, a3), _mm_loadu_si128((const __m128i *)(buf + 32)));
		x4 = _mm_xor_si128(_mm_xor_si128(x4, a4), _mm_loadu_si128((const __m128i *)(buf + 48)));
		buf += 64;
		len -= 64;
	}
	// four lanes -> one
	__m128i a = _mm_clmulepi64_si128(x1, k3k4, 0x00);
	x1 = _mm_xor_si128(_mm_xor_si128(_mm_clmulepi64_si128(x1, k3k4, 0x11), a), x2);
	a = _mm_clmulepi64_si128(x1, k3k4, 0x00);
	x1 = _mm_xor_si128(_mm_xor_si128(_mm_clmulepi64_si128(x1, k3k4, 0x11), a), x3);
	a = _mm_clmulepi64_si128(x1, k3k4, 0x00);
	x1 = _mm_xor_si128(_mm_xor_si128(_mm_clmulepi64_si128(x1, k3k4, 0x11), a), x4);
	while (len >= 16) {
		a = _mm_clmulepi64_si128(x1, k3k4, 0x00);
		x1 = _mm_xor_si128(_mm_xor_si128(_mm_clmulepi64_si128(x1, k3k4, 0x11), a), _mm_loadu_si128((const __m128i *)buf));
		buf += 16;
		len -= 16;
	}
	// 128 -> 64 -> 32 bits
	const __m128i mask32 = _mm_setr_epi32(~0, 0, ~0, 0);
	__m128i t = _mm_clmulepi64_si128(x1, k3k4, 0x10);
	x1 = _mm_xor_si128(_mm_srli_si128(x1, 8), t);
	t = _mm_srli_si128(x1, 4);
	x1 = _mm_and_si128(x1, mask32);
	x1 = _mm_xor_si128(_mm_clmulepi64_si128(x1, k5, 0x00), t);
	t = _mm_and_si128(x1, mask32);
	t = _mm_clmulepi64_si128(t, poly, 0x10);
	t = _mm_and_si128(t, mask32);
	t = _mm_clmulepi64_si128(t, poly, 0x00);
	x1 = _mm_xor_si128(x1, t);
	return (uint32_t)_mm_extract_epi32(x1, 1);
}

bool clmul_usable()
{
	static const bool ok = [] {
		if (!__builtin_cpu_supports("pclmul") || !__builtin_cpu_supports("sse4.1")) return false;
		if (const char *e = getenv("NTSM_CRC")) {
			if (!strcmp(e, "zlib")) return false;
		}
		uint8_t pat[64 * 5 + 16];
		for (size_t i = 0; i < sizeof pat; ++i) pat[i] = (uint8_t)(i * 131u + (i >> 3) * 7u + 5u);
		const uint32_t want = (uint32_t)crc32(0x12345678u, pat, sizeof pat);
		const uint32_t got = ~crc32_clmul(~0x12345678u, pat, sizeof pat);
		return want == got;
	}();
	return ok;
}

}  // namespace

uint32_t crc32_fast(uint32_t crc, const uint8_t *buf, size_t len)
{
	if (len >= 256 && clmul_usable()) {
		const size_t body = len & ~(size_t)15;
		crc = ~crc32_clmul(~crc, buf, body);
		buf += body;
		len -= body;
	}
	while (len) {                                             // zlib takes a 32-bit length on some builds
		const size_t n = len > (1u << 30) ? (1u << 30) : len;
		crc = (uint32_t)crc32(crc, buf, (uInt)n);
		buf += n;
		len -= n;
	}
	return crc;
}

// ---------------------------------------------------------------------------------------------
namespace {

constexpr size_t kHist = 32768;                 // DEFLATE window
constexpr size_t kChunk = 1u << 20;             // decoded per Inflater::run call in fast mode
constexpr uint32_t kBgzfMaxOut = 65536;         // a BGZF block never inflates to more than 64 KiB
constexpr int kBatchBlocks = 32;                // BGZF blocks one helper takes at a time (<= 2 MiB of output)

inline uint32_t le32(const uint8_t *p) { return p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

// gzip member header at p (RFC 1952).  Returns 0 = regular (hdr_len, and bsize != 0 when it is a BGZF
// block), 1 = not a gzip member at all, -1 = a member this code leaves to zlib (header CRC, reserved
// flags, unknown method, truncated header).
int parse_member_header(const uint8_t *p, size_t avail, size_t *hdr_len, uint32_t *bsize)
{
	*bsize = 0;
	if (avail < 2 || p[0] != 0x1f || p[1] != 0x8b) return 1;
	if (avail < 10 || p[2] != 8) return -1;
	const uint8_t flg = p[3];
	if (flg & 0xE2) return -1;                               // reserved bits, FHCRC
	size_t q = 10;
	if (flg & 4) {                                           // FEXTRA
		if (avail < q + 2) return -1;
		const size_t xlen = p[q] | ((size_t)p[q + 1] << 8);
		q += 2;
		if (avail < q + xlen) return -1;
		for (size_t s = q; s + 4 <= q + xlen;) {
			const size_t slen = p[s + 2] | ((size_t)p[s + 3] << 8);
			if (p[s] == 'B' && p[s + 1] == 'C' && slen == 2 && s + 6 <= q + xlen) *bsize = (uint32_t)(p[s + 4] | (p[s + 5] << 8)) + 1;
			s += 4 + slen;
		}
		q += xlen;
	}
	for (int f = 8; f <= 16; f <<= 1) {                      // FNAME, FCOMMENT: zero-terminated
		if (!(flg & f)) continue;
		const void *z = q < avail ? memchr(p + q, 0, avail - q) : nullptr;
		if (!z) return -1;
		q = (size_t)((const uint8_t *)z - p) + 1;
	}
	*hdr_len = q;
	return 0;
}

// one batch of consecutive BGZF blocks, inflated by one helper
struct Batch {
	size_t first_off = 0;                 // file offset of the first block
	uint32_t sizes[kBatchBlocks];         // compressed size of each block (BSIZE + 1)
	int n_blocks = 0;
	std::vector<uint8_t> out;             // inflated bytes of the blocks that decoded cleanly, in order
	size_t out_len = 0;
	int failed_at = -1;                   // first block that was not regular (its bytes are not in `out`)
	size_t failed_off = 0;
	bool done = false;
};

}  // namespace

struct GzSource::Impl {
	std::string path;
	enum Mode { kGzread, kFast, kBgzf, kZinflate, kPlainMap } mode = kGzread;
	size_t plain_pos = 0;                 // kPlainMap: bytes handed out by read() so far
	bool fell_back = false;
	bool eof = false;                     // the producers have nothing more (clean end, trailing garbage, or truncated input)
	bool failed = false;                  // a data error was met: what precedes it in whole 16 KiB reads is delivered, then -1
	gzFile gz = nullptr;
	bool gz_direct = false;
	const uint8_t *map = nullptr;
	size_t size = 0;
	bool map_is_heap = false;             // NTSM_GZ_INPUT=read: the file was read into memory instead of mapped (measurement knob)

	// ---- staging -------------------------------------------------------------------------------
	// How much of a DAMAGED gzip file the reference gets to see follows from zlib's gzread.c and from
	// kseq asking for 16 KiB at a time (vendor/kseq.h:229): gzread inflates a member in units of 16 KiB
	// of output counted from the member's start (gz_fetch's internal buffer, or straight into the
	// caller's buffer once the two line up), and when inflate reports a data error inside a unit,
	// gz_decomp returns before that unit's output is accounted and the gzread call in progress comes
	// back -1 with everything it had gathered.  So the reference sees the whole 16 KiB reads that end
	// at or before the start of the failing unit.  Producers therefore hand over segments of decoded
	// bytes together with the stream offset at which the unit now being decoded began (`risk_base`:
	// were the very next symbol bad, that unit would fail); bytes are released to the caller only
	// below the 16 KiB boundary at or under it -- or under `fail_base` once a real error is known, or
	// all of them once the input has ended cleanly (a truncated file just ends, gzread.c: "unexpected
	// end of file" is not an error for the caller).
	static constexpr uint64_t kRefRead = 16384;
	uint64_t risk_base = 0, fail_base = 0;
	static uint64_t unit_start(uint64_t member_gstart, uint64_t member_out)
	{
		return member_gstart + (member_out ? (member_out - 1) / kRefRead * kRefRead : 0);
	}
	const uint8_t *seg = nullptr;         // current segment (owned by the producer until the next produce call)
	size_t seg_len = 0, seg_pos = 0;
	std::vector<uint8_t> held;            // decoded bytes that precede the segment and were not releasable yet
	size_t held_out = 0;
	uint64_t produced = 0, delivered = 0; // global stream offsets

	// ---- fast mode: one member at a time through a sliding window --------------------------------
	size_t member_off = 0;
	uint64_t member_produced = 0;         // bytes of the current member handed to staging
	uint64_t member_gstart = 0;           // stream offset of its first byte
	bool in_member = false;
	std::unique_ptr<Inflater> inf;
	std::vector<uint8_t> win;
	uint8_t *out = nullptr;
	const uint8_t *hist = nullptr;
	uint32_t crc = 0, isize = 0;
	// a large single member is cut between worker threads (pargz.h); what they confirm arrives here in order
	std::unique_ptr<ParallelInflate> par;
	int par_workers = 0;
	size_t par_min = 16u << 20, par_chunk = 1u << 20;
	uint64_t par_chunks = 0;              // chunks accepted from workers so far (introspection)

	// ---- zlib inflate() over the mapping: takes over at a member the fast modes found irregular --
	z_stream zs;
	bool zs_init = false, z_in_member = false;
	size_t z_off = 0;                     // next input byte
	uint64_t z_skip = 0;                  // bytes of the member that were already handed to staging
	uint64_t z_gstart = 0, z_member_out = 0;   // stream offset of the member's first byte; bytes of it inflated so far
	std::vector<uint8_t> zbuf;

	// ---- bgzf mode: the caller's thread scans block headers and queues batches; helpers inflate ---
	std::vector<std::thread> helpers;
	std::vector<std::unique_ptr<Batch>> slots;
	std::deque<Batch *> todo;             // queued, not yet taken by a helper
	std::deque<Batch *> order;            // every batch in flight, in file order
	std::vector<Batch *> free_slots;
	std::mutex mu;
	std::condition_variable cv_work, cv_done;
	bool stop = false;
	size_t scan_off = 0;                  // next block header to look at
	bool scan_ended = false;              // scan_off is not a BGZF block (or is past the end): fast mode continues there
	Batch *cur = nullptr;                 // batch whose bytes are the current segment

	~Impl() { shutdown(); }

	void shutdown()
	{
		par.reset();
		stop_helpers();
		if (gz) gzclose(gz);
		gz = nullptr;
		if (zs_init) inflateEnd(&zs);
		zs_init = false;
		if (map && map_is_heap) free((void *)map);
		else if (map) munmap((void *)map, size);
		map = nullptr;
	}

	void stop_helpers()
	{
		if (helpers.empty()) return;
		{
			std::lock_guard<std::mutex> g(mu);
			stop = true;
		}
		cv_work.notify_all();
		for (auto &t : helpers) t.join();
		helpers.clear();
		std::lock_guard<std::mutex> g(mu);
		stop = false;
		todo.clear();
		order.clear();
	}

	void set_segment(const uint8_t *p, size_t n)
	{
		seg = p;
		seg_len = n;
		seg_pos = 0;
		produced += n;
		risk_base = produced;                 // producers in the middle of a member lower this right after
	}

	// hand the rest of the file to zlib's inflate from the gzip member at `off`, of which `skip` bytes
	// are already in staging
	void fallback(size_t off, uint64_t skip)
	{
		fell_back = true;
		mode = kZinflate;
		z_off = off;
		z_skip = skip;
		z_gstart = produced - skip;
		z_in_member = false;
		if (!zs_init) {
			memset(&zs, 0, sizeof zs);
			if (inflateInit2(&zs, 15 + 16) != Z_OK) {
				failed = true;
				return;
			}
			zs_init = true;
			zbuf.resize(kChunk);
		}
	}

	// zlib mode: one segment per call, mirroring gzread.c's gz_look / gz_decomp decisions
	bool produce_zinflate()
	{
		for (;;) {
			if (failed || eof) return false;
			if (!z_in_member) {
				if (size - z_off < 2 || map[z_off] != 0x1f || map[z_off + 1] != 0x8b) {   // gz_look: trailing garbage is ignored
					eof = true;
					return false;
				}
				inflateReset(&zs);
				z_in_member = true;
				z_member_out = 0;
				if (!z_skip) z_gstart = produced;
			}
			zs.next_out = zbuf.data();
			zs.avail_out = (uInt)zbuf.size();
			bool member_end = false;
			while (zs.avail_out) {
				const size_t left = size - z_off;
				if (left == 0) {                              // gz_decomp: "unexpected end of file" is not an error for gzread
					eof = true;
					break;
				}
				zs.next_in = (Bytef *)(map + z_off);
				zs.avail_in = (uInt)std::min<size_t>(left, 1u << 30);
				const uInt before = zs.avail_in;
				const int ret = inflate(&zs, Z_NO_FLUSH);
				z_off += before - zs.avail_in;
				if (ret == Z_STREAM_END) {
					member_end = true;
					break;
				}
				if (ret == Z_BUF_ERROR && zs.avail_out != 0 && zs.avail_in == 0) continue;   // wants more input: the loop ends when there is none
				if (ret != Z_OK && ret != Z_BUF_ERROR) {      // Z_DATA_ERROR and friends: everything decoded before it counts
					failed = true;
					break;
				}
			}
			if (member_end) z_in_member = false;
			size_t n = zbuf.size() - zs.avail_out;
			const uint8_t *p = zbuf.data();
			z_member_out += n;
			if (failed) fail_base = unit_start(z_gstart, z_member_out);
			if (z_skip) {
				const size_t d = (size_t)std::min<uint64_t>(z_skip, n);
				z_skip -= d;
				p += d;
				n -= d;
			}
			if (member_end) z_skip = 0;
			if (n) {
				set_segment(p, n);
				if (!member_end) risk_base = std::min(produced, unit_start(z_gstart, z_member_out));
				return true;
			}
		}
	}

	// fast mode: one segment per call.  false = none (end of input, or the mode changed).
	bool produce_fast()
	{
		uint8_t *const limit = win.data() + kHist + kChunk;
		for (;;) {
			if (!in_member) {
				size_t hdr = 0;
				uint32_t bsize = 0;
				const int r = parse_member_header(map + member_off, size - member_off, &hdr, &bsize);
				if (r == 1) {                                  // zlib: anything but a gzip magic after a member is ignored
					eof = true;
					return false;
				}
				if (r < 0) {
					fallback(member_off, 0);
					return false;
				}
				inf->begin(map + member_off + hdr, map + size);
				out = win.data();
				hist = win.data();
				crc = 0;
				isize = 0;
				member_produced = 0;
				member_gstart = produced;
				in_member = true;
				if (par_workers > 0 && size - (member_off + hdr) >= par_min)
					par.reset(new ParallelInflate(map, size, member_off + hdr, par_workers, par_chunk));
			}
			if (par) {
				const uint8_t *pp = nullptr;
				size_t pn = 0;
				uint32_t pcrc = 0;
				if (par->next(&pp, &pn, &pcrc)) {
					crc = (uint32_t)crc32_combine(crc, pcrc, (z_off_t)pn);   // the workers hashed their chunks
					isize += (uint32_t)pn;
					member_produced += pn;
					set_segment(pp, pn);
					risk_base = unit_start(member_gstart, member_produced);
					return true;
				}
				par_chunks += par->chunks_accepted();
				if (par->end() == ParallelInflate::kStreamEnd) {
					const uint8_t *t = map + par->end_byte();
					par.reset();
					if ((size_t)(map + size - t) < 8 || le32(t) != crc || le32(t + 4) != isize) {
						fallback(member_off, member_produced);
						return false;
					}
					member_off = (size_t)(t + 8 - map);
					in_member = false;
					risk_base = produced;                      // the member is complete and checked
					continue;
				}
				// the workers stopped at a confirmed block boundary: one thread carries on from that bit
				const std::vector<uint8_t> &w = par->window();
				if (!w.empty()) memcpy(win.data(), w.data(), w.size());
				hist = win.data();
				out = win.data() + w.size();
				inf->begin_bits(map, par->resume_bit(), map + size);
				par.reset();
			}
			if (out >= limit) {                               // the window is full (and handed over): keep the history, rewind
				const size_t keep = std::min<size_t>(kHist, (size_t)(out - hist));
				memmove(win.data(), out - keep, keep);
				hist = win.data();
				out = win.data() + keep;
			}
			uint8_t *const from = out;
			const Inflater::Status st = inf->run(hist, &out, limit);
			bool bad = st == Inflater::kError;
			if (!bad) {
				crc = crc32_fast(crc, from, (size_t)(out - from));
				isize += (uint32_t)(out - from);
				if (st == Inflater::kStreamEnd) {
					const uint8_t *t = inf->in_pos();
					if ((size_t)(map + size - t) < 8 || le32(t) != crc || le32(t + 4) != isize) bad = true;
					else {
						member_off = (size_t)(t + 8 - map);
						in_member = false;
					}
				}
			}
			if (bad) {                                        // zlib decides what the caller sees of this member
				fallback(member_off, member_produced);
				return false;
			}
			if (out > from) {
				member_produced += (uint64_t)(out - from);
				set_segment(from, (size_t)(out - from));
				if (in_member) risk_base = unit_start(member_gstart, member_produced);
				return true;
			}
		}
	}

	// ---- bgzf -------------------------------------------------------------------------------
	void helper_main()
	{
		std::unique_ptr<Inflater> my(new Inflater());
		for (;;) {
			Batch *b;
			{
				std::unique_lock<std::mutex> g(mu);
				cv_work.wait(g, [&] { return stop || !todo.empty(); });
				if (stop) return;
				b = todo.front();
				todo.pop_front();
			}
			size_t off = b->first_off;
			uint8_t *o = b->out.data();
			for (int i = 0; i < b->n_blocks; ++i) {
				size_t hdr = 0;
				uint32_t bsize = 0;
				const uint8_t *blk = map + off;
				bool ok = parse_member_header(blk, b->sizes[i], &hdr, &bsize) == 0 && bsize == b->sizes[i] && hdr + 8 <= bsize;
				if (ok) {
					uint8_t *w = o;
					my->begin(blk + hdr, blk + bsize - 8);
					const Inflater::Status st = my->run(o, &w, o + kBgzfMaxOut + 1);   // +1: a block of exactly 64 KiB must still reach its end-of-block
					ok = st == Inflater::kStreamEnd && (size_t)(w - o) <= kBgzfMaxOut && my->in_pos() == blk + bsize - 8;
					if (ok) {
						const uint32_t n = (uint32_t)(w - o);
						ok = le32(blk + bsize - 8) == crc32_fast(0, o, n) && le32(blk + bsize - 4) == n;
						if (ok) o = w;
					}
				}
				if (!ok) {
					b->failed_at = i;
					b->failed_off = off;
					break;
				}
				off += b->sizes[i];
			}
			b->out_len = (size_t)(o - b->out.data());
			{
				std::lock_guard<std::mutex> g(mu);
				b->done = true;
			}
			cv_done.notify_all();
		}
	}

	// look at block headers from scan_off on and queue batches while slots are free
	void scan_ahead()
	{
		while (!scan_ended) {
			Batch *b;
			{
				std::lock_guard<std::mutex> g(mu);
				if (free_slots.empty()) return;
				b = free_slots.back();
				free_slots.pop_back();
			}
			b->first_off = scan_off;
			b->n_blocks = 0;
			b->out_len = 0;
			b->failed_at = -1;
			b->done = false;
			while (b->n_blocks < kBatchBlocks) {
				size_t hdr = 0;
				uint32_t bsize = 0;
				if (parse_member_header(map + scan_off, size - scan_off, &hdr, &bsize) != 0 || bsize == 0 || bsize > size - scan_off || hdr + 8 > bsize) {
					scan_ended = true;
					break;
				}
				b->sizes[b->n_blocks++] = bsize;
				scan_off += bsize;
			}
			std::lock_guard<std::mutex> g(mu);
			if (b->n_blocks == 0) {
				free_slots.push_back(b);
				return;
			}
			todo.push_back(b);
			order.push_back(b);
			cv_work.notify_one();
		}
	}

	// bgzf mode: the next batch becomes the segment.  false = none (the mode changed).
	bool produce_bgzf()
	{
		for (;;) {
			if (cur) {                                        // its bytes are delivered or copied to `held` by now
				const bool bad = cur->failed_at >= 0;
				const size_t foff = cur->failed_off;
				{
					std::lock_guard<std::mutex> g(mu);
					free_slots.push_back(cur);
				}
				cur = nullptr;
				if (bad) {                                    // a block that is not what its header promised: zlib from there
					stop_helpers();
					fallback(foff, 0);
					return false;
				}
			}
			scan_ahead();
			std::unique_lock<std::mutex> g(mu);
			if (order.empty()) {                              // every BGZF block is decoded; whatever follows goes the ordinary way
				g.unlock();
				stop_helpers();
				mode = kFast;
				member_off = scan_off;
				in_member = false;
				return false;
			}
			Batch *b = order.front();
			cv_done.wait(g, [&] { return b->done; });
			order.pop_front();
			g.unlock();
			cur = b;
			if (b->out_len) {
				set_segment(b->out.data(), b->out_len);
				return true;
			}
		}
	}

	// gzread itself, for what cannot be mapped.  When the content is gzip this is the reference's own call
	// pattern -- aligned 16 KiB reads with zlib's default buffer, each inflated straight into the
	// destination -- so a data error costs exactly the bytes it costs the reference.
	std::vector<uint8_t> gzchunk;
	size_t gzchunk_len = 0, gzchunk_pos = 0;
	int read_gzread(uint8_t *d, unsigned n)
	{
		unsigned total = 0;
		while (total < n) {
			if (gzchunk_pos < gzchunk_len) {
				const size_t k = std::min<size_t>(gzchunk_len - gzchunk_pos, n - total);
				memcpy(d + total, gzchunk.data() + gzchunk_pos, k);
				gzchunk_pos += k;
				total += (unsigned)k;
				continue;
			}
			if (failed) return total ? (int)total : -1;
			if (eof) break;
			if (gz_direct) {                                  // plain file: no inflate, no error granularity to keep
				const int r = gzread(gz, d + total, n - total);
				if (r < 0) failed = true;
				else {
					if ((unsigned)r < n - total) eof = true;
					total += (unsigned)r;
				}
				continue;
			}
			gzchunk.resize(kRefRead);
			const int r = gzread(gz, gzchunk.data(), (unsigned)kRefRead);
			gzchunk_pos = 0;
			gzchunk_len = r > 0 ? (size_t)r : 0;
			if (r < 0) failed = true;
			else if ((uint64_t)r < kRefRead) eof = true;
		}
		return (int)total;
	}
};

GzSource::GzSource() {}
GzSource::~GzSource() { close(); }

void GzSource::close() { p_.reset(); }

const char *GzSource::mode() const
{
	if (!p_) return "closed";
	return p_->mode == Impl::kFast ? "fast" : p_->mode == Impl::kBgzf ? "bgzf" : p_->mode == Impl::kPlainMap ? "mapped" : "zlib";
}

bool GzSource::mapped(const uint8_t **base, size_t *size) const
{
	if (!p_ || p_->mode != Impl::kPlainMap) return false;
	*base = p_->map;
	*size = p_->size;
	return true;
}

bool GzSource::fell_back() const { return p_ && p_->fell_back; }
bool GzSource::bad() const { return p_ && p_->failed; }
uint64_t GzSource::parallel_chunks() const { return p_ ? p_->par_chunks + (p_->par ? p_->par->chunks_accepted() : 0) : 0; }

bool GzSource::open(const char *path, int helpers, bool map_plain)
{
	close();
	p_.reset(new Impl());
	Impl &s = *p_;
	s.path = path;
	const char *force = getenv("NTSM_INFLATE");
	const bool want_fast = !(force && !strcmp(force, "zlib"));
	const int fd = ::open(path, O_RDONLY);
	if (fd < 0) {
		p_.reset();
		return false;
	}
	struct stat sb;
	unsigned char magic[2] = { 0, 0 };
	if (want_fast && fstat(fd, &sb) == 0 && S_ISREG(sb.st_mode) && sb.st_size >= 18 && pread(fd, magic, 2, 0) == 2 && magic[0] == 0x1f &&
	    magic[1] == 0x8b) {
		void *m = MAP_FAILED;
		const char *gin = getenv("NTSM_GZ_INPUT");
		if (gin && !strcmp(gin, "read")) {
			m = malloc((size_t)sb.st_size);
			size_t got = 0;
			while (m && got < (size_t)sb.st_size) {
				const ssize_t r = pread(fd, (char *)m + got, (size_t)sb.st_size - got, (off_t)got);
				if (r <= 0) break;
				got += (size_t)r;
			}
			if (!m || got != (size_t)sb.st_size) {
				free(m);
				m = MAP_FAILED;
			} else s.map_is_heap = true;
		} else {
			m = mmap(nullptr, (size_t)sb.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
			if (m != MAP_FAILED && !(gin && !strcmp(gin, "mmap_plain"))) madvise(m, (size_t)sb.st_size, MADV_SEQUENTIAL);
		}
		if (m != MAP_FAILED) {
			s.map = (const uint8_t *)m;
			s.size = (size_t)sb.st_size;
			::close(fd);
			s.mode = Impl::kFast;
			s.inf.reset(new Inflater());
			s.win.resize(kHist + kChunk + Inflater::kSlack);
			s.held.reserve(4 * Impl::kRefRead);
			size_t hdr = 0;
			uint32_t bsize = 0;
			if (const char *e = getenv("NTSM_PARGZ_MIN")) s.par_min = (size_t)strtoull(e, nullptr, 10);
			if (const char *e = getenv("NTSM_PARGZ_CHUNK")) s.par_chunk = (size_t)strtoull(e, nullptr, 10);
			const char *pgz = getenv("NTSM_PARALLEL_GZ");
			// decoding with markers and resolving them costs ~2.2x the CPU of the plain decoder (measured), so
			// cutting a member only pays with four workers or more
			int min_workers = 4;
			if (const char *e = getenv("NTSM_PARGZ_MIN_WORKERS")) min_workers = atoi(e);
			if (helpers >= min_workers && helpers > 0 && !(pgz && !strcmp(pgz, "0")) && !(force && !strcmp(force, "serial"))) s.par_workers = helpers;
			if (force && !strcmp(force, "zinflate")) s.fallback(0, 0);          // tests: zlib's inflate over the mapping from the start
			else if (helpers > 0 && !(force && !strcmp(force, "serial")) && parse_member_header(s.map, s.size, &hdr, &bsize) == 0 && bsize != 0) {
				s.mode = Impl::kBgzf;
				const int n_slots = 2 * helpers + 2;
				for (int i = 0; i < n_slots; ++i) {
					s.slots.emplace_back(new Batch());
					s.slots.back()->out.resize((size_t)kBatchBlocks * kBgzfMaxOut + Inflater::kSlack);
					s.free_slots.push_back(s.slots.back().get());
				}
				for (int i = 0; i < helpers; ++i) s.helpers.emplace_back([&s] { s.helper_main(); });
			}
			return true;
		}
	}
	// A regular file that does not start with the gzip magic: gzread would hand its bytes through
	// untouched (zlib's transparent mode, which the reference relies on for plain FASTA/FASTQ,
	// src/FingerPrint.hpp:50) -- at the price of read() copying every byte out of the page cache, which is
	// three quarters of the reader's time on cached files.  Map it instead: the reader scans the page
	// cache's own pages (FastxReader takes the mapping as its window), nothing is copied.  The price is
	// page-table work per 4 KiB page, which is cheaper than copying the page for one reader (4.6 vs 2.8
	// Gbases/s per thread on the GPU box) but does not scale inside one process: at 16 parser threads
	// the mapped readers fall to 1.7 Gbases/s each and gzread's 2.4 win (profiles/r02e_hostpath.txt,
	// r02f_hostpath.txt: every way of setting the mapping up lands on the same number).  So the caller
	// says whether to map (map_plain): the file pipeline does for up to 6 parser threads.
	{
		struct stat sb2;
		if (want_fast && fstat(fd, &sb2) == 0 && S_ISREG(sb2.st_mode) && sb2.st_size > 0 && !(magic[0] == 0x1f && magic[1] == 0x8b)) {
			unsigned char m2[2] = { 0, 0 };
			const bool is_gz = pread(fd, m2, 2, 0) == 2 && m2[0] == 0x1f && m2[1] == 0x8b;
			void *m = is_gz || !map_plain ? MAP_FAILED : mmap(nullptr, (size_t)sb2.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
			if (m != MAP_FAILED) {
				madvise(m, (size_t)sb2.st_size, MADV_SEQUENTIAL);
				s.map = (const uint8_t *)m;
				s.size = (size_t)sb2.st_size;
				::close(fd);
				s.mode = Impl::kPlainMap;
				return true;
			}
		}
	}
	// pipes, gzip files too small to be one, empty files, or no mapping: gzread does everything, as in the reference
	s.gz = gzdopen(fd, "r");
	if (!s.gz) {
		::close(fd);
		p_.reset();
		return false;
	}
	s.gz_direct = gzdirect(s.gz) != 0;
	s.mode = Impl::kGzread;
	return true;
}

int GzSource::read(void *dst, unsigned n)
{
	if (!p_) return -1;
	Impl &s = *p_;
	uint8_t *d = (uint8_t *)dst;
	if (s.mode == Impl::kGzread) return s.read_gzread(d, n);
	if (s.mode == Impl::kPlainMap) {                          // callers that want copies (ntsm_gz_read) still get gzread's contract
		const size_t k = std::min<size_t>(n, s.size - s.plain_pos);
		memcpy(d, s.map + s.plain_pos, k);
		s.plain_pos += k;
		return (int)k;
	}
	unsigned total = 0;
	while (total < n) {
		// what may be released: everything once the input ended cleanly, else whole reference-sized reads only
		const uint64_t base = s.failed ? s.fail_base : s.risk_base;
		const uint64_t safe = s.eof && !s.failed ? s.produced : base - base % Impl::kRefRead;
		if (s.held_out < s.held.size()) {                     // held bytes come first; they start at `delivered`
			const uint64_t ok = safe > s.delivered ? safe - s.delivered : 0;
			const size_t k = (size_t)std::min<uint64_t>(std::min<uint64_t>(s.held.size() - s.held_out, ok), n - total);
			if (k) {
				memcpy(d + total, s.held.data() + s.held_out, k);
				s.held_out += k;
				s.delivered += k;
				total += (unsigned)k;
				if (s.held_out == s.held.size()) {
					s.held.clear();
					s.held_out = 0;
				}
				continue;
			}
		} else if (s.seg_pos < s.seg_len) {
			const uint64_t ok = safe > s.delivered ? safe - s.delivered : 0;
			const size_t k = (size_t)std::min<uint64_t>(std::min<uint64_t>(s.seg_len - s.seg_pos, ok), n - total);
			if (k) {
				memcpy(d + total, s.seg + s.seg_pos, k);
				s.seg_pos += k;
				s.delivered += k;
				total += (unsigned)k;
				continue;
			}
		}
		// nothing releasable: finish, or decode more (the unreleased tail moves to `held` first, the producer reuses its buffer)
		if (s.failed) return total ? (int)total : -1;
		if (s.eof) break;
		if (s.seg_pos < s.seg_len) {
			s.held.insert(s.held.end(), s.seg + s.seg_pos, s.seg + s.seg_len);
			s.seg_pos = s.seg_len;
		}
		if (s.mode == Impl::kFast) s.produce_fast();
		else if (s.mode == Impl::kBgzf) s.produce_bgzf();
		else s.produce_zinflate();
	}
	return (int)total;
}

}  // namespace ntsm

// ---------------------------------------------------------------- C ABI
#include "../../include/ntsm_b200.h"

struct ntsm_gz {
	ntsm::GzSource s;
};

extern "C" int ntsm_gz_open(ntsm_gz **out, const char *path, int helpers)
{
	if (!out || !path) return NTSM_ERR_ARG;
	ntsm_gz *g = new ntsm_gz();
	if (!g->s.open(path, helpers)) {
		delete g;
		return NTSM_ERR_IO;
	}
	*out = g;
	return NTSM_OK;
}

extern "C" int ntsm_gz_read(ntsm_gz *g, void *dst, unsigned n) { return g ? g->s.read(dst, n) : -1; }
extern "C" const char *ntsm_gz_mode(const ntsm_gz *g) { return g ? g->s.mode() : ""; }
extern "C" int ntsm_gz_fell_back(const ntsm_gz *g) { return g && g->s.fell_back() ? 1 : 0; }
extern "C" uint64_t ntsm_gz_parallel_chunks(const ntsm_gz *g) { return g ? g->s.parallel_chunks() : 0; }
extern "C" void ntsm_gz_close(ntsm_gz *g) { delete g; }
extern "C" uint32_t ntsm_crc32(uint32_t crc, const void *buf, uint64_t len) { return ntsm::crc32_fast(crc, (const uint8_t *)buf, (size_t)len); }
