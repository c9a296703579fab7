// pargz.cpp -- see pargz.h.
#include "pargz.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>

#include "gzsource.h"
#include "inflate.h"

namespace ntsm {

namespace {

constexpr size_t kWin = 32768;
constexpr uint64_t kSearchChunks = 3;          // look for a block start over at most this many chunks

struct Task {
	uint64_t index = 0;
	uint64_t from_bit = 0, stop_bit = 0;       // search / start from here; stop in front of the first block at or after stop_bit
	bool exact_start = false;                  // chunk 0: the stream's own start, no unknown window
	// decode results
	bool ok = false, stream_end = false;
	uint64_t start_bit = 0, end_bit = 0;
	size_t end_byte = 0;
	std::unique_ptr<uint16_t[]> sym;           // [n_prefix markers] + decoded symbols: 0-255 bytes, 256 + w = byte w of the unknown 32 KiB before the chunk
	size_t sym_cap = 0, n_prefix = 0, n_sym = 0;
	bool decoded = false;
	// resolve: symbols -> bytes, in place at the front of `sym`, through the window that the stitcher supplies
	uint8_t window[kWin];
	size_t window_have = 0;                    // valid bytes at the END of window[] (< kWin only near the stream's start)
	bool resolved = false;
	uint32_t crc = 0;                          // CRC-32 of the chunk's bytes, computed by the worker that resolved it
	int job = 0;                               // what a worker is asked to do with it: 0 decode, 1 resolve
};

// byte for symbol v given the last `have` bytes before the chunk at the end of win[kWin]; -1 = refers to before the stream
inline int resolve_one(uint16_t v, const uint8_t *win, size_t have)
{
	if (v < 256) return v;
	const size_t w = (size_t)v - 256;
	return w < kWin - have ? -1 : win[w];
}

}  // namespace

struct ParallelInflate::Impl {
	const uint8_t *base;
	size_t size, deflate_start, chunk;
	uint64_t n_chunks;
	std::vector<std::thread> workers;
	std::mutex mu;
	std::condition_variable cv_work, cv_done;
	std::deque<Task *> todo;                   // jobs for the workers (resolve jobs go to the front)
	std::deque<Task *> order;                  // chunks being decoded, in stream order
	std::deque<Task *> resolving;              // accepted chunks whose bytes are being produced, in stream order
	std::vector<std::unique_ptr<Task>> pool;
	std::vector<Task *> free_tasks;
	bool stop = false;
	uint64_t next_issue = 0;                   // next chunk index to hand out
	// stitcher state (the consumer's thread)
	End end = kRunning;
	bool no_more_accepts = false;              // stream end seen, or the chain broke: only hand out what is already accepted
	End pending_end = kRunning;
	uint64_t expect_bit;                       // where the next accepted chunk must start
	uint8_t win[kWin];                         // last bytes of accepted output, valid: the final win_have bytes
	size_t win_have = 0;
	std::vector<uint8_t> win_out;              // window() for the caller after a bail
	size_t end_byte = 0;
	uint64_t accepted = 0;
	Task *lent = nullptr;                      // the task whose bytes the caller is reading
	size_t max_resolving;

	void worker_main()
	{
		std::unique_ptr<Inflater> inf(new Inflater());
		for (;;) {
			Task *t;
			{
				std::unique_lock<std::mutex> g(mu);
				cv_work.wait(g, [&] { return stop || !todo.empty(); });
				if (stop) return;
				t = todo.front();
				todo.pop_front();
			}
			if (t->job == 0) decode(*inf, *t);
			else resolve(*t);
			{
				std::lock_guard<std::mutex> g(mu);
				if (t->job == 0) t->decoded = true;
				else t->resolved = true;
			}
			cv_done.notify_all();
		}
	}

	// symbols -> bytes in place (byte i lands at or before symbol i's own storage)
	static void resolve(Task &t)
	{
		uint8_t lut[256 + kWin];
		for (int i = 0; i < 256; ++i) lut[i] = (uint8_t)i;
		memcpy(lut + 256, t.window, kWin);
		const uint16_t *sy = t.sym.get() + t.n_prefix;
		uint8_t *by = reinterpret_cast<uint8_t *>(t.sym.get() + t.n_prefix);
		for (size_t i = 0; i < t.n_sym; ++i) by[i] = lut[sy[i]];
		t.crc = crc32_fast(0, by, t.n_sym);                        // the consumer only has to combine these
	}

	// One attempt from `bit`; false = this was not a block start (or the data is bad).  Decodes straight
	// into the task's own symbol buffer, which is kept from chunk to chunk: a fresh buffer of this size
	// costs more in page faults than the decoding itself, a warm one is the fastest place to decode to.
	bool attempt(Inflater &inf, Task &t, uint64_t bit)
	{
		const uint8_t *end = base + size;
		t.n_prefix = t.exact_start ? 0 : kWin;                     // symbols in front of the output a match may reach
		const size_t want = t.n_prefix + chunk * 5 + Inflater::kSlack;
		if (t.sym_cap < want) {
			t.sym.reset(new uint16_t[want]);                       // uninitialised on purpose
			t.sym_cap = want;
		}
		for (size_t i = 0; i < t.n_prefix; ++i) t.sym[i] = (uint16_t)(256 + i);
		size_t pos = t.n_prefix;
		inf.begin_bits(base, bit, end);
		inf.stop_at_block_boundary(base, t.stop_bit);
		for (;;) {
			uint16_t *o = t.sym.get() + pos;
			const Inflater::Status st = inf.run16(t.sym.get(), &o, t.sym.get() + t.sym_cap - Inflater::kSlack);
			pos = (size_t)(o - t.sym.get());
			if (st == Inflater::kError) return false;
			if (pos >= 0x7FFFFFF0ull) return false;
			if (st == Inflater::kNeedOutput) {                     // the data expands more than 5x: a bigger buffer
				std::unique_ptr<uint16_t[]> bigger(new uint16_t[t.sym_cap * 2]);
				memcpy(bigger.get(), t.sym.get(), pos * sizeof(uint16_t));
				t.sym.swap(bigger);
				t.sym_cap *= 2;
				continue;
			}
			t.stream_end = st == Inflater::kStreamEnd;
			t.end_bit = inf.bit_position(base);
			t.end_byte = (size_t)(inf.in_pos() - base);
			break;
		}
		t.n_sym = pos - t.n_prefix;
		t.start_bit = bit;
		return true;
	}

	void decode(Inflater &inf, Task &t)
	{
		t.ok = false;
		if (t.exact_start) {
			t.ok = attempt(inf, t, t.from_bit);
			return;
		}
		const uint8_t *end = base + size;
		const uint64_t limit = std::min<uint64_t>(t.from_bit + kSearchChunks * chunk * 8, (uint64_t)size * 8);
		int tries = 0;
		for (uint64_t bit = t.from_bit; bit < limit; ++bit) {
			bit = Inflater::find_plausible_dynamic_header(base, bit, limit, end);
			if (bit == ~0ull) return;
			if (attempt(inf, t, bit)) {
				t.ok = true;
				return;
			}
			// about one random position in ten thousand passes the cheap test and nearly all of those die in
			// the header parser within microseconds; a chunk that burns through this many is not worth it
			if (++tries > 20000) return;
			if ((tries & 63) == 0) {
				std::lock_guard<std::mutex> g(mu);
				if (stop) return;
			}
		}
	}

	void issue_locked()
	{
		while (!no_more_accepts && next_issue < n_chunks && !free_tasks.empty()) {
			Task *t = free_tasks.back();
			free_tasks.pop_back();
			t->index = next_issue;
			t->exact_start = next_issue == 0;
			t->from_bit = (uint64_t)(deflate_start + next_issue * chunk) * 8;
			t->stop_bit = (uint64_t)(deflate_start + (next_issue + 1) * chunk) * 8;
			t->decoded = t->resolved = false;
			t->job = 0;
			todo.push_back(t);
			order.push_back(t);
			++next_issue;
			cv_work.notify_one();
		}
	}

	void shutdown()
	{
		{
			std::lock_guard<std::mutex> g(mu);
			stop = true;
		}
		cv_work.notify_all();
		for (auto &w : workers) w.join();
		workers.clear();
	}

	// Stitcher, with mu held: take decoded chunks in order while they line up; each accepted chunk gets the
	// window in front of it and goes back to the workers to be turned into bytes, and its own last 32 KiB
	// are resolved right here to become the next chunk's window (the only serial dependency).
	void accept_ready_locked()
	{
		while (!no_more_accepts && !order.empty() && order.front()->decoded && resolving.size() < max_resolving) {
			Task *t = order.front();
			bool good = t->ok && t->start_bit == expect_bit;
			uint8_t tail[kWin];
			size_t n_tail = 0;
			if (good) {
				n_tail = std::min<size_t>(kWin, t->n_sym);
				const uint16_t *sy = t->sym.get() + t->n_prefix + (t->n_sym - n_tail);
				for (size_t i = 0; i < n_tail; ++i) {
					const int b = resolve_one(sy[i], win, win_have);
					if (b < 0) {
						good = false;                              // a match reaching before the start of the stream
						break;
					}
					tail[i] = (uint8_t)b;
				}
				// the body may also hold such references: they all point into the window's missing front
				if (good && win_have < kWin) {
					const uint16_t lim = (uint16_t)(256 + (kWin - win_have));
					const uint16_t *body = t->sym.get() + t->n_prefix;
					for (size_t i = 0; i + n_tail < t->n_sym; ++i)
						if (body[i] >= 256 && body[i] < lim) {
							good = false;
							break;
						}
				}
			}
			if (!good) {                                           // the chain ends at expect_bit
				no_more_accepts = true;
				pending_end = kBail;
				break;
			}
			order.pop_front();
			memcpy(t->window, win, kWin);
			t->window_have = win_have;
			// next window = (old window + this chunk)'s last 32 KiB
			if (n_tail == kWin) {
				memcpy(win, tail, kWin);
				win_have = kWin;
			} else if (n_tail) {
				memmove(win, win + n_tail, kWin - n_tail);
				memcpy(win + kWin - n_tail, tail, n_tail);
				win_have = std::min(kWin, win_have + n_tail);
			}
			expect_bit = t->end_bit;
			++accepted;
			t->job = 1;
			todo.push_front(t);
			resolving.push_back(t);
			cv_work.notify_one();
			if (t->stream_end) {
				no_more_accepts = true;
				pending_end = kStreamEnd;
				end_byte = t->end_byte;
			}
		}
	}
};

ParallelInflate::ParallelInflate(const uint8_t *base, size_t size, size_t deflate_start, int workers, size_t chunk_bytes) : p_(new Impl())
{
	Impl &s = *p_;
	s.base = base;
	s.size = size;
	s.deflate_start = deflate_start;
	s.chunk = std::max<size_t>(chunk_bytes, 4096);
	s.n_chunks = (size - deflate_start + s.chunk - 1) / s.chunk;
	s.expect_bit = (uint64_t)deflate_start * 8;
	workers = std::max(1, workers);
	s.max_resolving = (size_t)workers / 2 + 1;
	const int n_tasks = 2 * workers + 2;                       // each holds one chunk's symbols: ~10 MB at 1 MiB chunks of FASTQ
	for (int i = 0; i < n_tasks; ++i) {
		s.pool.emplace_back(new Task());
		s.free_tasks.push_back(s.pool.back().get());
	}
	for (int i = 0; i < workers; ++i) s.workers.emplace_back([&s] { s.worker_main(); });
	std::lock_guard<std::mutex> g(s.mu);
	s.issue_locked();
}

ParallelInflate::~ParallelInflate() { p_->shutdown(); }

ParallelInflate::End ParallelInflate::end() const { return p_->end; }
size_t ParallelInflate::end_byte() const { return p_->end_byte; }
uint64_t ParallelInflate::resume_bit() const { return p_->expect_bit; }
const std::vector<uint8_t> &ParallelInflate::window() const { return p_->win_out; }
uint64_t ParallelInflate::chunks_accepted() const { return p_->accepted; }

bool ParallelInflate::next(const uint8_t **p, size_t *n, uint32_t *crc)
{
	Impl &s = *p_;
	std::unique_lock<std::mutex> g(s.mu);
	if (s.lent) {                                              // the caller is done with the previous chunk's bytes
		s.free_tasks.push_back(s.lent);
		s.lent = nullptr;
		s.issue_locked();
	}
	for (;;) {
		if (s.end != kRunning) return false;
		s.accept_ready_locked();
		if (!s.resolving.empty()) {
			Task *t = s.resolving.front();
			if (!t->resolved) {
				s.cv_done.wait(g, [&] { return t->resolved || (s.resolving.size() < s.max_resolving && !s.no_more_accepts && !s.order.empty() && s.order.front()->decoded); });
				continue;
			}
			s.resolving.pop_front();
			if (t->n_sym == 0) {
				s.free_tasks.push_back(t);
				s.issue_locked();
				continue;
			}
			s.lent = t;
			*p = reinterpret_cast<const uint8_t *>(t->sym.get() + t->n_prefix);
			*n = t->n_sym;
			if (crc) *crc = t->crc;
			return true;
		}
		if (s.no_more_accepts || s.order.empty()) {
			// everything accepted has been handed out.  Stream end, a broken chain, or chunks exhausted with
			// the last block still open (the plain decoder will say what is wrong with the data)
			s.end = s.pending_end == kRunning ? kBail : s.pending_end;
			s.win_out.assign(s.win + (kWin - s.win_have), s.win + kWin);
			s.stop = true;
			s.cv_work.notify_all();
			return false;
		}
		Task *f = s.order.front();
		s.cv_done.wait(g, [&] { return f->decoded; });
	}
}

}  // namespace ntsm
