// parse_worker.cpp -- bin/ntsm_parse_worker: one parser of FingerPrint::computeCounts (src/FingerPrint.hpp:46-87)
// as a process of its own (see procpipe.h for why).  Started by ntsm_count_files with the shared mapping's file
// descriptor; takes files off the shared counter, reads them with the same FastxReader (kseq grammar,
// vendor/kseq.h:178-219) and packs them with the same Packer as the in-process parser threads (pipeline.cpp),
// into slots of the shared, page-locked mapping.  No CUDA here.
//   usage: ntsm_parse_worker <shm fd> <worker index>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include <algorithm>

#include "../../include/ntsm_b200.h"
#include "fastx.h"
#include "pack.h"
#include "procpipe.h"

using namespace ntsm;

namespace {

void nap()
{
	struct timespec ts = { 0, 20000 };     // 20 us
	nanosleep(&ts, nullptr);
}

struct Out {
	ProcHeader *h;
	uint8_t *base;
	ProcSlot *slots;
	uint32_t me;
	int cur = -1;                          // slot being filled
	Packer pk;
	uint64_t n_bases = 0, n_reads = 0;

	bool claim()
	{
		for (;;) {
			if (h->stop.load(std::memory_order_acquire)) return false;
			for (uint32_t i = 0; i < h->n_slots; ++i) {
				const uint32_t s = (i + me) % h->n_slots;      // workers start their scan at different slots
				uint32_t want = kSlotFree;
				if (slots[s].state.load(std::memory_order_relaxed) == kSlotFree &&
				    slots[s].state.compare_exchange_strong(want, kSlotFilling, std::memory_order_acquire)) {
					cur = (int)s;
					slots[s].worker = me;
					uint8_t *b = base + h->data_off + (uint64_t)s * h->slot_stride;
					pk.reset_streaming(reinterpret_cast<uint64_t *>(b), reinterpret_cast<uint32_t *>(b + h->bases_bytes));
					n_bases = n_reads = 0;
					return true;
				}
			}
			nap();
		}
	}
	void publish()
	{
		if (cur < 0) return;
		ProcSlot &s = slots[cur];
		if (pk.pos == 0) {
			s.state.store(kSlotFree, std::memory_order_release);
		} else {
			s.n_pos = pk.finish();             // pads, streams the last lines out, sfence
			s.n_bases = n_bases;
			s.n_reads = n_reads;
			s.state.store(kSlotReady, std::memory_order_release);
		}
		cur = -1;
	}
	// FingerPrint::insertCount(seq, len) for this parser: the packing half of ntsm_batch_append (ctx.cu) without the
	// -m bookkeeping (a cap keeps the parsers in-process): a read that does not fit is split, its last k-1 bases
	// re-packed in the next slot, so every window is counted exactly once
	bool append(const char *seq, uint64_t len)
	{
		const uint32_t k = h->k;
		uint64_t pos = 0;
		for (;;) {
			if (cur < 0 && !claim()) return false;
			const uint64_t from = pos >= (uint64_t)(k - 1) ? pos - (k - 1) : 0;
			const uint64_t start = pos == 0 ? 0 : from;
			const uint64_t need = len - start;
			const uint64_t room = h->cap_pos - pk.pos;
			if (read_span(need) <= room) {
				pk.put_read(seq + start, need);
				n_bases += len - pos;
				n_reads += (pos == 0);
				return true;
			}
			const uint64_t split_min = std::max<uint64_t>(2 * k, std::min<uint64_t>(4096, h->cap_pos / 4));
			if (pk.pos != 0 && room < split_min + 1) {             // full: hand it over and go on in a fresh slot
				publish();
				continue;
			}
			const uint64_t take = room - 1;
			pk.put_read(seq + start, take);
			n_bases += start + take - pos;
			n_reads += (pos == 0);
			pos = start + take;
			publish();
		}
	}
};

void set_error(ProcHeader *h, int code, const char *text)
{
	uint32_t want = 0;
	if (h->error_lock.compare_exchange_strong(want, 1)) {
		snprintf(h->error_text, sizeof h->error_text, "%s", text);
		h->error.store(code, std::memory_order_release);
	}
	h->stop.store(1, std::memory_order_release);
}

}  // namespace

int main(int argc, char **argv)
{
	if (argc < 3) return 64;
	const int fd = atoi(argv[1]);
	const uint32_t me = (uint32_t)atoi(argv[2]);
	struct stat sb;
	if (fstat(fd, &sb) != 0 || sb.st_size < (off_t)sizeof(ProcHeader)) return 65;
	void *m = mmap(nullptr, (size_t)sb.st_size, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
	if (m == MAP_FAILED) return 66;
	ProcHeader *h = static_cast<ProcHeader *>(m);
	if (h->magic != kProcMagic || h->version != kProcVersion || h->total_bytes != (uint64_t)sb.st_size) return 67;
	uint8_t *base = static_cast<uint8_t *>(m);
	Out out;
	out.h = h;
	out.base = base;
	out.slots = reinterpret_cast<ProcSlot *>(base + h->slots_off);
	out.me = me;
	const uint64_t *path_off = reinterpret_cast<const uint64_t *>(base + h->paths_off);
	FastxReader rd;
	for (;;) {
		if (h->stop.load(std::memory_order_acquire)) break;
		const uint32_t fi = h->next_file.fetch_add(1);
		if (fi >= h->n_files) break;
		const char *path = reinterpret_cast<const char *>(base + path_off[fi]);
		if (!rd.open(path, 0, true)) {                               // src/FingerPrint.hpp:51-57
			char msg[480];
			snprintf(msg, sizeof msg, "file %s cannot be opened", path);
			set_error(h, NTSM_ERR_IO, msg);
			break;
		}
		if (h->verbose) fprintf(stderr, "Opening %s\n", path);       // :58-62
		int64_t l;
		bool ok = true;
		while (ok && (l = rd.next()) >= 0) ok = out.append(rd.seq(), (uint64_t)l);   // :67 (any negative code ends the file)
		rd.close();
		if (!ok) break;
	}
	out.publish();
	return 0;
}
