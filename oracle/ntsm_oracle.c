/*
 * ntsm_oracle.c -- CPU restatement of the ntsmCount counting path.
 * TEST INFRASTRUCTURE ONLY (see ntsm_oracle.h).  Parity: PINNED against the
 * real reference binary (tests/golden/, oracle/_ref/ntsmCount).
 *
 * Plain C99 + zlib (+ OpenMP for the bulk helper).  Written from the behaviour
 * of the reference, not from its text: every block cites the file:line under
 * /root/reference that it restates.
 */
#include "ntsm_oracle.h"

#include <errno.h>
#include <stdlib.h>
#include <string.h>
#include <sys/types.h>
#include <zlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ */
/* base decoding: vendor/KseqHashIterator.hpp:114-127                  */
/* A,a->0  C,c->1  G,g->2  T,t,U,u->3  bytes 0..3 -> themselves  else 4 */
int ntsm_oracle_nt4(unsigned char c)
{
	switch (c) {
	case 0: case 'A': case 'a': return 0;
	case 1: case 'C': case 'c': return 1;
	case 2: case 'G': case 'g': return 2;
	case 3: case 'T': case 't': case 'U': case 'u': return 3;
	default: return 4;
	}
}

/* vendor/KseqHashIterator.hpp:29 -- mask = (1<<2k)-1 (k<=31 here; k=32 is UB upstream) */
static uint64_t kmask(unsigned k) { return (1ULL << (2 * k)) - 1; }

/* vendor/KseqHashIterator.hpp:129-139 */
uint64_t ntsm_oracle_hash64(uint64_t x, unsigned k)
{
	const uint64_t m = kmask(k);
	x = (~x + (x << 21)) & m;       /* :131 */
	x ^= x >> 24;                   /* :132 */
	x = (x + (x << 3) + (x << 8)) & m; /* :133  (*265) */
	x ^= x >> 14;                   /* :134 */
	x = (x + (x << 2) + (x << 4)) & m; /* :135  (*21) */
	x ^= x >> 28;                   /* :136 */
	x = (x + (x << 31)) & m;        /* :137 */
	return x;
}

/* vendor/KseqHashIterator.hpp:28-33,87-112: the iterator as a plain loop.
 * A k-mer is yielded at every position where the last k decoded bytes were
 * all <4; fw shifts in at the low end, rv at bit 2(k-1); key=min(fw,rv). */
typedef void (*kmer_cb)(void *ctx, uint64_t hv, uint64_t pos, uint64_t fw, uint64_t rv);

static size_t iterate(const char *seq, uint64_t len, unsigned k, kmer_cb cb, void *ctx)
{
	const uint64_t m = kmask(k);
	const unsigned shift = 2 * (k - 1);
	uint64_t fw = 0, rv = 0;
	unsigned run = 0;
	size_t n = 0;
	for (uint64_t p = 0; p < len; ++p) {
		int c = ntsm_oracle_nt4((unsigned char)seq[p]);
		if (c < 4) {
			fw = ((fw << 2) | (uint64_t)c) & m;               /* :99  */
			rv = (rv >> 2) | ((uint64_t)(3 - c) << shift);    /* :100 */
			if (++run >= k) {                                 /* :101 */
				++n;
				if (cb) cb(ctx, ntsm_oracle_hash64(fw < rv ? fw : rv, k), p + 1, fw, rv);
			}
		} else {                                              /* :106-107 */
			fw = rv = 0;
			run = 0;
		}
	}
	return n;
}

struct iter_out { uint64_t *h, *p, *fw, *rv; size_t cap, n; };
static void iter_store(void *c, uint64_t hv, uint64_t pos, uint64_t fw, uint64_t rv)
{
	struct iter_out *o = (struct iter_out *)c;
	if (o->n < o->cap) {
		if (o->h) o->h[o->n] = hv;
		if (o->p) o->p[o->n] = pos;
		if (o->fw) o->fw[o->n] = fw;
		if (o->rv) o->rv[o->n] = rv;
	}
	o->n++;
}

size_t ntsm_oracle_iter(const char *seq, uint64_t len, unsigned k, uint64_t *out_hash,
                        uint64_t *out_pos, uint64_t *out_fw, uint64_t *out_rv, size_t cap)
{
	struct iter_out o = { out_hash, out_pos, out_fw, out_rv, cap, 0 };
	return iterate(seq, len, k, iter_store, &o);
}

/* ------------------------------------------------------------------ */
/* record reader: vendor/kseq.h:68-79 (getc), :94-146 (getuntil2),     */
/* :178-219 (kseq_read); 16 KiB gzread buffer (:229)                   */
struct ntsm_oracle_reader {
	gzFile f;
	unsigned char buf[16384];
	int beg, end, eof, err;
	int last;                 /* kseq_t::last_char */
	char *name, *seq, *qual;  /* growable strings */
	size_t nl, nm, sl, sm, ql, qm;
};

static int rd_fill(ntsm_oracle_reader *r)
{   /* returns 1 if bytes are available */
	if (r->beg < r->end) return 1;
	if (r->eof) return 0;
	r->beg = 0;
	r->end = gzread(r->f, r->buf, sizeof r->buf);
	if (r->end == 0) { r->eof = 1; return 0; }
	if (r->end < 0) { r->eof = 1; r->err = 1; r->end = 0; return 0; }
	return 1;
}

/* ks_getc: next byte, -1 at EOF, -3 on stream error */
static int rd_getc(ntsm_oracle_reader *r)
{
	if (r->err) return -3;
	if (!rd_fill(r)) return r->err ? -3 : -1;
	return r->buf[r->beg++];
}

static void str_put(char **s, size_t *l, size_t *m, const unsigned char *src, size_t n)
{
	if (*l + n + 1 > *m) {
		size_t nm = *m ? *m : 256;
		while (nm < *l + n + 1) nm *= 2;
		*s = (char *)realloc(*s, nm);
		*m = nm;
	}
	if (n) memcpy(*s + *l, src, n);
	*l += n;
	(*s)[*l] = 0;
}

/* ks_getuntil2 for the two delimiters kseq_read uses.
 * mode 0: stop at any isspace() byte (KS_SEP_SPACE); mode 2: stop at '\n' and then
 * drop one trailing '\r' if the accumulated string is longer than 1 (kseq.h:141).
 * Returns string length, -1 if nothing could be read because of EOF, -3 on error.
 * *dret receives the delimiter byte (0 if the stream ended first). */
static long rd_until(ntsm_oracle_reader *r, int mode, char **s, size_t *l, size_t *m,
                     int append, int *dret)
{
	int got = 0;
	if (dret) *dret = 0;
	if (!append) *l = 0;
	for (;;) {
		if (r->err) return -3;
		if (!rd_fill(r)) { if (r->err) return -3; break; }
		int i = r->beg;
		if (mode == 2) {
			while (i < r->end && r->buf[i] != '\n') ++i;
		} else {
			while (i < r->end) {
				unsigned char c = r->buf[i];
				if (c == ' ' || (c >= '\t' && c <= '\r')) break; /* isspace in the C locale */
				++i;
			}
		}
		got = 1;
		str_put(s, l, m, r->buf + r->beg, (size_t)(i - r->beg));
		r->beg = i + 1;
		if (i < r->end) { if (dret) *dret = r->buf[i]; break; }
	}
	if (!got && r->eof && r->beg >= r->end) return -1;
	if (*s == NULL) str_put(s, l, m, NULL, 0);
	else if (mode == 2 && *l > 1 && (*s)[*l - 1] == '\r') { --*l; }
	(*s)[*l] = 0;
	return (long)*l;
}

ntsm_oracle_reader *ntsm_oracle_reader_open(const char *path)
{
	gzFile f = gzopen(path, "r");
	if (!f) return NULL;
	ntsm_oracle_reader *r = (ntsm_oracle_reader *)calloc(1, sizeof *r);
	r->f = f;
	return r;
}

void ntsm_oracle_reader_close(ntsm_oracle_reader *r)
{
	if (!r) return;
	gzclose(r->f);
	free(r->name); free(r->seq); free(r->qual);
	free(r);
}

long ntsm_oracle_reader_next(ntsm_oracle_reader *r, const char **seq, const char **name)
{
	int c;
	long rc;
	if (r->last == 0) {                       /* kseq.h:182-186: hunt for a header */
		while ((c = rd_getc(r)) >= 0 && c != '>' && c != '@') {}
		if (c < 0) return c;
		r->last = c;
	}
	r->sl = r->ql = 0;
	if ((rc = rd_until(r, 0, &r->name, &r->nl, &r->nm, 0, &c)) < 0) return rc;   /* :188 */
	if (c != '\n') {                          /* :189 rest of header line = comment (discarded) */
		char *tmp = NULL; size_t tl = 0, tm = 0;
		rd_until(r, 2, &tmp, &tl, &tm, 0, NULL);
		free(tmp);
	}
	if (!r->seq) str_put(&r->seq, &r->sl, &r->sm, NULL, 0);
	while ((c = rd_getc(r)) >= 0 && c != '>' && c != '+' && c != '@') {      /* :194 */
		if (c == '\n') continue;                                              /* :195 */
		unsigned char ch = (unsigned char)c;
		str_put(&r->seq, &r->sl, &r->sm, &ch, 1);                             /* :196 */
		rd_until(r, 2, &r->seq, &r->sl, &r->sm, 1, NULL);                     /* :197 */
	}
	if (c == '>' || c == '@') r->last = c;                                    /* :199 */
	r->seq[r->sl] = 0;
	if (seq) *seq = r->seq;
	if (name) *name = r->name;
	if (c != '+') return (long)r->sl;                                         /* :206-207 FASTA */
	while ((c = rd_getc(r)) >= 0 && c != '\n') {}                             /* :212 */
	if (c == -1) return -2;                                                   /* :213 */
	if (!r->qual) str_put(&r->qual, &r->ql, &r->qm, NULL, 0);
	while (rd_until(r, 2, &r->qual, &r->ql, &r->qm, 1, NULL) >= 0 && r->ql < r->sl) {} /* :214 */
	r->last = 0;                                                              /* :216 */
	if (r->sl != r->ql) return -2;                                            /* :217 */
	return (long)r->sl;
}

/* ------------------------------------------------------------------ */
/* m_counts: hashed k-mer -> count.  Any exact map gives the same      */
/* observable behaviour as tsl::robin_map (src/FingerPrint.hpp:466).   */
typedef struct { uint64_t key; uint64_t cnt; uint8_t state; /*0 empty,1 live,2 erased*/ } slot_t;
typedef struct { slot_t *s; uint64_t cap, live; } map_t;

static uint64_t mix(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 29; return x; }

static void map_init(map_t *m, uint64_t cap) { m->cap = cap; m->live = 0; m->s = (slot_t *)calloc(cap, sizeof(slot_t)); }
static slot_t *map_find_any(const map_t *m, uint64_t key)
{   /* finds live or erased entry with this key */
	uint64_t i = mix(key) & (m->cap - 1);
	while (m->s[i].state) {
		if (m->s[i].key == key) return &m->s[i];
		i = (i + 1) & (m->cap - 1);
	}
	return NULL;
}
static slot_t *map_find(const map_t *m, uint64_t key)
{
	slot_t *s = map_find_any(m, key);
	return (s && s->state == 1) ? s : NULL;
}
static void map_grow(map_t *m);
static slot_t *map_insert(map_t *m, uint64_t key)
{
	if ((m->live + 1) * 2 > m->cap) map_grow(m);
	uint64_t i = mix(key) & (m->cap - 1);
	while (m->s[i].state) i = (i + 1) & (m->cap - 1);
	m->s[i].key = key; m->s[i].cnt = 0; m->s[i].state = 1;
	m->live++;
	return &m->s[i];
}
static void map_grow(map_t *m)
{
	map_t n;
	map_init(&n, m->cap * 2);
	for (uint64_t i = 0; i < m->cap; ++i)
		if (m->s[i].state) {
			slot_t *d = map_insert(&n, m->s[i].key);
			d->cnt = m->s[i].cnt; d->state = m->s[i].state;
		}
	free(m->s);
	*m = n;
}

typedef struct { uint64_t *v; uint32_t n, cap; int present; } list_t;
static void list_push(list_t *l, uint64_t x)
{
	if (l->n == l->cap) { l->cap = l->cap ? l->cap * 2 : 16; l->v = (uint64_t *)realloc(l->v, l->cap * sizeof(uint64_t)); }
	l->v[l->n++] = x;
}

struct ntsm_oracle_fp {
	unsigned k;
	map_t counts;             /* m_counts */
	uint64_t table_size;      /* m_counts.size() after dupe removal */
	list_t *ref, *var;        /* m_alleleIDToKmerRef / Var */
	char **names;             /* m_alleleIDs */
	uint32_t n_ref, n_var, cap_sites;
	uint64_t total_counts, total_kmers, total_bases, max_counts;
	int early;
};

struct site_cb { ntsm_oracle_fp *fp; list_t *list; const char *name; const char *tag; FILE *warn; map_t *dupes; };
static void site_kmer(void *c, uint64_t hv, uint64_t pos, uint64_t fw, uint64_t rv)
{   /* src/FingerPrint.hpp:516-528 (REF) / :537-549 (VAR) */
	struct site_cb *s = (struct site_cb *)c;
	(void)fw; (void)rv;
	if (map_find_any(&s->fp->counts, hv)) {
		if (s->warn)
			fprintf(s->warn, "Warning: %s of %s file has a k-mer collision at pos: %llu\n",
			        s->name, s->tag, (unsigned long long)pos);
		if (!map_find_any(s->dupes, hv)) map_insert(s->dupes, hv);
	} else {
		list_push(s->list, hv);
		map_insert(&s->fp->counts, hv);
	}
}

ntsm_oracle_fp *ntsm_oracle_fp_create(const char *sites_path, unsigned k, int dupes,
                                      double cov_thresh, FILE *warn)
{
	ntsm_oracle_reader *r = ntsm_oracle_reader_open(sites_path);
	if (!r) return NULL;
	ntsm_oracle_fp *fp = (ntsm_oracle_fp *)calloc(1, sizeof *fp);
	fp->k = k;
	map_init(&fp->counts, 1 << 16);
	map_t du; map_init(&du, 1 << 10);
	const char *seq, *name;
	long l;
	uint64_t entry = 0;
	while ((l = ntsm_oracle_reader_next(r, &seq, &name)) >= 0) {   /* :508-553 */
		uint32_t idx = (uint32_t)(entry / 2);
		if (idx >= fp->cap_sites) {
			uint32_t nc = fp->cap_sites ? fp->cap_sites * 2 : 1024;
			fp->ref = (list_t *)realloc(fp->ref, nc * sizeof(list_t));
			fp->var = (list_t *)realloc(fp->var, nc * sizeof(list_t));
			fp->names = (char **)realloc(fp->names, nc * sizeof(char *));
			memset(fp->ref + fp->cap_sites, 0, (nc - fp->cap_sites) * sizeof(list_t));
			memset(fp->var + fp->cap_sites, 0, (nc - fp->cap_sites) * sizeof(list_t));
			fp->cap_sites = nc;
		}
		struct site_cb cb = { fp, NULL, name, NULL, warn, &du };
		if (entry % 2 == 0) {
			cb.list = &fp->ref[idx]; cb.tag = "REF";
			fp->ref[idx].present = 1;
			iterate(seq, (uint64_t)l, k, site_kmer, &cb);
			fp->names[idx] = strdup(name);                         /* :530 */
			fp->n_ref = idx + 1;
		} else {
			cb.list = &fp->var[idx]; cb.tag = "VAR";
			fp->var[idx].present = 1;
			iterate(seq, (uint64_t)l, k, site_kmer, &cb);
			fp->n_var = idx + 1;
		}
		entry++;
	}
	ntsm_oracle_reader_close(r);
	fp->table_size = fp->counts.live;
	if (!dupes) {                                                  /* :557-563 */
		for (uint64_t i = 0; i < du.cap; ++i)
			if (du.s[i].state) {
				slot_t *s = map_find(&fp->counts, du.s[i].key);
				if (s) { s->state = 2; fp->table_size--; }
			}
	}
	free(du.s);
	if (cov_thresh > 0)                                            /* :41-43 */
		fp->max_counts = (uint64_t)(((double)fp->table_size * cov_thresh) / 2);
	return fp;
}

void ntsm_oracle_fp_destroy(ntsm_oracle_fp *fp)
{
	if (!fp) return;
	for (uint32_t i = 0; i < fp->cap_sites; ++i) { free(fp->ref[i].v); free(fp->var[i].v); }
	for (uint32_t i = 0; i < fp->n_ref; ++i) free(fp->names[i]);
	free(fp->ref); free(fp->var); free(fp->names); free(fp->counts.s);
	free(fp);
}

struct cnt_cb { ntsm_oracle_fp *fp; uint64_t hits, kmers; int atomic; };
static void count_kmer(void *c, uint64_t hv, uint64_t pos, uint64_t fw, uint64_t rv)
{   /* src/FingerPrint.hpp:92-99 */
	struct cnt_cb *s = (struct cnt_cb *)c;
	(void)pos; (void)fw; (void)rv;
	slot_t *e = map_find(&s->fp->counts, hv);
	if (e) {
		if (s->atomic) {
#pragma omp atomic update
			e->cnt += 1;
		} else e->cnt += 1;
		s->hits++;
	}
	s->kmers++;
}

void ntsm_oracle_fp_insert(ntsm_oracle_fp *fp, const char *seq, uint64_t len)
{
	struct cnt_cb cb = { fp, 0, 0, 0 };
	iterate(seq, len, fp->k, count_kmer, &cb);
	fp->total_counts += cb.hits;
	fp->total_kmers += cb.kmers;
	fp->total_bases += len;                                        /* :101-102 */
}

void ntsm_oracle_fp_insert_many(ntsm_oracle_fp *fp, const char *buf, const uint64_t *off,
                                uint64_t n_reads, int threads)
{
	uint64_t hits = 0, kmers = 0, bases = 0;
#ifdef _OPENMP
	if (threads < 1) threads = 1;
#pragma omp parallel num_threads(threads) reduction(+ : hits, kmers, bases)
#endif
	{
		struct cnt_cb cb = { fp, 0, 0, 1 };
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 4096)
#endif
		for (long long i = 0; i < (long long)n_reads; ++i) {
			iterate(buf + off[i], off[i + 1] - off[i], fp->k, count_kmer, &cb);
			bases += off[i + 1] - off[i];
		}
		hits += cb.hits;
		kmers += cb.kmers;
	}
	(void)threads;
	fp->total_counts += hits;
	fp->total_kmers += kmers;
	fp->total_bases += bases;
}

int ntsm_oracle_fp_count_file(ntsm_oracle_fp *fp, const char *path)
{
	ntsm_oracle_reader *r = ntsm_oracle_reader_open(path);
	if (!r) return -1;
	const char *seq;
	long l = ntsm_oracle_reader_next(r, &seq, NULL);
	while (l >= 0 && !fp->early) {                                 /* :67 */
		ntsm_oracle_fp_insert(fp, seq, (uint64_t)l);               /* :475 */
		if (fp->max_counts != 0 && fp->total_counts > fp->max_counts) fp->early = 1; /* :476,486 */
		l = ntsm_oracle_reader_next(r, &seq, NULL);
	}
	ntsm_oracle_reader_close(r);
	return 0;
}

static int list_stats(const ntsm_oracle_fp *fp, const list_t *l, uint32_t *mx, uint32_t *sum)
{   /* src/FingerPrint.hpp:281-287 -- `unsigned` arithmetic, i.e. mod 2^32 */
	uint32_t m = 0, s = 0;
	for (uint32_t j = 0; j < l->n; ++j) {
		const slot_t *e = map_find(&fp->counts, l->v[j]);
		if (!e) return -134;                                       /* m_counts.at() throws */
		uint32_t c = (uint32_t)e->cnt;
		if (m < c) m = c;
		s += c;
	}
	*mx = m; *sum = s;
	return 0;
}

int ntsm_oracle_fp_rows(const ntsm_oracle_fp *fp, uint32_t *max_ref, uint32_t *max_var,
                        uint32_t *sum_ref, uint32_t *sum_var, uint32_t *n_ref, uint32_t *n_var)
{
	for (uint32_t i = 0; i < fp->n_ref; ++i) {
		if (i >= fp->n_var) return -134;                           /* :276 vector::at throws */
		if (list_stats(fp, &fp->ref[i], &max_ref[i], &sum_ref[i])) return -134;
		if (list_stats(fp, &fp->var[i], &max_var[i], &sum_var[i])) return -134;
		n_ref[i] = fp->ref[i].n; n_var[i] = fp->var[i].n;
	}
	return 0;
}

int ntsm_oracle_fp_print(const ntsm_oracle_fp *fp, FILE *out)
{
	fprintf(out, "#@TK\t%llu\n#@KS\t%u", (unsigned long long)fp->total_kmers, fp->k); /* :261-268 */
	fprintf(out, "\n#locusID\tcountAT\tcountCG\tsumAT\tsumCG\tdistinctAT\tdistinctCG\n"); /* :271 */
	for (uint32_t i = 0; i < fp->n_ref; ++i) {
		uint32_t mr, mv, sr, sv;
		if (i >= fp->n_var) return -134;
		if (list_stats(fp, &fp->ref[i], &mr, &sr)) return -134;
		if (list_stats(fp, &fp->var[i], &mv, &sv)) return -134;
		fprintf(out, "%s\t%u\t%u\t%u\t%u\t%u\t%u\n", fp->names[i], mr, mv, sr, sv,
		        fp->ref[i].n, fp->var[i].n);                       /* :295-309 */
	}
	return 0;
}

int ntsm_oracle_fp_summary(ntsm_oracle_fp *fp, char *buf, size_t cap)
{   /* src/FingerPrint.hpp:313-333 + getSitesCoveredInSample :389-413 (operator[] there: an
	   erased k-mer reads as 0 instead of throwing) */
	unsigned covered = 0;
	for (uint32_t i = 0; i < fp->n_ref; ++i) {
		int any = 0;
		const list_t *ls[2] = { &fp->ref[i], i < fp->n_var ? &fp->var[i] : NULL };
		for (int a = 0; a < 2 && ls[a]; ++a)
			for (uint32_t j = 0; j < ls[a]->n; ++j) {
				const slot_t *e = map_find(&fp->counts, ls[a]->v[j]);
				if (e && e->cnt > 0) any = 1;
			}
		covered += any;
	}
	return snprintf(buf, cap,
	                "Total Bases Considered: %llu\nTotal k-mers Considered: %llu\n"
	                "Total k-mers Recorded: %llu\nDistinct k-mers in initial set: %llu\n"
	                "Total Sites: %u\nSites Covered by at least one k-mer: %u\n",
	                (unsigned long long)fp->total_bases, (unsigned long long)fp->total_kmers,
	                (unsigned long long)fp->total_counts, (unsigned long long)fp->table_size,
	                fp->n_ref, covered);
}

uint64_t ntsm_oracle_fp_total_kmers(const ntsm_oracle_fp *fp) { return fp->total_kmers; }
uint64_t ntsm_oracle_fp_total_counts(const ntsm_oracle_fp *fp) { return fp->total_counts; }
uint64_t ntsm_oracle_fp_total_bases(const ntsm_oracle_fp *fp) { return fp->total_bases; }
uint64_t ntsm_oracle_fp_max_counts(const ntsm_oracle_fp *fp) { return fp->max_counts; }
int ntsm_oracle_fp_early_term(const ntsm_oracle_fp *fp) { return fp->early; }
uint64_t ntsm_oracle_fp_table_size(const ntsm_oracle_fp *fp) { return fp->table_size; }
uint32_t ntsm_oracle_fp_n_sites(const ntsm_oracle_fp *fp) { return fp->n_ref; }
uint32_t ntsm_oracle_fp_n_listed(const ntsm_oracle_fp *fp)
{
	uint32_t n = 0;
	for (uint32_t i = 0; i < fp->n_ref; ++i) n += fp->ref[i].n + (i < fp->n_var ? fp->var[i].n : 0);
	return n;
}
const char *ntsm_oracle_fp_site_name(const ntsm_oracle_fp *fp, uint32_t i) { return fp->names[i]; }

void ntsm_oracle_fp_lists(const ntsm_oracle_fp *fp, uint64_t *hashes, uint32_t *allele_off,
                          uint32_t *counts)
{
	uint32_t n = 0;
	for (uint32_t i = 0; i < fp->n_ref; ++i) {
		const list_t *ls[2] = { &fp->ref[i], i < fp->n_var ? &fp->var[i] : NULL };
		for (int a = 0; a < 2; ++a) {
			if (allele_off) allele_off[2 * i + a] = n;
			if (!ls[a]) continue;
			for (uint32_t j = 0; j < ls[a]->n; ++j, ++n) {
				if (hashes) hashes[n] = ls[a]->v[j];
				if (counts) {
					const slot_t *e = map_find(&fp->counts, ls[a]->v[j]);
					counts[n] = e ? (uint32_t)e->cnt : 0xFFFFFFFFu;
				}
			}
		}
	}
	if (allele_off) allele_off[2 * fp->n_ref] = n;
}

/* ================================================================== */
/* Multi-sample matrix path: MultiCount driven by   */
/* VCFConvert.  Pinned against tools/ref_vcf_harness.cpp, which runs   */
/* the reference's two classes unmodified (tests/golden/vcf/).         */

/* MultiCount (src/MultiCount.hpp:36-289).  initCountsHash (:214-288) is FingerPrint's
 * (src/FingerPrint.hpp:490-564) with `m_kmerToHash[hv] = kmerCount++` instead of a zero count:
 * the dense index of a k-mer is its position in the site lists taken in file order.  The fp's
 * map is reused with slot.cnt holding that index. */
struct ntsm_oracle_mc {
	ntsm_oracle_fp *fp;
	uint32_t n_samples;
	uint64_t listed;          /* kmerCount (:217) */
	uint64_t stride;          /* m_kmerToHash.size() (:55,:108): the row stride */
	uint8_t *mat;             /* m_matCounts (:266): listed * n_samples bytes */
};

ntsm_oracle_mc *ntsm_oracle_mc_create(const char *sites_path, unsigned k, int dupes, uint32_t n_samples, FILE *warn)
{
	ntsm_oracle_fp *fp = ntsm_oracle_fp_create(sites_path, k, dupes, 0, warn);
	if (!fp) return NULL;
	ntsm_oracle_mc *mc = (ntsm_oracle_mc *)calloc(1, sizeof *mc);
	mc->fp = fp;
	mc->n_samples = n_samples;
	uint64_t n = 0;
	for (uint32_t i = 0; i < fp->n_ref; ++i) {
		const list_t *ls[2] = { &fp->ref[i], i < fp->n_var ? &fp->var[i] : NULL };
		for (int a = 0; a < 2 && ls[a]; ++a)
			for (uint32_t j = 0; j < ls[a]->n; ++j) map_find_any(&fp->counts, ls[a]->v[j])->cnt = n++;   /* :230,:250 */
	}
	mc->listed = n;
	mc->stride = fp->table_size;
	mc->mat = (uint8_t *)calloc(n * n_samples + 1, 1);
	return mc;
}

void ntsm_oracle_mc_destroy(ntsm_oracle_mc *mc)
{
	if (!mc) return;
	ntsm_oracle_fp_destroy(mc->fp);
	free(mc->mat);
	free(mc);
}

/* MultiCount::insertCount (:52-70), one thread: the first non-zero value written to a cell stays;
 * a later insert of a different value only warns.  The stored value is `multi` truncated to a byte,
 * the comparison is between the byte and the unsigned, and the byte goes to cerr as a CHARACTER. */
void ntsm_oracle_mc_insert(ntsm_oracle_mc *mc, unsigned sample, uint64_t hash, unsigned multi, FILE *warn)
{
	const slot_t *e = map_find(&mc->fp->counts, hash);
	if (!e) return;                                                /* :53 */
	uint8_t *cell = &mc->mat[mc->stride * sample + e->cnt];        /* :56-57 */
	if (*cell > 0) {                                               /* :58 */
		if (*cell != multi && warn) {                              /* :59-62 */
			fputs("Warning: Inconsistent k-mer counts, check for overlapping sites: ", warn);
			fputc(*cell, warn);
			fprintf(warn, " vs %u\n", multi);
		}
		return;
	}
	*cell = (uint8_t)multi;                                        /* :65-67 */
}

const uint8_t *ntsm_oracle_mc_matrix(const ntsm_oracle_mc *mc, uint64_t *bytes)
{
	if (bytes) *bytes = mc->listed * mc->n_samples;
	return mc->mat;
}
ntsm_oracle_fp *ntsm_oracle_mc_fp(ntsm_oracle_mc *mc) { return mc->fp; }

static int mc_list_max(const ntsm_oracle_mc *mc, const list_t *l, unsigned sample, uint32_t *mx, uint32_t *sum)
{
	uint32_t m = 0, s = 0;
	for (uint32_t j = 0; j < l->n; ++j) {
		const slot_t *e = map_find(&mc->fp->counts, l->v[j]);
		if (!e) return -134;                                       /* m_kmerToHash.at() throws (:108,:117,:169,:176) */
		const uint32_t c = mc->mat[mc->stride * sample + e->cnt];
		if (m < c) m = c;
		s += c;
	}
	*mx = m;
	*sum = s;
	return 0;
}

/* MultiCount::printCountsMax (:93-138): a counts file WITHOUT the #@TK / #@KS header lines */
int ntsm_oracle_mc_print_counts_max(const ntsm_oracle_mc *mc, unsigned index, FILE *out)
{
	const ntsm_oracle_fp *fp = mc->fp;
	fprintf(out, "\n#locusID\tcountAT\tcountCG\tsumAT\tsumCG\tdistinctAT\tdistinctCG\n");     /* :94 */
	for (uint32_t i = 0; i < fp->n_ref; ++i) {
		uint32_t mr, mv, sr, sv;
		if (i >= fp->n_var) return -134;                           /* :99 */
		if (mc_list_max(mc, &fp->ref[i], index, &mr, &sr)) return -134;
		if (mc_list_max(mc, &fp->var[i], index, &mv, &sv)) return -134;
		fprintf(out, "%s\t%u\t%u\t%u\t%u\t%u\t%u\n", fp->names[i], mr, mv, sr, sv, fp->ref[i].n, fp->var[i].n);
	}
	return 0;
}

/* MultiCount::printNormMatrix (:148-203).  value = maxREF / (maxREF + maxVAR) as doubles, missing
 * (both zero) = the site's mean over the samples that have one -- the sum runs in sample order and
 * is divided in long double by the number of ALL samples (:181,:185: `size` counts every sample).
 * ostream state carried by `out`: doubles print as %.6g until the first missing value has been
 * written with setprecision(19) (:190), from then on every value of that stream prints as %.19g. */
int ntsm_oracle_mc_print_norm_matrix(const ntsm_oracle_mc *mc, const char *const *sample_ids, FILE *out, FILE *center_file)
{
	const ntsm_oracle_fp *fp = mc->fp;
	const uint32_t S = mc->n_samples;
	double *values = (double *)malloc((S + 1) * sizeof(double));
	int precision = 6;
	fputs("alleleID", out);                                        /* :149 */
	for (uint32_t j = 0; j < S; ++j) fprintf(out, "\t%s", sample_ids[j]);
	fputc('\n', out);
	for (uint32_t i = 0; i < fp->n_ref; ++i) {
		if (i >= fp->n_var) { free(values); return -134; }         /* :158 */
		double sum = 0.0;
		uint64_t size = 0;
		for (uint32_t j = 0; j < S; ++j) {
			uint32_t mr, mv, unused;
			if (mc_list_max(mc, &fp->ref[i], j, &mr, &unused) || mc_list_max(mc, &fp->var[i], j, &mv, &unused)) { free(values); return -134; }
			const unsigned denom = mr + mv;                        /* :179 */
			if (denom == 0) values[j] = 1.7976931348623157e308;    /* UNDEF = numeric_limits<double>::max() (:41) */
			else {
				values[j] = (double)mr / (double)denom;            /* :183 */
				sum += values[j];
			}
			++size;
		}
		fputs(fp->names[i], out);                                  /* :187 */
		const long double size_f = (long double)size;
		const long double center = (long double)sum / size_f;      /* :188-189 */
		for (uint32_t j = 0; j < S; ++j) {
			if (values[j] == 1.7976931348623157e308) {
				precision = 19;                                    /* :192, and it sticks */
				fprintf(out, "\t%.19Lg", center);
			} else fprintf(out, "\t%.*g", precision, values[j]);   /* :194 */
		}
		fprintf(center_file, "%.19Lg\n", center);                  /* :198 */
		fputc('\n', out);
	}
	free(values);
	return 0;
}

/* VCFConvert (src/VCFConvert.hpp:40-218) with one thread.  Return codes: 0; -1 a file cannot be
 * opened; -134 where the reference dies on an uncaught exception or a failed assert (empty line
 * before the #CHROM line: string::at :74; unknown chromosome: robin_map::at :204; non-numeric POS:
 * stoi :113; a data line whose sample columns are not as many as the header's: assert :146);
 * -2 where the reference has undefined behaviour and no result to restate (a site closer than
 * window/2 to the start of its chromosome, or beyond its end: getSeqFromSite reads outside the
 * sequence, :207-209).
 * n_samples_out / sample_ids_out (malloc'ed array of malloc'ed strings) describe the header. */
struct chr_rec { char *name; char *seq; size_t len; };

/* Tab-separated fields of a line of `len` bytes (which may hold NUL bytes: std::string items do): pointers + lengths;
 * every field is also NUL-terminated in place for the calls that take C strings as the reference's do (stoi). */
static char **split_tabs(char *line, size_t len, size_t *n_out, size_t **len_out)
{
	size_t n = 1;
	for (size_t p = 0; p < len; ++p) n += line[p] == '\t';
	char **f = (char **)malloc(n * sizeof(char *));
	size_t *fl = (size_t *)malloc(n * sizeof(size_t));
	size_t i = 0, start = 0;
	for (size_t p = 0; p <= len; ++p)
		if (p == len || line[p] == '\t') {
			f[i] = line + start;
			fl[i++] = p - start;
			line[p] = 0;                                           /* line has len + 1 bytes */
			start = p + 1;
		}
	*n_out = n;
	*len_out = fl;
	return f;
}
static int field_is(char *const *f, const size_t *fl, size_t i, const char *lit)
{
	return fl[i] == strlen(lit) && memcmp(f[i], lit, fl[i]) == 0;
}

static void mc_count_window(void *c, uint64_t hv, uint64_t pos, uint64_t fw, uint64_t rv);
struct win_cb { ntsm_oracle_mc *mc; const uint8_t *geno; uint32_t S; unsigned multi; int var; FILE *warn; };
static void mc_count_window(void *c, uint64_t hv, uint64_t pos, uint64_t fw, uint64_t rv)
{   /* :148-169 */
	struct win_cb *w = (struct win_cb *)c;
	(void)pos; (void)fw; (void)rv;
	for (uint32_t i = 0; i < w->S; ++i) {
		const uint8_t g = w->geno[i];                              /* 0 hom1, 1 het, 2 hom2 */
		if (g == (w->var ? 2 : 0)) ntsm_oracle_mc_insert(w->mc, i, hv, w->multi * 2, w->warn);
		else if (g == 1) ntsm_oracle_mc_insert(w->mc, i, hv, w->multi, w->warn);
	}
}

int ntsm_oracle_vcf_convert(const char *sites_path, const char *ref_path, const char *vcf_path, unsigned k, int dupes,
                            unsigned multi, unsigned window, FILE *warn, ntsm_oracle_mc **mc_out, char ***sample_ids_out,
                            uint32_t *n_samples_out)
{
	*mc_out = NULL;
	/* the reference genome: every record kept whole, last record of a name wins (:47-58) */
	ntsm_oracle_reader *r = ntsm_oracle_reader_open(ref_path);
	if (!r) return -1;
	struct chr_rec *chr = NULL;
	size_t n_chr = 0;
	const char *seq, *name;
	long l;
	while ((l = ntsm_oracle_reader_next(r, &seq, &name)) >= 0) {
		chr = (struct chr_rec *)realloc(chr, (n_chr + 1) * sizeof *chr);
		chr[n_chr].name = strdup(name);
		chr[n_chr].seq = (char *)malloc((size_t)l + 1);
		memcpy(chr[n_chr].seq, seq, (size_t)l);
		chr[n_chr].seq[l] = 0;
		chr[n_chr].len = (size_t)l;
		++n_chr;
	}
	ntsm_oracle_reader_close(r);

	FILE *fh = fopen(vcf_path, "rb");
	int rc = fh ? 0 : -1;
	char *line = NULL;
	size_t cap = 0;
	ssize_t got;
	char **ids = NULL;
	uint32_t S = 0;
	/* header: lines are looked at until the one whose first field is "#CHROM" (:71-93) */
	while (rc == 0 && (got = getline(&line, &cap, fh)) >= 0) {
		if (got > 0 && line[got - 1] == '\n') line[--got] = 0;
		if (got == 0) { rc = -134; break; }                        /* line.at(0) throws */
		if (line[0] != '#') continue;
		size_t nf, *fl;
		char **f = split_tabs(line, (size_t)got, &nf, &fl);
		if (field_is(f, fl, 0, "#CHROM")) {
			/* `while (getline(ss, item, '\t'))` (:88): a tab at the very end of the line leaves nothing to extract and
			 * ends the loop -- no empty last ID (an empty field between two tabs does count) */
			if (nf > 9 && fl[nf - 1] == 0) --nf;
			for (size_t i = 9; i < nf; ++i) {                      /* 8 more fields skipped, the rest are sample IDs */
				ids = (char **)realloc(ids, (S + 1) * sizeof(char *));
				ids[S++] = strdup(f[i]);
			}
			free(f); free(fl);
			break;
		}
		free(f); free(fl);
	}
	ntsm_oracle_mc *mc = NULL;
	if (rc == 0) {
		mc = ntsm_oracle_mc_create(sites_path, k, dupes, S, warn);
		if (!mc) rc = -1;
	}
	uint8_t *geno = (uint8_t *)malloc(S + 1);
	char *wref = (char *)malloc(window + 1), *wvar = (char *)malloc(window + 1);
	while (rc == 0 && fh && (got = getline(&line, &cap, fh)) >= 0) {
		if (got == 0 || line[got - 1] != '\n') break;              /* :101-108: a last line without its newline leaves the stream at eof and is dropped */
		line[--got] = 0;
		size_t nf, *fl;
		char **f = split_tabs(line, (size_t)got, &nf, &fl);
		do {
			/* getline on an exhausted stringstream leaves `item` as it was (:110-125): a missing field
			 * reads as the last one present */
#define FIDX(i) ((size_t)(i) < nf ? (size_t)(i) : nf - 1)
#define FIELD(i) (f[FIDX(i)])
			char *end;
			errno = 0;
			const long loc = strtol(FIELD(1), &end, 10);           /* stoi (:113): leading integer, or it throws */
			if (end == FIELD(1) || errno == ERANGE || loc > 2147483647L || loc < -2147483647L - 1) { rc = -134; break; }
			if (field_is(f, fl, FIDX(3), ".")) break;              /* :121-123 */
			if (fl[FIDX(4)] != 1) break;                           /* :125-127: ALT must be one character; REF is not looked at */
			const char alt = FIELD(4)[0];
			/* getSeqFromSite (:202-215) */
			const struct chr_rec *c = NULL;
			for (size_t i = 0; i < n_chr; ++i)
				if (field_is(f, fl, 0, chr[i].name)) c = &chr[i];  /* later record of the same name wins (:52) */
			if (!c) { rc = -134; break; }
			const unsigned half = window / 2;
			if (loc < (long)half + 1 || (size_t)(loc - half - 1) > c->len) { rc = -2; break; }
			const size_t offset = (size_t)loc - half - 1;
			memset(wref, 0, window + 1);
			memset(wvar, 0, window + 1);
			size_t avail = c->len - offset < window ? c->len - offset : window;   /* strncpy stops at the sequence's NUL (or one inside it) and pads */
			avail = strnlen(c->seq + offset, avail);
			memcpy(wref, c->seq + offset, avail);
			memcpy(wvar, c->seq + offset, avail);
			wvar[half] = alt;                                      /* :211 */
			/* sample columns (:136-146) */
			size_t n_cols = nf > 9 ? nf - 9 : 0;
			if (n_cols && fl[nf - 1] == 0) --n_cols;               /* the same `while (getline(...))` (:136): a trailing tab adds no column */
			if (n_cols != (size_t)S) { rc = -134; break; }         /* assert(sampleIndex == m_sampleIDs.size()) (:146) */
			for (uint32_t i = 0; i < S; ++i) {
				geno[i] = field_is(f, fl, 9 + i, "0|0") ? 0 : (field_is(f, fl, 9 + i, "0|1") || field_is(f, fl, 9 + i, "1|0")) ? 1
				          : field_is(f, fl, 9 + i, "1|1") ? 2 : 0;   /* anything else keeps the vector's initial hom1 */
			}
			struct win_cb cb = { mc, geno, S, multi, 0, warn };
			iterate(wref, strlen(wref), k, mc_count_window, &cb);  /* :148-158 */
			cb.var = 1;
			iterate(wvar, strlen(wvar), k, mc_count_window, &cb);  /* :159-169 */
		} while (0);
		free(f); free(fl);
	}
	free(geno); free(wref); free(wvar); free(line);
	if (fh) fclose(fh);
	for (size_t i = 0; i < n_chr; ++i) { free(chr[i].name); free(chr[i].seq); }
	free(chr);
	if (rc != 0) {
		ntsm_oracle_mc_destroy(mc);
		for (uint32_t i = 0; i < S; ++i) free(ids[i]);
		free(ids);
		return rc;
	}
	*mc_out = mc;
	*sample_ids_out = ids;
	*n_samples_out = S;
	return 0;
}

/* Whole run with the file layout of tools/ref_vcf_harness.cpp: <prefix>_matrix.tsv, <prefix>_center.txt,
 * <prefix>_counts_<j>.txt per sample, <prefix>_mat.bin, and the warnings in <prefix>_stderr.txt. */
int ntsm_oracle_vcf_run(const char *sites_path, const char *ref_path, const char *vcf_path, unsigned k, int dupes,
                        unsigned multi, unsigned window, const char *prefix)
{
	char path[4096];
	snprintf(path, sizeof path, "%s_stderr.txt", prefix);
	FILE *warn = fopen(path, "wb");
	if (!warn) return -1;
	ntsm_oracle_mc *mc;
	char **ids;
	uint32_t S;
	int rc = ntsm_oracle_vcf_convert(sites_path, ref_path, vcf_path, k, dupes, multi, window, warn, &mc, &ids, &S);
	fclose(warn);
	if (rc) return rc;
	snprintf(path, sizeof path, "%s_mat.bin", prefix);
	FILE *f = fopen(path, "wb");
	uint64_t bytes;
	const uint8_t *m = ntsm_oracle_mc_matrix(mc, &bytes);
	fwrite(m, 1, bytes, f);
	fclose(f);
	for (uint32_t j = 0; j < S && rc == 0; ++j) {
		snprintf(path, sizeof path, "%s_counts_%u.txt", prefix, j);
		f = fopen(path, "wb");
		rc = ntsm_oracle_mc_print_counts_max(mc, j, f);
		fclose(f);
	}
	if (rc == 0) {
		snprintf(path, sizeof path, "%s_matrix.tsv", prefix);
		f = fopen(path, "wb");
		snprintf(path, sizeof path, "%s_center.txt", prefix);
		FILE *c = fopen(path, "wb");
		rc = ntsm_oracle_mc_print_norm_matrix(mc, (const char *const *)ids, f, c);
		fclose(f);
		fclose(c);
	}
	for (uint32_t j = 0; j < S; ++j) free(ids[j]);
	free(ids);
	ntsm_oracle_mc_destroy(mc);
	return rc;
}

/* ------------------------------------------------------------------ */
#ifdef NTSM_ORACLE_MAIN
/* Minimal driver with the reference's flags (-s -k -d -m, files...), single
 * thread: stdout is the counts file, stderr the warnings + summary. */
#include <unistd.h>
int main(int argc, char **argv)
{
	const char *sites = NULL;
	unsigned k = 19;
	int dupes = 0, c;
	double m = 0;
	while ((c = getopt(argc, argv, "s:k:dm:t:")) != -1) {
		if (c == 's') sites = optarg;
		else if (c == 'k') k = (unsigned)atoi(optarg);
		else if (c == 'd') dupes = 1;
		else if (c == 'm') m = atof(optarg);
	}
	if (!sites || optind >= argc) { fprintf(stderr, "usage: ntsm_oracle -s sites.fa [-k K] [-d] [-m COV] files...\n"); return 1; }
	ntsm_oracle_fp *fp = ntsm_oracle_fp_create(sites, k, dupes, m, stderr);
	if (!fp) { fprintf(stderr, "file %s cannot be opened\n", sites); return 1; }
	for (int i = optind; i < argc; ++i)
		if (ntsm_oracle_fp_count_file(fp, argv[i])) { fprintf(stderr, "file %s cannot be opened\n", argv[i]); return 1; }
	if (ntsm_oracle_fp_early_term(fp)) fprintf(stderr, "Reached desired (-m) threshold\n");
	int rc = ntsm_oracle_fp_print(fp, stdout);
	if (rc) { fflush(stdout); fprintf(stderr, "terminate: std::out_of_range (Couldn't find key.)\n"); return 134; }
	char buf[1024];
	ntsm_oracle_fp_summary(fp, buf, sizeof buf);
	fprintf(stderr, "%s\n", buf);
	ntsm_oracle_fp_destroy(fp);
	return 0;
}
#endif
