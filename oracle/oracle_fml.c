/*
 * oracle_fml.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, single-threaded restatement of fermi-lite's BFC stages as fml_assemble runs them: k-mer counting
 * (fml_count), per-read error correction (fml_correct -> bfc_ec1) and the unique-k-mer filter (fml_fltuniq).  It is
 * the checker of the CUDA path where the compiled reference (oracle/_ref) is not available, and an independent second
 * opinion where it is.  Pinned: tests/test_cpu_fml.py checks it against the committed golden vectors
 * (tests/golden/fml_*.npz, made by the reference's own fermi-lite) and, when oracle/_ref exists, against the live
 * reference on fresh inputs.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this; the
 * product never does.  The assembly half (FMD-index, unitigs, graph cleaning) has no restatement here: its oracle is the
 * reference compiled in place plus the committed stage dumps (BWT digest, rank answers, graph text, unitigs).
 *
 * Written independently of the device code: the count table is a sorted array searched by bisection (the reference
 * uses 2^l_pre khash tables, the device one open-addressing table); only the equivalence classes of keys and the
 * saturating counts are observable.  Every function names the reference code it restates (paths relative to the tree).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "../include/seqlib_b200.h"

typedef uint64_t u64;
typedef uint8_t u8;

/* ------------------------------------------------------------------ k-mers (fermi-lite/kmer.h) */
typedef struct { u64 x[4]; } kmer_t;

static const kmer_t kmer_null = {{0, 0, 0, 0}};

static int nt4(unsigned char c) /* seq_nt6_table[c] - 1 (fermi-lite/misc.c:12-29) */
{
	switch (c) {
	case 'A': case 'a': return 0;
	case 'C': case 'c': return 1;
	case 'G': case 'g': return 2;
	case 'T': case 't': return 3;
	default: return 4;
	}
}

static void kmer_append(int k, u64 x[4], int c) /* bfc_kmer_append (kmer.h:10-17) */
{
	u64 mask = (1ULL << k) - 1;
	x[0] = (x[0] << 1 | (u64)(c & 1)) & mask;
	x[1] = (x[1] << 1 | (u64)(c >> 1)) & mask;
	x[2] = x[2] >> 1 | (1ULL ^ (u64)(c & 1)) << (k - 1);
	x[3] = x[3] >> 1 | (1ULL ^ (u64)(c >> 1)) << (k - 1);
}

static void kmer_change(int k, u64 x[4], int d, int c) /* bfc_kmer_change (kmer.h:19-28) */
{
	u64 t = ~(1ULL << d);
	x[0] = (u64)(c & 1) << d | (x[0] & t);
	x[1] = (u64)(c >> 1) << d | (x[1] & t);
	t = ~(1ULL << (k - 1 - d));
	x[2] = (u64)(1 ^ (c & 1)) << (k - 1 - d) | (x[2] & t);
	x[3] = (u64)(1 ^ (c >> 1)) << (k - 1 - d) | (x[3] & t);
}

static u64 hash64(u64 key, u64 mask) /* bfc_hash_64 (kmer.h:31-41) */
{
	key = (~key + (key << 21)) & mask;
	key = key ^ key >> 24;
	key = ((key + (key << 3)) + (key << 8)) & mask;
	key = key ^ key >> 14;
	key = ((key + (key << 2)) + (key << 4)) & mask;
	key = key ^ key >> 28;
	key = (key + (key << 31)) & mask;
	return key;
}

/* bfc_kmer_hash (kmer.h:79-88) followed by get_subhash (htab.c:45-58): the (sub-table, stored key) pair */
typedef struct { u64 hi, lo; } ckey_t;

static ckey_t kmer_key(int k, int l_pre, const u64 x[4])
{
	int t = k >> 1, u = ((x[1] >> t & 1) > (x[3] >> t & 1));
	u64 mask = (1ULL << k) - 1, h0, h1, y0, y1;
	ckey_t r;
	h0 = hash64((x[u << 1 | 0] + x[u << 1 | 1]) & mask, mask);
	h1 = hash64(h0 ^ x[u << 1 | 1], mask);
	y0 = (h0 + h1) & mask; y1 = h1;
	if (k <= 32) {
		int tt = k * 2 - l_pre;
		u64 z = y0 << k | y1;
		r.lo = tt >= 64 ? z : (z & ((1ULL << tt) - 1));
		r.hi = tt >= 64 ? 0 : z >> tt;
	} else {
		int tt = k - l_pre;
		int shift = tt + k < 50 ? k : 50 - tt;
		r.lo = ((((y0 & ((1ULL << tt) - 1)) << shift) ^ y1) << 14) >> 14;
		r.hi = y0 >> tt;
	}
	return r;
}

/* ------------------------------------------------------------------ count table (fermi-lite/htab.c, bfc.c:66-99) */
typedef struct { ckey_t key; int cnt, high; } cent_t;
typedef struct { int k, l_pre; size_t n; cent_t *a; } ctab_t;

static int ckey_cmp(const ckey_t *a, const ckey_t *b)
{
	if (a->hi != b->hi) return a->hi < b->hi ? -1 : 1;
	if (a->lo != b->lo) return a->lo < b->lo ? -1 : 1;
	return 0;
}
static int cent_cmp(const void *a, const void *b) { return ckey_cmp(&((const cent_t*)a)->key, &((const cent_t*)b)->key); }

/* bfc_ch_init's clamping of l_pre (htab.c:20-33) */
static int clamp_l_pre(int k, int l_pre)
{
	if (k * 2 - l_pre > 50) l_pre = k * 2 - 50;
	if (l_pre > 20) l_pre = 20;
	if (l_pre < 0) l_pre = 0;
	return l_pre;
}

/* fml_count -> worker_count -> bfc_ch_insert: one entry per k-mer, merged with saturating counters */
static ctab_t *count_kmers(int64_t n, const char *seqs, const char *quals, const int64_t *off, int k, int q, int l_pre)
{
	ctab_t *t = (ctab_t*)calloc(1, sizeof(ctab_t));
	int64_t r, tot = n > 0 ? off[n] - off[0] : 0;
	size_t m = 0, i, j;
	t->k = k; t->l_pre = clamp_l_pre(k, l_pre);
	t->a = (cent_t*)malloc((size_t)(tot > 0 ? tot : 1) * sizeof(cent_t));
	for (r = 0; r < n; ++r) {
		const char *s = seqs + off[r], *ql = quals ? quals + off[r] : 0;
		int len = (int)(off[r + 1] - off[r]), l = 0, p;
		kmer_t x = kmer_null;
		u64 qmer = 0, mask = (1ULL << k) - 1;
		for (p = 0; p < len; ++p) {
			int c = nt4((unsigned char)s[p]);
			if (c < 4) {
				kmer_append(k, x.x, c);
				qmer = (qmer << 1 | (u64)(ql == 0 || ql[p] - 33 >= q)) & mask;
				if (++l >= k) {
					t->a[m].key = kmer_key(k, t->l_pre, x.x);
					t->a[m].cnt = 1; t->a[m].high = (qmer == mask);
					++m;
				}
			} else { l = 0; qmer = 0; x = kmer_null; }
		}
	}
	qsort(t->a, m, sizeof(cent_t), cent_cmp);
	for (i = j = 0; i < m; ++j) {      /* first insert: count 1 (+ high); later ones saturate at 255 / 63 (htab.c:74-80) */
		size_t e = i; int cnt = 0, high = 0;
		while (e < m && ckey_cmp(&t->a[e].key, &t->a[i].key) == 0) { ++cnt; high += t->a[e].high; ++e; }
		t->a[j].key = t->a[i].key;
		t->a[j].cnt = cnt > 255 ? 255 : cnt;
		t->a[j].high = high > 63 ? 63 : high;
		i = e;
	}
	t->n = j;
	return t;
}

static void ctab_destroy(ctab_t *t) { if (t) { free(t->a); free(t); } }

/* bfc_ch_kmer_occ (htab.c:85-93): -1 if absent, else high << 8 | total */
static int kmer_occ(const ctab_t *t, const kmer_t *z)
{
	ckey_t key = kmer_key(t->k, t->l_pre, z->x);
	size_t lo = 0, hi = t->n;
	while (lo < hi) {
		size_t mid = (lo + hi) >> 1;
		int c = ckey_cmp(&t->a[mid].key, &key);
		if (c == 0) return t->a[mid].high << 8 | t->a[mid].cnt;
		if (c < 0) lo = mid + 1; else hi = mid;
	}
	return -1;
}

/* bfc_ch_hist (htab.c:104-127) */
static int ctab_hist(const ctab_t *t, u64 cnt[256], u64 high[64])
{
	size_t i; int max_i = -1; u64 max = 0;
	memset(cnt, 0, 256 * 8); memset(high, 0, 64 * 8);
	for (i = 0; i < t->n; ++i) { ++cnt[t->a[i].cnt]; ++high[t->a[i].high]; }
	for (i = 3; i < 256; ++i) if (cnt[i] > max) { max = cnt[i]; max_i = (int)i; }
	return max_i;
}

/* ------------------------------------------------------------------ error correction (fermi-lite/bfc.c:101-466) */
typedef struct { int k, q, min_cov, max_end_ext, win_multi_ec, w_ec, w_ec_high, w_absent, w_absent_high, max_path_diff, max_heap; float min_trim_frac; } eopt_t;

static void eopt_init(eopt_t *o) /* bfc_opt_init (bfc.c:18-37) */
{
	o->q = 20; o->k = -1; o->min_cov = 4; o->win_multi_ec = 10; o->max_end_ext = 5; o->min_trim_frac = .8f;
	o->w_ec = 1; o->w_ec_high = 7; o->w_absent = 3; o->w_absent_high = 1; o->max_path_diff = 15; o->max_heap = 100;
}

typedef struct { int b, q, ob, oq, lcov, hcov, solid_end, high_end; } ebase_t;   /* ecbase_t (bfc.h:86-91); lcov/hcov are 6-bit */
typedef struct { int ec, ec_high, absent, absent_high, b; } epen_t;              /* bfc_penalty_t */
typedef struct { int tot_pen, i, k; int32_t ecpos_high[2], ecpos[5]; kmer_t x; } eheap_t;   /* echeap1_t */
typedef struct { int parent, i, tot_pen, b; epen_t pen; } estack_t;              /* ecstack1_t (cnt is never read back) */

typedef struct {
	const eopt_t *opt; const ctab_t *ch; int mode;
	eheap_t *heap; size_t n_heap, m_heap;
	estack_t *stack; size_t n_stack, m_stack;
	ebase_t *seq; int *ec[2]; int n, m;
} ebuf_t;

#define WPEN(o, p) ((o)->w_ec * (p).ec + (o)->w_ec_high * (p).ec_high + (o)->w_absent * (p).absent + (o)->w_absent_high * (p).absent_high)

static void heap_up(size_t n, eheap_t *l) /* ks_heapup with heap_lt(a,b) = a.tot_pen > b.tot_pen (ksort.h:125-136) */
{
	size_t i, k = n - 1;
	eheap_t tmp = l[k];
	while (k) {
		i = (k - 1) >> 1;
		if (tmp.tot_pen > l[i].tot_pen) break;
		l[k] = l[i]; k = i;
	}
	l[k] = tmp;
}

static void heap_down(size_t i, size_t n, eheap_t *l) /* ks_heapdown (ksort.h:137-146) */
{
	size_t k = i;
	eheap_t tmp = l[i];
	while ((k = (k << 1) + 1) < n) {
		if (k != n - 1 && l[k].tot_pen > l[k + 1].tot_pen) ++k;
		if (l[k].tot_pen > tmp.tot_pen) break;
		l[i] = l[k]; i = k;
	}
	l[i] = tmp;
}

static void buf_update(ebuf_t *e, const eheap_t *prev, epen_t pen) /* bfc.c:231-263 */
{
	estack_t *q; eheap_t *r;
	if (e->n_stack == e->m_stack) { e->m_stack = e->m_stack ? e->m_stack << 1 : 256; e->stack = (estack_t*)realloc(e->stack, e->m_stack * sizeof(estack_t)); }
	if (e->n_heap == e->m_heap) { e->m_heap = e->m_heap ? e->m_heap << 1 : 256; e->heap = (eheap_t*)realloc(e->heap, e->m_heap * sizeof(eheap_t)); }
	q = &e->stack[e->n_stack++];
	q->parent = prev->k; q->i = prev->i; q->b = pen.b; q->pen = pen;
	q->tot_pen = prev->tot_pen + WPEN(e->opt, pen);
	r = &e->heap[e->n_heap++];
	r->i = prev->i + 1; r->k = (int)e->n_stack - 1; r->x = prev->x;
	if (pen.ec_high) { r->ecpos_high[1] = prev->ecpos_high[0]; r->ecpos_high[0] = prev->i; }
	else memcpy(r->ecpos_high, prev->ecpos_high, sizeof(r->ecpos_high));
	if (pen.ec) { memmove(r->ecpos + 1, prev->ecpos, 4 * sizeof(int32_t)); r->ecpos[0] = prev->i; }
	else memcpy(r->ecpos, prev->ecpos, sizeof(r->ecpos));
	r->tot_pen = q->tot_pen;
	kmer_append(e->opt->k, r->x.x, pen.b);
	heap_up(e->n_heap, e->heap);
}

/* bfc_ec1dir (bfc.c:280-399); ec[i] receives the corrected base, 4 outside the corrected range */
static int ec1dir(ebuf_t *e, int *ec, int start, int end)
{
	const eopt_t *o = e->opt;
	const ebase_t *seq = e->seq;
	int n = e->n, i, l, rv = -1, path[4], n_paths = 0, min_path = -1, min_path_pen = 0x7fffffff, n_failures = 0;
	eheap_t z;
	e->n_heap = e->n_stack = 0;
	memset(&z, 0, sizeof(z));
	for (z.i = start, l = 0; z.i < end; ++z.i) {
		int c = seq[z.i].b;
		if (c < 4) {
			if (++l == o->k) break;
			kmer_append(o->k, z.x.x, c);
		} else { l = 0; z.x = kmer_null; }
	}
	z.k = -1;
	for (i = 0; i < 5; ++i) z.ecpos[i] = -1;
	for (i = 0; i < 2; ++i) z.ecpos_high[i] = -1;
	if (e->m_heap == 0) { e->m_heap = 256; e->heap = (eheap_t*)malloc(256 * sizeof(eheap_t)); }
	e->heap[e->n_heap++] = z;
	for (i = 0; i < n; ++i) ec[i] = seq[i].b;
	for (;;) {
		int stop = 0;
		if (e->n_heap == 0) { rv = -2; break; }
		z = e->heap[0];
		e->heap[0] = e->heap[--e->n_heap];
		heap_down(0, e->n_heap, e->heap);
		if (min_path >= 0 && z.tot_pen > min_path_pen + o->max_path_diff) break;
		if (z.i - end > o->max_end_ext) stop = 1;
		if (!stop) {
			const ebase_t *c = z.i < n ? &seq[z.i] : 0;
			int b, os = -1, fixed = 0, other_ext = 0, n_added = 0;
			epen_t added[4];
			if (z.i > end) fixed = 1;
			if (c && c->b < 4) {
				kmer_t x = z.x;
				kmer_append(o->k, x.x, c->b);
				os = kmer_occ(e->ch, &x);
				if (c->q && (os & 0xff) >= o->min_cov + 1 && c->lcov >= o->min_cov + 1) fixed = 1;
				else if (c->hcov > o->k * .75) fixed = 1;
			}
			for (b = 0; b < 4; ++b) {
				epen_t pen;
				if (fixed && c && b != c->b) continue;
				if (c == 0 || b != c->b) {
					int s;
					kmer_t x = z.x;
					if (c) {
						if (c->q && z.ecpos_high[1] >= 0 && z.i - z.ecpos_high[1] < o->win_multi_ec) continue;
						if (z.ecpos[4] >= 0 && z.i - z.ecpos[4] < o->win_multi_ec) continue;
					}
					kmer_append(o->k, x.x, b);
					s = kmer_occ(e->ch, &x);
					if (s < 0 || (s & 0xff) < o->min_cov) continue;
					pen.ec = c && c->b < 4 ? 1 : 0;
					pen.ec_high = pen.ec ? c->oq : 0;
					pen.absent = 0;
					pen.absent_high = ((s >> 8 & 0xff) < o->min_cov);
					pen.b = b;
					added[n_added++] = pen;
					++other_ext;
				} else {
					pen.ec = pen.ec_high = 0;
					pen.absent = (os < 0 || (os & 0xff) < o->min_cov);
					pen.absent_high = (os < 0 || (os >> 8 & 0xff) < o->min_cov);
					pen.b = b;
					added[n_added++] = pen;
				}
			}
			if (fixed == 0 && other_ext == 0) ++n_failures;
			if (n_failures > n * 2) { rv = -3; break; }
			if (c || n_added == 1) {
				if (n_added > 1 && (int)e->n_heap > o->max_heap) {
					int min_b = -1, min = 0x7fffffff;
					for (b = 0; b < n_added; ++b) { int t = WPEN(o, added[b]); if (min > t) { min = t; min_b = b; } }
					buf_update(e, &z, added[min_b]);
				} else for (b = 0; b < n_added; ++b) buf_update(e, &z, added[b]);
			} else {
				if (n_added == 0) e->stack[z.k].tot_pen += o->w_absent * (o->max_end_ext - (z.i - end));
				stop = 1;
			}
		}
		if (stop) {
			if (e->stack[z.k].tot_pen < min_path_pen) { min_path_pen = e->stack[z.k].tot_pen; min_path = n_paths; }
			path[n_paths++] = z.k;
			if (n_paths == 4) break;
		}
	}
	if (n_paths == 0) return rv;
	{	/* buf_backtrack (bfc.c:265-278) */
		int endp = path[min_path];
		rv = 0;
		while (endp >= 0) {
			if ((i = e->stack[endp].i) < n) { ec[i] = e->stack[endp].b; rv += e->stack[endp].pen.absent; }
			endp = e->stack[endp].parent;
		}
	}
	for (i = 0; i < n; ++i) if (i < start + o->k || i >= end) ec[i] = 4;
	return rv;
}

static void seq_revcomp(ebase_t *a, int n) /* bfc_seq_revcomp (bfc.c:126-137) */
{
	int i;
	for (i = 0; i < n >> 1; ++i) {
		ebase_t t = a[i]; a[i] = a[n - 1 - i]; a[n - 1 - i] = t;
	}
	for (i = 0; i < n; ++i) { a[i].b = a[i].b < 4 ? 3 - a[i].b : 4; a[i].ob = a[i].ob < 4 ? 3 - a[i].ob : 4; }
}

/* bfc_ec1 (bfc.c:401-466): corrects seq / qual (len bytes) in place; returns the ec_code */
static int ec1(ebuf_t *e, char *seq, char *qual, int len)
{
	const eopt_t *o = e->opt;
	int i, l, start = 0, end = 0, n_n = 0, rv0, rv1, k = o->k;
	ebase_t *a;
	kmer_t x;
	if (len > e->m) {
		e->m = len + 64;
		e->seq = (ebase_t*)realloc(e->seq, e->m * sizeof(ebase_t));
		e->ec[0] = (int*)realloc(e->ec[0], e->m * sizeof(int)); e->ec[1] = (int*)realloc(e->ec[1], e->m * sizeof(int));
	}
	e->n = len; a = e->seq;
	for (i = 0; i < len; ++i) {          /* bfc_seq_conv (bfc.c:101-116) */
		ebase_t *c = &a[i];
		c->b = c->ob = nt4((unsigned char)seq[i]);
		c->q = c->oq = !qual ? 1 : (qual[i] - 33 >= o->q ? 1 : 0);
		if (c->b > 3) c->q = c->oq = 0;
		c->lcov = c->hcov = c->solid_end = c->high_end = 0;
		if (c->ob > 3) ++n_n;
	}
	if (n_n > len * .05) return 2;       /* ECCODE_MANY_N */
	x = kmer_null;                       /* bfc_ec_kcov (bfc.c:175-196) */
	for (i = l = 0; i < len; ++i) {
		ebase_t *c = &a[i];
		c->high_end = c->solid_end = c->lcov = c->hcov = 0;
		if (c->b < 4) {
			kmer_append(k, x.x, c->b);
			if (++l >= k) {
				int r = kmer_occ(e->ch, &x), j;
				if (r >= 0) {
					if ((r >> 8 & 0x3f) >= o->min_cov + 1) c->high_end = 1;
					if ((r & 0xff) >= o->min_cov) {
						c->solid_end = 1;
						for (j = i - k + 1; j <= i; ++j) { a[j].lcov = (a[j].lcov + 1) & 63; a[j].hcov = (a[j].hcov + c->high_end) & 63; }
					}
				}
			}
		} else { l = 0; x = kmer_null; }
	}
	{	/* bfc_ec_best_island (bfc.c:198-209) */
		int max = 0, max_i = -1;
		for (i = k - 1, l = 0; i < len; ++i) {
			if (!a[i].solid_end) { if (l > max) { max = l; max_i = i; } l = 0; }
			else ++l;
		}
		if (l > max) { max = l; max_i = i; }
		if (max > 0) { start = max_i - max - k + 1; end = max_i; }
		else {   /* no solid k-mer: bfc_ec_first_kmer + bfc_ec_greedy_k (bfc.c:142-173,417-432) */
			int ec = -1;
			for (;;) {
				int ll = 0;
				x = kmer_null;
				for (end = start; end < len; ++end) {
					if (a[end].b < 4) { kmer_append(k, x.x, a[end].b); if (++ll == k) break; }
					else { ll = 0; x = kmer_null; }
				}
				if (end >= len) break;
				{
					int p, j, mx = 0, mx2 = 0, mx_ec = -1;
					for (p = 0; p < k; ++p) {
						int c0 = (int)(x.x[1] >> p & 1) << 1 | (int)(x.x[0] >> p & 1);
						for (j = 0; j < 4; ++j) {
							kmer_t y = x; int ret;
							if (j == c0) continue;
							kmer_change(k, y.x, p, j);
							ret = kmer_occ(e->ch, &y);
							if (ret < 0) continue;
							if ((mx & 0xff) < (ret & 0xff)) { mx2 = mx; mx = ret; mx_ec = p << 2 | j; }
							else if ((mx2 & 0xff) < (ret & 0xff)) mx2 = ret;
						}
					}
					ec = (mx & 0xff) * 3 > e->mode && (mx2 & 0xff) < 3 ? mx_ec : -1;
				}
				if (ec >= 0) break;
				if (end + (k >> 1) >= len) break;
				start = end - (k >> 1);
			}
			if (ec < 0) return 3;            /* ECCODE_NO_SOLID */
			a[end - (ec >> 2)].b = ec & 3;
			++end; start = end - k;
		}
	}
	if ((rv0 = ec1dir(e, e->ec[0], start, len)) < 0) return rv0 == -2 ? 4 : rv0 == -3 ? 5 : 1;
	seq_revcomp(a, len);
	if ((rv1 = ec1dir(e, e->ec[1], len - end, len)) < 0) return rv1 == -2 ? 4 : rv1 == -3 ? 5 : 1;
	for (i = 0; i < len >> 1; ++i) { int t = e->ec[1][i]; e->ec[1][i] = e->ec[1][len - 1 - i]; e->ec[1][len - 1 - i] = t; }
	for (i = 0; i < len; ++i) e->ec[1][i] = e->ec[1][i] < 4 ? 3 - e->ec[1][i] : 4;
	seq_revcomp(a, len);
	for (i = 0; i < len; ++i) {
		int e0 = e->ec[0][i], e1 = e->ec[1][i];
		if (e0 == e1) a[i].b = e0 > 3 ? a[i].b : e0;
		else if (e1 > 3) a[i].b = e0;
		else if (e0 > 3) a[i].b = e1;
		else a[i].b = a[i].ob;
	}
	for (i = 0; i < len; ++i) {
		int is_diff = !(a[i].b == a[i].ob);
		seq[i] = (is_diff ? "acgtn" : "ACGTN")[a[i].b];
		if (qual) qual[i] = is_diff ? (char)(34 + a[i].ob) : "+?"[a[i].q];
	}
	return 0;
}

/* max_streak + the flt_uniq branch of worker_ec (bfc.c:469-506); returns the new length (0 = dropped) */
static int fltuniq1(const eopt_t *o, const ctab_t *ch, char *seq, char *qual, int len)
{
	int i, l;
	u64 max = 0, t = 0;
	kmer_t x = kmer_null;
	for (i = l = 0; i < len; ++i) {
		int c = nt4((unsigned char)seq[i]);
		if (c < 4) {
			kmer_append(o->k, x.x, c);
			if (++l >= o->k) { if (kmer_occ(ch, &x) > 0) t += 1ULL << 32; else t = i + 1; }
			else t = i + 1;
		} else { l = 0; x = kmer_null; t = i + 1; }
		max = max > t ? max : t;
	}
	if (max >> 32 && (double)((max >> 32) + o->k - 1) / len > o->min_trim_frac) {
		int start = (int)(uint32_t)max, end = start + (int)(max >> 32);
		start -= o->k - 1;
		memmove(seq, seq + start, end - start);
		if (qual) memmove(qual, qual + start, end - start);
		return end - start;
	}
	return 0;
}

/* ------------------------------------------------------------------ entry points */

/* fml_correct_core (bfc.c:513-553) on flat pools: same contract as b200_fml_correct_flat */
int oracle_fml_correct_flat(const b200_fml_opt_t *opt, int flt_uniq, int64_t n, char *seqs, char *quals, const int64_t *off,
                            int32_t *len_out, float *kcov_out, uint64_t *hist_out /* 320 entries, optional */)
{
	eopt_t bo;
	ctab_t *ch;
	u64 hist[256], high[64], tot_len = n > 0 ? (u64)off[n] : 0, sum_k = 0, tot_k = 0;
	ebuf_t e;
	int64_t r;
	int i, mode;
	float kcov;
	eopt_init(&bo);
	bo.k = flt_uniq ? opt->min_asm_ovlp : opt->ec_k;
	if (bo.k <= 0) {                     /* the reference's k = 0 run changes nothing and reports 255 */
		if (kcov_out) *kcov_out = 255.0f;
		if (len_out) for (r = 0; r < n; ++r) len_out[r] = (int32_t)(off[r + 1] - off[r]);
		return 0;
	}
	ch = count_kmers(n, seqs, quals, off, bo.k, bo.q, tot_len - 8 < 20 ? (int)(tot_len - 8) : 20);
	mode = ctab_hist(ch, hist, high);
	if (hist_out) { memcpy(hist_out, hist, 256 * 8); memcpy(hist_out + 256, high, 64 * 8); }
	for (i = opt->min_cnt < 0 ? 0 : opt->min_cnt; i < 256; ++i) { sum_k += hist[i]; tot_k += (u64)i * hist[i]; }
	kcov = (float)tot_k / sum_k;
	bo.min_cov = sum_k ? (int)(.1 * kcov + .499) : opt->min_cnt;   /* (int)NaN is INT_MIN on x86-64; the clamps give min_cnt */
	bo.min_cov = bo.min_cov < opt->max_cnt ? bo.min_cov : opt->max_cnt;
	bo.min_cov = bo.min_cov > opt->min_cnt ? bo.min_cov : opt->min_cnt;
	if (kcov_out) *kcov_out = kcov;
	memset(&e, 0, sizeof(e));
	e.opt = &bo; e.ch = ch; e.mode = mode;
	for (r = 0; r < n; ++r) {
		int len = (int)(off[r + 1] - off[r]);
		char *s = seqs + off[r], *q = quals ? quals + off[r] : 0;
		if (flt_uniq) len_out[r] = len > 0 ? fltuniq1(&bo, ch, s, q, len) : 0;
		else { ec1(&e, s, q, len); if (len_out) len_out[r] = len; }
	}
	free(e.heap); free(e.stack); free(e.seq); free(e.ec[0]); free(e.ec[1]);
	ctab_destroy(ch);
	return 0;
}
