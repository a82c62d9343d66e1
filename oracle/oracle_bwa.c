/*
 * oracle_bwa.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, single-threaded restatement of the reference's per-read seed-and-extend path, working directly on
 * bwa's own index layout (the arrays of b200_index_view_t = bwt_t/bntseq_t/pac).  It is the checker of the CUDA
 * path where the compiled reference (oracle/_ref) is not available, and an independent second opinion where it
 * is.  Pinned: tests/test_cpu_suite.py checks it against the committed golden vectors (the reference's KAT,
 * tiny.fa reads, config-1 reads) and, when oracle/_ref exists, against the live reference on random inputs.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this; the product never does.
 *
 * Every function names the reference code it restates (paths relative to the SeqLib tree).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "../include/seqlib_b200.h"

typedef int64_t i64;
typedef uint64_t u64;
typedef uint32_t u32;
typedef uint8_t u8;

typedef struct {
	const b200_index_view_t *v;
	const b200_mem_opt_t *o;
} ctx_t;

/* ------------------------------------------------------------------ FM-index (bwa/bwt.c) */

static inline int pop2(u32 w, int c, int n) /* symbols equal to c among the first n (0..16) of a 16-symbol word, MSB first */
{
	int k, r = 0;
	for (k = 0; k < n; ++k) r += ((w >> ((15 - k) << 1)) & 3) == (u32)c;
	return r;
}

/* bwt_occ4 (bwa/bwt.c:169-186): counts of A,C,G,T in B0[0..k'] */
static void occ4(const b200_index_view_t *v, u64 k, u64 cnt[4])
{
	int c, w, nw, rem;
	const u32 *p;
	if (k == (u64)-1) { cnt[0] = cnt[1] = cnt[2] = cnt[3] = 0; return; }
	k -= (k >= v->primary);
	p = v->bwt + ((k >> 7) << 4);
	for (c = 0; c < 4; ++c) cnt[c] = ((const u64*)p)[c];
	p += 8;
	rem = (int)(k & 127) + 1;          /* symbols of this block to count */
	nw = rem >> 4;
	for (w = 0; w < nw; ++w) for (c = 0; c < 4; ++c) cnt[c] += pop2(p[w], c, 16);
	if (rem & 15) for (c = 0; c < 4; ++c) cnt[c] += pop2(p[nw], c, rem & 15);
}

typedef struct { u64 x[3], info; } intv_t;     /* bwtintv_t (bwa/bwt.h:62-64) */

/* bwt_extend (bwa/bwt.c:262-275) */
static void bwt_ext(const b200_index_view_t *v, const intv_t *ik, intv_t ok[4], int is_back)
{
	u64 tk[4], tl[4];
	int i;
	occ4(v, ik->x[!is_back] - 1, tk);
	occ4(v, ik->x[!is_back] - 1 + ik->x[2], tl);
	for (i = 0; i != 4; ++i) {
		ok[i].x[!is_back] = v->L2[i] + 1 + tk[i];
		ok[i].x[2] = tl[i] - tk[i];
	}
	ok[3].x[is_back] = ik->x[is_back] + (ik->x[!is_back] <= v->primary && ik->x[!is_back] + ik->x[2] - 1 >= v->primary);
	ok[2].x[is_back] = ok[3].x[is_back] + ok[3].x[2];
	ok[1].x[is_back] = ok[2].x[is_back] + ok[2].x[2];
	ok[0].x[is_back] = ok[1].x[is_back] + ok[1].x[2];
}

static void set_intv(const b200_index_view_t *v, int c, intv_t *ik) /* bwt_set_intv (bwa/bwt.h:82) */
{
	ik->x[0] = v->L2[c] + 1; ik->x[2] = v->L2[c+1] - v->L2[c]; ik->x[1] = v->L2[3-c] + 1; ik->info = 0;
}

typedef struct { size_t n, m; intv_t *a; } intv_v;
static void iv_push(intv_v *v, const intv_t *x)
{
	if (v->n == v->m) { v->m = v->m ? v->m << 1 : 16; v->a = realloc(v->a, v->m * sizeof(intv_t)); }
	v->a[v->n++] = *x;
}
static void iv_reverse(intv_t *a, size_t n) { size_t j; for (j = 0; j < n >> 1; ++j) { intv_t t = a[j]; a[j] = a[n-1-j]; a[n-1-j] = t; } }

/* bwt_smem1a with max_intv = 0 (bwa/bwt.c:289-351) */
static int smem1(const b200_index_view_t *v, int len, const u8 *q, int x, int min_intv, intv_v *mem, intv_v *t0, intv_v *t1)
{
	int i, j, c, ret;
	intv_t ik, ok[4];
	intv_v *prev = t0, *curr = t1, *swap;
	mem->n = 0;
	if (q[x] > 3) return x + 1;
	if (min_intv < 1) min_intv = 1;
	set_intv(v, q[x], &ik);
	ik.info = x + 1;
	for (i = x + 1, curr->n = 0; i < len; ++i) {
		if (q[i] < 4) {
			c = 3 - q[i];
			bwt_ext(v, &ik, ok, 0);
			if (ok[c].x[2] != ik.x[2]) {
				iv_push(curr, &ik);
				if (ok[c].x[2] < (u64)min_intv) break;
			}
			ik = ok[c]; ik.info = i + 1;
		} else { iv_push(curr, &ik); break; }
	}
	if (i == len) iv_push(curr, &ik);
	iv_reverse(curr->a, curr->n);
	ret = (int)curr->a[0].info;
	swap = curr; curr = prev; prev = swap;
	for (i = x - 1; i >= -1; --i) {
		c = i < 0? -1 : q[i] < 4? q[i] : -1;
		for (j = 0, curr->n = 0; j < (int)prev->n; ++j) {
			intv_t *p = &prev->a[j];
			if (c >= 0) bwt_ext(v, p, ok, 1);
			if (c < 0 || ok[c].x[2] < (u64)min_intv) {
				if (curr->n == 0) {
					if (mem->n == 0 || (u64)(i + 1) < mem->a[mem->n-1].info>>32) {
						ik = *p; ik.info |= (u64)(i + 1)<<32;
						iv_push(mem, &ik);
					}
				}
			} else if (curr->n == 0 || ok[c].x[2] != curr->a[curr->n-1].x[2]) {
				ok[c].info = p->info;
				iv_push(curr, &ok[c]);
			}
		}
		if (curr->n == 0) break;
		swap = curr; curr = prev; prev = swap;
	}
	iv_reverse(mem->a, mem->n);
	return ret;
}

/* bwt_seed_strategy1 (bwa/bwt.c:358-379) */
static int seed_strategy1(const b200_index_view_t *v, int len, const u8 *q, int x, int min_len, int max_intv, intv_t *mem)
{
	int i, c;
	intv_t ik, ok[4];
	memset(mem, 0, sizeof(intv_t));
	if (q[x] > 3) return x + 1;
	set_intv(v, q[x], &ik);
	for (i = x + 1; i < len; ++i) {
		if (q[i] < 4) {
			c = 3 - q[i];
			bwt_ext(v, &ik, ok, 0);
			if (ok[c].x[2] < (u64)max_intv && i - x >= min_len) {
				*mem = ok[c];
				mem->info = (u64)x<<32 | (u64)(i + 1);
				return i + 1;
			}
			ik = ok[c];
		} else return i + 1;
	}
	return len;
}

/* ------------------------------------------------------------------ klib introsort (bwa/ksort.h:137-226), generic over element size */

typedef int (*lt_f)(const void *a, const void *b);
static void swp(char *a, char *b, size_t sz) { char t[160]; memcpy(t, a, sz); memcpy(a, b, sz); memcpy(b, t, sz); }

static void insertsort(char *s, char *t, size_t sz, lt_f lt)
{
	char *i, *j;
	for (i = s + sz; i < t; i += sz)
		for (j = i; j > s && lt(j, j - sz); j -= sz) swp(j, j - sz, sz);
}

static void combsort(size_t n, char *a, size_t sz, lt_f lt)
{
	const double shrink_factor = 1.2473309501039786540366528676643;
	int do_swap;
	size_t gap = n;
	char *i, *j;
	do {
		if (gap > 2) {
			gap = (size_t)(gap / shrink_factor);
			if (gap == 9 || gap == 10) gap = 11;
		}
		do_swap = 0;
		for (i = a; i < a + (n - gap) * sz; i += sz) {
			j = i + gap * sz;
			if (lt(j, i)) { swp(i, j, sz); do_swap = 1; }
		}
	} while (do_swap || gap > 2);
	if (gap != 1) insertsort(a, a + n * sz, sz, lt);
}

static void introsort(size_t n, void *base, size_t sz, lt_f lt)
{
	typedef struct { char *left, *right; int depth; } frame_t;
	frame_t stack[130], *top = stack;
	char *a = base, *s, *t, *i, *j, *k, rp[160];
	int d;
	if (n < 1) return;
	if (n == 2) { if (lt(a + sz, a)) swp(a, a + sz, sz); return; }
	for (d = 2; 1ul<<d < n; ++d);
	s = a; t = a + (n - 1) * sz; d <<= 1;
	for (;;) {
		if (s < t) {
			if (--d == 0) { combsort((size_t)(t - s) / sz + 1, s, sz, lt); t = s; continue; }
			i = s; j = t; k = i + (((size_t)(j - i) / sz) >> 1) * sz + sz;
			if (lt(k, i)) { if (lt(k, j)) k = j; }
			else k = lt(j, i)? i : j;
			memcpy(rp, k, sz);
			if (k != t) swp(k, t, sz);
			for (;;) {
				do i += sz; while (lt(i, rp));
				do j -= sz; while (i <= j && lt(rp, j));
				if (j <= i) break;
				swp(i, j, sz);
			}
			swp(i, t, sz);
			if (i - s > t - i) {
				if ((size_t)(i - s) > 16 * sz) { top->left = s; top->right = i - sz; top->depth = d; ++top; }
				s = (size_t)(t - i) > 16 * sz? i + sz : t;
			} else {
				if ((size_t)(t - i) > 16 * sz) { top->left = i + sz; top->right = t; top->depth = d; ++top; }
				t = (size_t)(i - s) > 16 * sz? i - sz : s;
			}
		} else {
			if (top == stack) { insertsort(a, a + n * sz, sz, lt); return; }
			--top; s = top->left; t = top->right; d = top->depth;
		}
	}
}

static int intv_lt(const void *a, const void *b) { return ((const intv_t*)a)->info < ((const intv_t*)b)->info; }
static int u64_lt(const void *a, const void *b) { return *(const u64*)a < *(const u64*)b; }

/* mem_collect_intv (bwa/bwamem.c:140-188) */
static void collect_intv(const ctx_t *cx, int len, const u8 *seq, intv_v *mem)
{
	const b200_mem_opt_t *opt = cx->o;
	intv_v mem1 = {0,0,0}, t0 = {0,0,0}, t1 = {0,0,0};
	int i, k, x = 0, old_n;
	int split_len = (int)(opt->min_seed_len * opt->split_factor + .499);
	mem->n = 0;
	while (x < len) {
		if (seq[x] < 4) {
			x = smem1(cx->v, len, seq, x, 1, &mem1, &t0, &t1);
			for (i = 0; i < (int)mem1.n; ++i) {
				intv_t *p = &mem1.a[i];
				int slen = (u32)p->info - (p->info>>32);
				if (slen >= opt->min_seed_len) iv_push(mem, p);
			}
		} else ++x;
	}
	old_n = mem->n;
	for (k = 0; k < old_n; ++k) {
		intv_t p = mem->a[k];
		int start = p.info>>32, end = (int32_t)p.info;
		if (end - start < split_len || p.x[2] > (u64)opt->split_width) continue;
		smem1(cx->v, len, seq, (start + end)>>1, p.x[2]+1, &mem1, &t0, &t1);
		for (i = 0; i < (int)mem1.n; ++i)
			if ((u32)mem1.a[i].info - (mem1.a[i].info>>32) >= (u32)opt->min_seed_len) iv_push(mem, &mem1.a[i]);
	}
	if (opt->max_mem_intv > 0) {
		x = 0;
		while (x < len) {
			if (seq[x] < 4) {
				intv_t m;
				x = seed_strategy1(cx->v, len, seq, x, opt->min_seed_len, opt->max_mem_intv, &m);
				if (m.x[2] > 0) iv_push(mem, &m);
			} else ++x;
		}
	}
	introsort(mem->n, mem->a, sizeof(intv_t), intv_lt);
	free(mem1.a); free(t0.a); free(t1.a);
}

/* ------------------------------------------------------------------ suffix array and reference text */

static inline int bwt_B0(const b200_index_view_t *v, u64 k) { return v->bwt[((k>>7)<<4) + 8 + ((k&0x7f)>>4)] >> ((~k&0xf)<<1) & 3; } /* bwa/bwt.h:74-80 */

static u64 inv_psi(const b200_index_view_t *v, u64 k) /* bwt_invPsi (bwa/bwt.c:53-59) via bwt_occ (:107-129) */
{
	u64 x = k - (k > v->primary), cnt[4];
	int c = bwt_B0(v, x);
	if (k == v->primary) return 0;
	if (k == v->seq_len) return v->L2[c] + (v->L2[c+1] - v->L2[c]);
	occ4(v, k, cnt);
	return v->L2[c] + cnt[c];
}

static u64 bwt_sa(const b200_index_view_t *v, u64 k) /* bwa/bwt.c:86-96 */
{
	u64 sa = 0, mask = v->sa_intv - 1;
	while (k & mask) { ++sa; k = inv_psi(v, k); }
	return sa + v->sa[k / v->sa_intv];
}

static inline int get_pac(const u8 *pac, i64 l) { return pac[l>>2] >> ((~l&3)<<1) & 3; }

/* bns_get_seq (bwa/bntseq.c:403-424); returns malloc'd bases or NULL */
static u8 *get_seq(const b200_index_view_t *v, i64 beg, i64 end, i64 *len)
{
	i64 l_pac = v->l_pac, k, l = 0;
	u8 *seq = 0;
	if (end < beg) { i64 t = beg; beg = end; end = t; }
	if (end > l_pac<<1) end = l_pac<<1;
	if (beg < 0) beg = 0;
	if (beg >= l_pac || end <= l_pac) {
		*len = end - beg;
		seq = malloc(end - beg + 1);
		if (beg >= l_pac) {
			i64 beg_f = (l_pac<<1) - 1 - end, end_f = (l_pac<<1) - 1 - beg;
			for (k = end_f; k > beg_f; --k) seq[l++] = 3 - get_pac(v->pac, k);
		} else for (k = beg; k < end; ++k) seq[l++] = get_pac(v->pac, k);
	} else *len = 0;
	return seq;
}

static int pos2rid(const b200_index_view_t *v, i64 pos_f) /* bns_pos2rid (bwa/bntseq.c:354-368) */
{
	int left = 0, mid = 0, right = v->n_seqs;
	if (pos_f >= v->l_pac) return -1;
	while (left < right) {
		mid = (left + right) >> 1;
		if (pos_f >= v->contigs[mid].offset) {
			if (mid == v->n_seqs - 1) break;
			if (pos_f < v->contigs[mid+1].offset) break;
			left = mid + 1;
		} else right = mid;
	}
	return mid;
}
static i64 depos(const b200_index_view_t *v, i64 pos, int *is_rev) { return (*is_rev = (pos >= v->l_pac))? (v->l_pac<<1) - 1 - pos : pos; } /* bwa/bntseq.h:87-90 */
static int intv2rid(const b200_index_view_t *v, i64 rb, i64 re) /* bwa/bntseq.c:370-378 */
{
	int is_rev, rid_b, rid_e;
	if (rb < v->l_pac && re > v->l_pac) return -2;
	rid_b = pos2rid(v, depos(v, rb, &is_rev));
	rid_e = rb < re? pos2rid(v, depos(v, re - 1, &is_rev)) : rid_b;
	return rid_b == rid_e? rid_b : -1;
}

/* ------------------------------------------------------------------ chaining (bwa/bwamem.c:194-411) */

typedef struct { i64 rbeg; int32_t qbeg, len; int score; } seed_t;
typedef struct { int n, m, first, rid; u32 w, kept, is_alt; float frac_rep; i64 pos; seed_t *seeds; } chain_t;

/* The reference keeps chains in a klib B-tree (bwa/kbtree.h, node size 512 => t = 5).  With duplicate `pos` keys the chain
 * found by kb_intervalp and the in-order traversal depend on the tree shape, so the tree is restated node for node. */
#define BT 5
typedef struct btnode { int n, internal; chain_t key[2*BT-1]; struct btnode *ptr[2*BT]; } btnode_t;

static int bt_find(const btnode_t *x, i64 k, int *r) /* __kb_getp_aux (bwa/kbtree.h:123-138) */
{
	int begin = 0, end = x->n;
	if (x->n == 0) return -1;
	while (begin < end) {
		int mid = (begin + end) >> 1;
		if (x->key[mid].pos < k) begin = mid + 1; else end = mid;
	}
	if (begin == x->n) { *r = 1; return x->n - 1; }
	*r = (x->key[begin].pos < k) - (k < x->key[begin].pos);
	if (*r < 0) --begin;
	return begin;
}
static chain_t *bt_lower(btnode_t *root, i64 k) /* kb_intervalp (bwa/kbtree.h:159-178), lower bound only */
{
	chain_t *lower = 0; int r = 0; btnode_t *x = root;
	while (x) {
		int i = bt_find(x, k, &r);
		if (i >= 0 && r == 0) return &x->key[i];
		if (i >= 0) lower = &x->key[i];
		if (!x->internal) return lower;
		x = x->ptr[i + 1];
	}
	return lower;
}
static void bt_split(btnode_t *x, int i, btnode_t *y) /* __kb_split (bwa/kbtree.h:187-204) */
{
	btnode_t *z = calloc(1, sizeof(btnode_t));
	z->internal = y->internal; z->n = BT - 1;
	memcpy(z->key, y->key + BT, sizeof(chain_t) * (BT - 1));
	if (y->internal) memcpy(z->ptr, y->ptr + BT, sizeof(void*) * BT);
	y->n = BT - 1;
	memmove(x->ptr + i + 2, x->ptr + i + 1, sizeof(void*) * (x->n - i));
	x->ptr[i + 1] = z;
	memmove(x->key + i + 1, x->key + i, sizeof(chain_t) * (x->n - i));
	x->key[i] = y->key[BT - 1];
	++x->n;
}
static void bt_put_aux(btnode_t *x, const chain_t *k) /* __kb_putp_aux (bwa/kbtree.h:205-226) */
{
	int i, r;
	if (!x->internal) {
		i = bt_find(x, k->pos, &r);
		if (i != x->n - 1) memmove(x->key + i + 2, x->key + i + 1, (x->n - i - 1) * sizeof(chain_t));
		x->key[i + 1] = *k;
		++x->n;
	} else {
		i = bt_find(x, k->pos, &r) + 1;
		if (x->ptr[i]->n == 2 * BT - 1) {
			bt_split(x, i, x->ptr[i]);
			if (((x->key[i].pos < k->pos) - (k->pos < x->key[i].pos)) > 0) ++i;
		}
		bt_put_aux(x->ptr[i], k);
	}
}
static btnode_t *bt_put(btnode_t *root, const chain_t *k) /* kb_putp (bwa/kbtree.h:227-243) */
{
	if (root->n == 2 * BT - 1) {
		btnode_t *s = calloc(1, sizeof(btnode_t));
		s->internal = 1; s->ptr[0] = root;
		bt_split(s, 0, root);
		root = s;
	}
	bt_put_aux(root, k);
	return root;
}
static void bt_traverse(btnode_t *x, chain_t *out, int *n) /* in-order (__kb_traverse, bwa/kbtree.h:346-370) */
{
	int i;
	if (!x) return;
	for (i = 0; i < x->n; ++i) { if (x->internal) bt_traverse(x->ptr[i], out, n); out[(*n)++] = x->key[i]; }
	if (x->internal) bt_traverse(x->ptr[x->n], out, n);
}
static void bt_free(btnode_t *x) { int i; if (!x) return; if (x->internal) for (i = 0; i <= x->n; ++i) bt_free(x->ptr[i]); free(x); }

static int test_and_merge(const b200_mem_opt_t *opt, i64 l_pac, chain_t *c, const seed_t *p, int seed_rid) /* bwa/bwamem.c:216-237 */
{
	i64 qend, rend, x, y;
	const seed_t *last = &c->seeds[c->n-1];
	qend = last->qbeg + last->len; rend = last->rbeg + last->len;
	if (seed_rid != c->rid) return 0;
	if (p->qbeg >= c->seeds[0].qbeg && p->qbeg + p->len <= qend && p->rbeg >= c->seeds[0].rbeg && p->rbeg + p->len <= rend) return 1;
	if ((last->rbeg < l_pac || c->seeds[0].rbeg < l_pac) && p->rbeg >= l_pac) return 0;
	x = p->qbeg - last->qbeg; y = p->rbeg - last->rbeg;
	if (y >= 0 && x - y <= opt->w && y - x <= opt->w && x - last->len < opt->max_chain_gap && y - last->len < opt->max_chain_gap) {
		if (c->n == c->m) { c->m <<= 1; c->seeds = realloc(c->seeds, c->m * sizeof(seed_t)); }
		c->seeds[c->n++] = *p;
		return 1;
	}
	return 0;
}

static int chain_weight(const chain_t *c) /* mem_chain_weight (bwa/bwamem.c:239-258) */
{
	i64 end; int j, w = 0, tmp;
	for (j = 0, end = 0; j < c->n; ++j) {
		const seed_t *s = &c->seeds[j];
		if (s->qbeg >= end) w += s->len;
		else if (s->qbeg + s->len > end) w += s->qbeg + s->len - end;
		end = end > s->qbeg + s->len? end : s->qbeg + s->len;
	}
	tmp = w; w = 0;
	for (j = 0, end = 0; j < c->n; ++j) {
		const seed_t *s = &c->seeds[j];
		if (s->rbeg >= end) w += s->len;
		else if (s->rbeg + s->len > end) w += s->rbeg + s->len - end;
		end = end > s->rbeg + s->len? end : s->rbeg + s->len;
	}
	w = w < tmp? w : tmp;
	return w < 1<<30? w : (1<<30)-1;
}

/* mem_chain (bwa/bwamem.c:277-341) */
static chain_t *mem_chain(const ctx_t *cx, int len, const u8 *seq, int *n_out)
{
	const b200_mem_opt_t *opt = cx->o; const b200_index_view_t *v = cx->v;
	int i, b, e, l_rep, n = 0, n_keys = 0;
	intv_v mem = {0,0,0};
	btnode_t *root;
	chain_t *out;
	*n_out = 0;
	if (len < opt->min_seed_len) return 0;
	root = calloc(1, sizeof(btnode_t));
	collect_intv(cx, len, seq, &mem);
	for (i = 0, b = e = l_rep = 0; i < (int)mem.n; ++i) {
		intv_t *p = &mem.a[i];
		int sb = (p->info>>32), se = (u32)p->info;
		if (p->x[2] <= (u64)opt->max_occ) continue;
		if (sb > e) l_rep += e - b, b = sb, e = se;
		else e = e > se? e : se;
	}
	l_rep += e - b;
	for (i = 0; i < (int)mem.n; ++i) {
		intv_t *p = &mem.a[i];
		int step, count, slen = (u32)p->info - (p->info>>32);
		i64 k;
		step = p->x[2] > (u64)opt->max_occ? p->x[2] / opt->max_occ : 1;
		for (k = count = 0; k < (i64)p->x[2] && count < opt->max_occ; k += step, ++count) {
			chain_t tmp, *lower; seed_t s; int rid, to_add = 0;
			s.rbeg = tmp.pos = bwt_sa(v, p->x[0] + k);
			s.qbeg = p->info>>32;
			s.score = s.len = slen;
			rid = intv2rid(v, s.rbeg, s.rbeg + s.len);
			if (rid < 0) continue;
			if (n_keys) {
				lower = bt_lower(root, tmp.pos);
				if (!lower || !test_and_merge(opt, v->l_pac, lower, &s, rid)) to_add = 1;
			} else to_add = 1;
			if (to_add) {
				memset(&tmp, 0, sizeof(tmp));
				tmp.pos = s.rbeg; tmp.n = 1; tmp.m = 4;
				tmp.seeds = calloc(tmp.m, sizeof(seed_t));
				tmp.seeds[0] = s; tmp.rid = rid; tmp.is_alt = !!v->contigs[rid].is_alt;
				root = bt_put(root, &tmp); ++n_keys;
			}
		}
	}
	out = malloc((n_keys + 1) * sizeof(chain_t));
	bt_traverse(root, out, &n);
	for (i = 0; i < n; ++i) out[i].frac_rep = (float)l_rep / len;
	bt_free(root); free(mem.a);
	*n_out = n;
	return out;
}

static int flt_lt(const void *a, const void *b) { return ((const chain_t*)a)->w > ((const chain_t*)b)->w; }
#define chn_beg(ch) ((ch).seeds->qbeg)
#define chn_end(ch) ((ch).seeds[(ch).n-1].qbeg + (ch).seeds[(ch).n-1].len)

/* mem_chain_flt (bwa/bwamem.c:353-411) */
static int chain_flt(const b200_mem_opt_t *opt, int n_chn, chain_t *a)
{
	int i, k, n_kept = 0, *chains;
	if (n_chn == 0) return 0;
	for (i = k = 0; i < n_chn; ++i) {
		chain_t *c = &a[i];
		c->first = -1; c->kept = 0;
		c->w = chain_weight(c);
		if ((int)c->w < opt->min_chain_weight) free(c->seeds);
		else a[k++] = *c;
	}
	n_chn = k;
	introsort(n_chn, a, sizeof(chain_t), flt_lt);
	chains = malloc((n_chn + 1) * sizeof(int));
	a[0].kept = 3;
	chains[n_kept++] = 0;
	for (i = 1; i < n_chn; ++i) {
		int large_ovlp = 0;
		for (k = 0; k < n_kept; ++k) {
			int j = chains[k];
			int b_max = chn_beg(a[j]) > chn_beg(a[i])? chn_beg(a[j]) : chn_beg(a[i]);
			int e_min = chn_end(a[j]) < chn_end(a[i])? chn_end(a[j]) : chn_end(a[i]);
			if (e_min > b_max && (!a[j].is_alt || a[i].is_alt)) {
				int li = chn_end(a[i]) - chn_beg(a[i]);
				int lj = chn_end(a[j]) - chn_beg(a[j]);
				int min_l = li < lj? li : lj;
				if (e_min - b_max >= min_l * opt->mask_level && min_l < opt->max_chain_gap) {
					large_ovlp = 1;
					if (a[j].first < 0) a[j].first = i;
					if ((int)a[i].w < (int)a[j].w * opt->drop_ratio && (int)a[j].w - (int)a[i].w >= opt->min_seed_len<<1) break;
				}
			}
		}
		if (k == n_kept) { chains[n_kept++] = i; a[i].kept = large_ovlp? 2 : 3; }
	}
	for (i = 0; i < n_kept; ++i) { chain_t *c = &a[chains[i]]; if (c->first >= 0) a[c->first].kept = 1; }
	free(chains);
	for (i = k = 0; i < n_chn; ++i) {
		if (a[i].kept == 0 || a[i].kept == 3) continue;
		if (++k >= opt->max_chain_extend) break;
	}
	for (; i < n_chn; ++i) if (a[i].kept < 3) a[i].kept = 0;
	for (i = k = 0; i < n_chn; ++i) {
		chain_t *c = &a[i];
		if (c->kept == 0) free(c->seeds);
		else a[k++] = a[i];
	}
	return k;
}

/* ------------------------------------------------------------------ ksw (bwa/ksw.c) */

typedef struct { int32_t h, e; } eh_t;

/* ksw_extend2 (bwa/ksw.c:416-515) */
static int ksw_extend2(int qlen, const u8 *query, int tlen, const u8 *target, const int8_t *mat, int o_del, int e_del, int o_ins, int e_ins,
                       int w, int end_bonus, int zdrop, int h0, int *_qle, int *_tle, int *_gtle, int *_gscore, int *_max_off)
{
	eh_t *eh = calloc(qlen + 1, 8);
	int i, j, oe_del = o_del + e_del, oe_ins = o_ins + e_ins, beg, end, max, max_i, max_j, max_ins, max_del, max_ie, gscore, max_off;
	eh[0].h = h0; eh[1].h = h0 > oe_ins? h0 - oe_ins : 0;
	for (j = 2; j <= qlen && eh[j-1].h > e_ins; ++j) eh[j].h = eh[j-1].h - e_ins;
	for (i = 0, max = 0; i < 25; ++i) max = max > mat[i]? max : mat[i];
	max_ins = (int)((double)(qlen * max + end_bonus - o_ins) / e_ins + 1.);
	max_ins = max_ins > 1? max_ins : 1;
	w = w < max_ins? w : max_ins;
	max_del = (int)((double)(qlen * max + end_bonus - o_del) / e_del + 1.);
	max_del = max_del > 1? max_del : 1;
	w = w < max_del? w : max_del;
	max = h0, max_i = max_j = -1; max_ie = -1, gscore = -1; max_off = 0;
	beg = 0, end = qlen;
	for (i = 0; i < tlen; ++i) {
		int t, f = 0, h1, m = 0, mj = -1;
		const int8_t *q = &mat[target[i] * 5];
		if (beg < i - w) beg = i - w;
		if (end > i + w + 1) end = i + w + 1;
		if (end > qlen) end = qlen;
		if (beg == 0) { h1 = h0 - (o_del + e_del * (i + 1)); if (h1 < 0) h1 = 0; } else h1 = 0;
		for (j = beg; j < end; ++j) {
			eh_t *p = &eh[j];
			int h, M = p->h, e = p->e;
			p->h = h1;
			M = M? M + q[query[j]] : 0;
			h = M > e? M : e;
			h = h > f? h : f;
			h1 = h;
			mj = m > h? mj : j;
			m = m > h? m : h;
			t = M - oe_del; t = t > 0? t : 0;
			e -= e_del; e = e > t? e : t;
			p->e = e;
			t = M - oe_ins; t = t > 0? t : 0;
			f -= e_ins; f = f > t? f : t;
		}
		eh[end].h = h1; eh[end].e = 0;
		if (j == qlen) { max_ie = gscore > h1? max_ie : i; gscore = gscore > h1? gscore : h1; }
		if (m == 0) break;
		if (m > max) {
			max = m, max_i = i, max_j = mj;
			max_off = max_off > abs(mj - i)? max_off : abs(mj - i);
		} else if (zdrop > 0) {
			if (i - max_i > mj - max_j) { if (max - m - ((i - max_i) - (mj - max_j)) * e_del > zdrop) break; }
			else { if (max - m - ((mj - max_j) - (i - max_i)) * e_ins > zdrop) break; }
		}
		for (j = beg; j < end && eh[j].h == 0 && eh[j].e == 0; ++j);
		beg = j;
		for (j = end; j >= beg && eh[j].h == 0 && eh[j].e == 0; --j);
		end = j + 2 < qlen? j + 2 : qlen;
	}
	free(eh);
	*_qle = max_j + 1; *_tle = max_i + 1; *_gtle = max_ie + 1; *_gscore = gscore; *_max_off = max_off;
	return max;
}

#define MINUS_INF -0x40000000

/* ksw_global2 (bwa/ksw.c:540-642); cigar_ may be NULL for score only */
static int ksw_global2(int qlen, const u8 *query, int tlen, const u8 *target, const int8_t *mat, int o_del, int e_del, int o_ins, int e_ins, int w,
                       int *n_cigar_, u32 **cigar_)
{
	eh_t *eh;
	int i, j, k, oe_del = o_del + e_del, oe_ins = o_ins + e_ins, score, n_col;
	u8 *z;
	if (n_cigar_) *n_cigar_ = 0;
	n_col = qlen < 2*w+1? qlen : 2*w+1;
	z = n_cigar_ && cigar_? malloc((long)n_col * tlen + 1) : 0;
	eh = calloc(qlen + 1, 8);
	eh[0].h = 0; eh[0].e = MINUS_INF;
	for (j = 1; j <= qlen && j <= w; ++j) eh[j].h = -(o_ins + e_ins * j), eh[j].e = MINUS_INF;
	for (; j <= qlen; ++j) eh[j].h = eh[j].e = MINUS_INF;
	for (i = 0; i < tlen; ++i) {
		int32_t f = MINUS_INF, h1, beg, end, t;
		const int8_t *q = &mat[target[i] * 5];
		u8 *zi = z? &z[(long)i * n_col] : 0;
		beg = i > w? i - w : 0;
		end = i + w + 1 < qlen? i + w + 1 : qlen;
		h1 = beg == 0? -(o_del + e_del * (i + 1)) : MINUS_INF;
		for (j = beg; j < end; ++j) {
			eh_t *p = &eh[j];
			int32_t h, m = p->h, e = p->e;
			u8 d;
			p->h = h1;
			m += q[query[j]];
			d = m >= e? 0 : 1;
			h = m >= e? m : e;
			d = h >= f? d : 2;
			h = h >= f? h : f;
			h1 = h;
			t = m - oe_del;
			e -= e_del;
			d |= e > t? 1<<2 : 0;
			e  = e > t? e : t;
			p->e = e;
			t = m - oe_ins;
			f -= e_ins;
			d |= f > t? 2<<4 : 0;
			f  = f > t? f : t;
			if (zi) zi[j - beg] = d;
		}
		eh[end].h = h1; eh[end].e = MINUS_INF;
	}
	score = eh[qlen].h;
	if (z) {
		int n_cigar = 0, m_cigar = 0, which = 0;
		u32 *cigar = 0, tmp;
#define PUSH(op, len) do { if (n_cigar == 0 || (u32)(op) != (cigar[n_cigar-1]&0xf)) { if (n_cigar == m_cigar) { m_cigar = m_cigar? m_cigar<<1 : 4; cigar = realloc(cigar, m_cigar << 2); } \
		cigar[n_cigar++] = (u32)(len)<<4 | (op); } else cigar[n_cigar-1] += (u32)(len)<<4; } while (0)
		i = tlen - 1; k = (i + w + 1 < qlen? i + w + 1 : qlen) - 1;
		while (i >= 0 && k >= 0) {
			which = z[(long)i * n_col + (k - (i > w? i - w : 0))] >> (which<<1) & 3;
			if (which == 0) { PUSH(0, 1); --i; --k; }
			else if (which == 1) { PUSH(2, 1); --i; }
			else { PUSH(1, 1); --k; }
		}
		if (i >= 0) PUSH(2, i + 1);
		if (k >= 0) PUSH(1, k + 1);
#undef PUSH
		for (i = 0; i < n_cigar>>1; ++i) tmp = cigar[i], cigar[i] = cigar[n_cigar-1-i], cigar[n_cigar-1-i] = tmp;
		*n_cigar_ = n_cigar; *cigar_ = cigar;
	}
	free(eh); free(z);
	return score;
}

/* ------------------------------------------------------------------ chain -> regions (bwa/bwamem.c:647-812) */

typedef struct {
	i64 rb, re; int qb, qe, rid, score, truesc, sub, alt_sc, csub, sub_n, w, seedcov, secondary, secondary_all, seedlen0, n_comp, is_alt;
	float frac_rep; u64 hash;
} reg_t;
typedef struct { size_t n, m; reg_t *a; } reg_v;

static int cal_max_gap(const b200_mem_opt_t *opt, int qlen) /* bwa/bwamem.c:647-654 */
{
	int l_del = (int)((double)(qlen * opt->a - opt->o_del) / opt->e_del + 1.);
	int l_ins = (int)((double)(qlen * opt->a - opt->o_ins) / opt->e_ins + 1.);
	int l = l_del > l_ins? l_del : l_ins;
	l = l > 1? l : 1;
	return l < opt->w<<1? l : opt->w<<1;
}

#define MAX_BAND_TRY 2

/* mem_chain2aln (bwa/bwamem.c:658-812) */
static void chain2aln(const ctx_t *cx, int l_query, const u8 *query, const chain_t *c, reg_v *av)
{
	const b200_mem_opt_t *opt = cx->o; const b200_index_view_t *v = cx->v;
	int i, k, max_off[2], aw[2];
	i64 l_pac = v->l_pac, rmax[2], tmp, max = 0, rlen;
	const seed_t *s;
	u8 *rseq = 0;
	u64 *srt;
	if (c->n == 0) return;
	rmax[0] = l_pac<<1; rmax[1] = 0;
	for (i = 0; i < c->n; ++i) {
		i64 b, e; const seed_t *t = &c->seeds[i];
		b = t->rbeg - (t->qbeg + cal_max_gap(opt, t->qbeg));
		e = t->rbeg + t->len + ((l_query - t->qbeg - t->len) + cal_max_gap(opt, l_query - t->qbeg - t->len));
		rmax[0] = rmax[0] < b? rmax[0] : b;
		rmax[1] = rmax[1] > e? rmax[1] : e;
		if (t->len > max) max = t->len;
	}
	rmax[0] = rmax[0] > 0? rmax[0] : 0;
	rmax[1] = rmax[1] < l_pac<<1? rmax[1] : l_pac<<1;
	if (rmax[0] < l_pac && l_pac < rmax[1]) {
		if (c->seeds[0].rbeg < l_pac) rmax[1] = l_pac; else rmax[0] = l_pac;
	}
	{ /* bns_fetch_seq (bwa/bntseq.c:426-451) */
		int is_rev, rid = pos2rid(v, depos(v, c->seeds[0].rbeg, &is_rev));
		i64 far_beg = v->contigs[rid].offset, far_end = far_beg + v->contigs[rid].len;
		if (is_rev) { i64 t2 = far_beg; far_beg = (l_pac<<1) - far_end; far_end = (l_pac<<1) - t2; }
		rmax[0] = rmax[0] > far_beg? rmax[0] : far_beg;
		rmax[1] = rmax[1] < far_end? rmax[1] : far_end;
		rseq = get_seq(v, rmax[0], rmax[1], &rlen);
	}
	srt = malloc(c->n * 8);
	for (i = 0; i < c->n; ++i) srt[i] = (u64)c->seeds[i].score<<32 | i;
	introsort(c->n, srt, 8, u64_lt);
	for (k = c->n - 1; k >= 0; --k) {
		reg_t *a;
		s = &c->seeds[(u32)srt[k]];
		for (i = 0; i < (int)av->n; ++i) {
			reg_t *p = &av->a[i]; i64 rd; int qd, w, max_gap;
			if (s->rbeg < p->rb || s->rbeg + s->len > p->re || s->qbeg < p->qb || s->qbeg + s->len > p->qe) continue;
			if (s->len - p->seedlen0 > .1 * l_query) continue;
			qd = s->qbeg - p->qb; rd = s->rbeg - p->rb;
			max_gap = cal_max_gap(opt, qd < rd? qd : rd);
			w = max_gap < p->w? max_gap : p->w;
			if (qd - rd < w && rd - qd < w) break;
			qd = p->qe - (s->qbeg + s->len); rd = p->re - (s->rbeg + s->len);
			max_gap = cal_max_gap(opt, qd < rd? qd : rd);
			w = max_gap < p->w? max_gap : p->w;
			if (qd - rd < w && rd - qd < w) break;
		}
		if (i < (int)av->n) {
			for (i = k + 1; i < c->n; ++i) {
				const seed_t *t;
				if (srt[i] == 0) continue;
				t = &c->seeds[(u32)srt[i]];
				if (t->len < s->len * .95) continue;
				if (s->qbeg <= t->qbeg && s->qbeg + s->len - t->qbeg >= s->len>>2 && t->qbeg - s->qbeg != t->rbeg - s->rbeg) break;
				if (t->qbeg <= s->qbeg && t->qbeg + t->len - s->qbeg >= s->len>>2 && s->qbeg - t->qbeg != s->rbeg - t->rbeg) break;
			}
			if (i == c->n) { srt[k] = 0; continue; }
		}
		if (av->n == av->m) { av->m = av->m? av->m<<1 : 4; av->a = realloc(av->a, av->m * sizeof(reg_t)); }
		a = &av->a[av->n++];
		memset(a, 0, sizeof(reg_t));
		a->w = aw[0] = aw[1] = opt->w;
		a->score = a->truesc = -1;
		a->rid = c->rid;
		if (s->qbeg) {
			u8 *rs, *qs; int qle, tle, gtle, gscore;
			qs = malloc(s->qbeg);
			for (i = 0; i < s->qbeg; ++i) qs[i] = query[s->qbeg - 1 - i];
			tmp = s->rbeg - rmax[0];
			rs = malloc(tmp + 1);
			for (i = 0; i < tmp; ++i) rs[i] = rseq[tmp - 1 - i];
			for (i = 0; i < MAX_BAND_TRY; ++i) {
				int prev = a->score;
				aw[0] = opt->w << i;
				a->score = ksw_extend2(s->qbeg, qs, tmp, rs, opt->mat, opt->o_del, opt->e_del, opt->o_ins, opt->e_ins, aw[0], opt->pen_clip5, opt->zdrop, s->len * opt->a, &qle, &tle, &gtle, &gscore, &max_off[0]);
				if (a->score == prev || max_off[0] < (aw[0]>>1) + (aw[0]>>2)) break;
			}
			if (gscore <= 0 || gscore <= a->score - opt->pen_clip5) { a->qb = s->qbeg - qle, a->rb = s->rbeg - tle; a->truesc = a->score; }
			else { a->qb = 0, a->rb = s->rbeg - gtle; a->truesc = gscore; }
			free(qs); free(rs);
		} else a->score = a->truesc = s->len * opt->a, a->qb = 0, a->rb = s->rbeg;
		if (s->qbeg + s->len != l_query) {
			int qle, tle, qe, re, gtle, gscore, sc0 = a->score;
			qe = s->qbeg + s->len;
			re = s->rbeg + s->len - rmax[0];
			for (i = 0; i < MAX_BAND_TRY; ++i) {
				int prev = a->score;
				aw[1] = opt->w << i;
				a->score = ksw_extend2(l_query - qe, query + qe, rmax[1] - rmax[0] - re, rseq + re, opt->mat, opt->o_del, opt->e_del, opt->o_ins, opt->e_ins, aw[1], opt->pen_clip3, opt->zdrop, sc0, &qle, &tle, &gtle, &gscore, &max_off[1]);
				if (a->score == prev || max_off[1] < (aw[1]>>1) + (aw[1]>>2)) break;
			}
			if (gscore <= 0 || gscore <= a->score - opt->pen_clip3) { a->qe = qe + qle, a->re = rmax[0] + re + tle; a->truesc += a->score - sc0; }
			else { a->qe = l_query, a->re = rmax[0] + re + gtle; a->truesc += gscore - sc0; }
		} else a->qe = l_query, a->re = s->rbeg + s->len;
		for (i = 0, a->seedcov = 0; i < c->n; ++i) {
			const seed_t *t = &c->seeds[i];
			if (t->qbeg >= a->qb && t->qbeg + t->len <= a->qe && t->rbeg >= a->rb && t->rbeg + t->len <= a->re) a->seedcov += t->len;
		}
		a->w = aw[0] > aw[1]? aw[0] : aw[1];
		a->seedlen0 = s->len;
		a->frac_rep = c->frac_rep;
	}
	free(srt); free(rseq);
}

/* ------------------------------------------------------------------ CIGAR (bwa/bwa.c:148-234) */

typedef struct { char *s; size_t l, m; } str_t;
static void sputc(str_t *s, int c) { if (s->l + 2 > s->m) { s->m = s->m? s->m << 1 : 64; s->s = realloc(s->s, s->m); } s->s[s->l++] = c; s->s[s->l] = 0; }
static void sputw(str_t *s, int v) { char b[16]; int l = 0; if (v == 0) b[l++] = '0'; while (v > 0) { b[l++] = '0' + v % 10; v /= 10; } while (l > 0) sputc(s, b[--l]); }

/* bwa_gen_cigar2: returns malloc'd cigar (or NULL); *md (optional) receives a malloc'd MD string */
static u32 *gen_cigar2(const ctx_t *cx, int w_, int l_query, u8 *query, i64 rb, i64 re, int *score, int *n_cigar, int *NM, char **md)
{
	const b200_mem_opt_t *opt = cx->o; const b200_index_view_t *v = cx->v;
	const int8_t *mat = opt->mat;
	u32 *cigar = 0;
	u8 tmp, *rseq;
	int i;
	i64 rlen, l_pac = v->l_pac;
	if (n_cigar) *n_cigar = 0;
	if (NM) *NM = -1;
	if (md) *md = 0;
	if (l_query <= 0 || rb >= re || (rb < l_pac && re > l_pac)) return 0;
	rseq = get_seq(v, rb, re, &rlen);
	if (re - rb != rlen) goto ret;
	if (rb >= l_pac) {
		for (i = 0; i < l_query>>1; ++i) tmp = query[i], query[i] = query[l_query - 1 - i], query[l_query - 1 - i] = tmp;
		for (i = 0; i < rlen>>1; ++i) tmp = rseq[i], rseq[i] = rseq[rlen - 1 - i], rseq[rlen - 1 - i] = tmp;
	}
	if (l_query == re - rb && w_ == 0) {
		if (n_cigar) { cigar = malloc(4); cigar[0] = l_query<<4 | 0; *n_cigar = 1; }
		for (i = 0, *score = 0; i < l_query; ++i) *score += mat[rseq[i]*5 + query[i]];
	} else {
		int w, max_gap, max_ins, max_del, min_w;
		max_ins = (int)((double)(((l_query+1)>>1) * mat[0] - opt->o_ins) / opt->e_ins + 1.);
		max_del = (int)((double)(((l_query+1)>>1) * mat[0] - opt->o_del) / opt->e_del + 1.);
		max_gap = max_ins > max_del? max_ins : max_del;
		max_gap = max_gap > 1? max_gap : 1;
		w = (max_gap + abs((int)rlen - l_query) + 1) >> 1;
		w = w < w_? w : w_;
		min_w = abs((int)rlen - l_query) + 3;
		w = w > min_w? w : min_w;
		*score = ksw_global2(l_query, query, rlen, rseq, mat, opt->o_del, opt->e_del, opt->o_ins, opt->e_ins, w, n_cigar, n_cigar? &cigar : 0);
	}
	if (NM && n_cigar) {
		int k, x, y, u, n_mm = 0, n_gap = 0;
		str_t str = {0,0,0};
		const char *int2base = rb < l_pac? "ACGTN" : "TGCAN";
		for (k = 0, x = y = u = 0; k < *n_cigar; ++k) {
			int op = cigar[k]&0xf, len = cigar[k]>>4;
			if (op == 0) {
				for (i = 0; i < len; ++i) {
					if (query[x + i] != rseq[y + i]) { sputw(&str, u); sputc(&str, int2base[rseq[y+i]]); ++n_mm; u = 0; }
					else ++u;
				}
				x += len; y += len;
			} else if (op == 2) {
				if (k > 0 && k < *n_cigar - 1) {
					sputw(&str, u); sputc(&str, '^');
					for (i = 0; i < len; ++i) sputc(&str, int2base[rseq[y+i]]);
					u = 0; n_gap += len;
				}
				y += len;
			} else if (op == 1) x += len, n_gap += len;
		}
		sputw(&str, u);
		*NM = n_mm + n_gap;
		if (md) *md = str.s; else free(str.s);
	}
	if (rb >= l_pac)
		for (i = 0; i < l_query>>1; ++i) tmp = query[i], query[i] = query[l_query - 1 - i], query[l_query - 1 - i] = tmp;
ret:
	free(rseq);
	return cigar;
}

/* ------------------------------------------------------------------ de-duplication, primary marking (bwa/bwamem.c:418-584) */

static int ars2_lt(const void *a, const void *b) { return ((const reg_t*)a)->re < ((const reg_t*)b)->re; }
static int ars_lt(const void *a_, const void *b_) { const reg_t *a = a_, *b = b_; return a->score > b->score || (a->score == b->score && (a->rb < b->rb || (a->rb == b->rb && a->qb < b->qb))); }
static int hash_lt(const void *a_, const void *b_) { const reg_t *a = a_, *b = b_; return a->score > b->score || (a->score == b->score && (a->is_alt < b->is_alt || (a->is_alt == b->is_alt && a->hash < b->hash))); }
static int hash2_lt(const void *a_, const void *b_) { const reg_t *a = a_, *b = b_; return a->is_alt < b->is_alt || (a->is_alt == b->is_alt && (a->score > b->score || (a->score == b->score && a->hash < b->hash))); }

#define PATCH_MAX_R_BW 0.05f
#define PATCH_MIN_SC_RATIO 0.90f

static int patch_reg(const ctx_t *cx, u8 *query, const reg_t *a, const reg_t *b, int *_w) /* mem_patch_reg (bwa/bwamem.c:432-461) */
{
	const b200_mem_opt_t *opt = cx->o;
	int w, score = 0, q_s, r_s;
	double r;
	if (a->rb < cx->v->l_pac && b->rb >= cx->v->l_pac) return 0;
	if (a->qb >= b->qb || a->qe >= b->qe || a->re >= b->re) return 0;
	w = (a->re - b->rb) - (a->qe - b->qb);
	w = w > 0? w : -w;
	r = (double)(a->re - b->rb) / (b->re - a->rb) - (double)(a->qe - b->qb) / (b->qe - a->qb);
	r = r > 0.? r : -r;
	if (a->re < b->rb || a->qe < b->qb) { if (w > opt->w<<1 || r >= PATCH_MAX_R_BW) return 0; }
	else if (w > opt->w<<2 || r >= PATCH_MAX_R_BW*2) return 0;
	w += a->w + b->w;
	w = w < opt->w<<2? w : opt->w<<2;
	gen_cigar2(cx, w, b->qe - a->qb, query + a->qb, a->rb, b->re, &score, 0, 0, 0);
	q_s = (int)((double)(b->qe - a->qb) / ((b->qe - b->qb) + (a->qe - a->qb)) * (b->score + a->score) + .499);
	r_s = (int)((double)(b->re - a->rb) / ((b->re - b->rb) + (a->re - a->rb)) * (b->score + a->score) + .499);
	if ((double)score / (q_s > r_s? q_s : r_s) < PATCH_MIN_SC_RATIO) return 0;
	*_w = w;
	return score;
}

static int sort_dedup_patch(const ctx_t *cx, u8 *query, int n, reg_t *a) /* mem_sort_dedup_patch (bwa/bwamem.c:463-515) */
{
	const b200_mem_opt_t *opt = cx->o;
	int m, i, j;
	if (n <= 1) return n;
	introsort(n, a, sizeof(reg_t), ars2_lt);
	for (i = 0; i < n; ++i) a[i].n_comp = 1;
	for (i = 1; i < n; ++i) {
		reg_t *p = &a[i];
		if (p->rid != a[i-1].rid || p->rb >= a[i-1].re + opt->max_chain_gap) continue;
		for (j = i - 1; j >= 0 && p->rid == a[j].rid && p->rb < a[j].re + opt->max_chain_gap; --j) {
			reg_t *q = &a[j];
			i64 or_, oq, mr, mq; int score, w;
			if (q->qe == q->qb) continue;
			or_ = q->re - p->rb;
			oq = q->qb < p->qb? q->qe - p->qb : p->qe - q->qb;
			mr = q->re - q->rb < p->re - p->rb? q->re - q->rb : p->re - p->rb;
			mq = q->qe - q->qb < p->qe - p->qb? q->qe - q->qb : p->qe - p->qb;
			if (or_ > opt->mask_level_redun * mr && oq > opt->mask_level_redun * mq) {
				if (p->score < q->score) { p->qe = p->qb; break; }
				else q->qe = q->qb;
			} else if (q->rb < p->rb && (score = patch_reg(cx, query, q, p, &w)) > 0) {
				p->n_comp += q->n_comp + 1;
				p->seedcov = p->seedcov > q->seedcov? p->seedcov : q->seedcov;
				p->sub = p->sub > q->sub? p->sub : q->sub;
				p->csub = p->csub > q->csub? p->csub : q->csub;
				p->qb = q->qb, p->rb = q->rb;
				p->truesc = p->score = score;
				p->w = w;
				q->qb = q->qe;
			}
		}
	}
	for (i = 0, m = 0; i < n; ++i) if (a[i].qe > a[i].qb) { if (m != i) a[m++] = a[i]; else ++m; }
	n = m;
	introsort(n, a, sizeof(reg_t), ars_lt);
	for (i = 1; i < n; ++i) if (a[i].score == a[i-1].score && a[i].rb == a[i-1].rb && a[i].qb == a[i-1].qb) a[i].qe = a[i].qb;
	for (i = 1, m = 1; i < n; ++i) if (a[i].qe > a[i].qb) { if (m != i) a[m++] = a[i]; else ++m; }
	return m;
}

static u64 hash_64(u64 key) /* bwa/utils.h:98-109 */
{
	key += ~(key << 32); key ^= (key >> 22); key += ~(key << 13); key ^= (key >> 8);
	key += (key << 3); key ^= (key >> 15); key += ~(key << 27); key ^= (key >> 31);
	return key;
}

static void mark_primary_core(const b200_mem_opt_t *opt, int n, reg_t *a, int *z, int *nz_) /* bwa/bwamem.c:519-545 */
{
	int i, k, tmp, nz = 0;
	tmp = opt->a + opt->b;
	tmp = opt->o_del + opt->e_del > tmp? opt->o_del + opt->e_del : tmp;
	tmp = opt->o_ins + opt->e_ins > tmp? opt->o_ins + opt->e_ins : tmp;
	z[nz++] = 0;
	for (i = 1; i < n; ++i) {
		for (k = 0; k < nz; ++k) {
			int j = z[k];
			int b_max = a[j].qb > a[i].qb? a[j].qb : a[i].qb;
			int e_min = a[j].qe < a[i].qe? a[j].qe : a[i].qe;
			if (e_min > b_max) {
				int min_l = a[i].qe - a[i].qb < a[j].qe - a[j].qb? a[i].qe - a[i].qb : a[j].qe - a[j].qb;
				if (e_min - b_max >= min_l * opt->mask_level) {
					if (a[j].sub == 0) a[j].sub = a[i].score;
					if (a[j].score - a[i].score <= tmp && (a[j].is_alt || !a[i].is_alt)) ++a[j].sub_n;
					break;
				}
			}
		}
		if (k == nz) z[nz++] = i;
		else a[i].secondary = z[k];
	}
	*nz_ = nz;
}

static int mark_primary_se(const b200_mem_opt_t *opt, int n, reg_t *a, i64 id) /* bwa/bwamem.c:547-584 */
{
	int i, n_pri, nz, *z;
	if (n == 0) return 0;
	z = malloc(2 * (size_t)(n > 0 ? n : 0) * sizeof(int) + sizeof(int));
	for (i = n_pri = 0; i < n; ++i) {
		a[i].sub = a[i].alt_sc = 0, a[i].secondary = a[i].secondary_all = -1, a[i].hash = hash_64(id+i);
		if (!a[i].is_alt) ++n_pri;
	}
	introsort(n, a, sizeof(reg_t), hash_lt);
	mark_primary_core(opt, n, a, z, &nz);
	for (i = 0; i < n; ++i) {
		reg_t *p = &a[i];
		p->secondary_all = i;
		if (!p->is_alt && p->secondary >= 0 && a[p->secondary].is_alt) p->alt_sc = a[p->secondary].score;
	}
	if (n_pri >= 0 && n_pri < n) {
		if (n_pri > 0) introsort(n, a, sizeof(reg_t), hash2_lt);
		for (i = 0; i < n; ++i) z[a[i].secondary_all] = i;
		for (i = 0; i < n; ++i) {
			if (a[i].secondary >= 0) {
				a[i].secondary_all = z[a[i].secondary];
				if (a[i].is_alt) a[i].secondary = 0x7fffffff;
			} else a[i].secondary_all = -1;
		}
		if (n_pri > 0) {
			for (i = 0; i < n_pri; ++i) a[i].sub = 0, a[i].secondary = -1;
			mark_primary_core(opt, n_pri, a, z, &nz);
		}
	} else for (i = 0; i < n; ++i) a[i].secondary_all = a[i].secondary;
	free(z);
	return n_pri;
}

/* ------------------------------------------------------------------ region -> alignment (bwa/bwamem.c:818-825,982-1006,1119-1189) */

static int approx_mapq_se(const b200_mem_opt_t *opt, const reg_t *a)
{
	int mapq, l, sub = a->sub? a->sub : opt->min_seed_len * opt->a;
	double identity;
	sub = a->csub > sub? a->csub : sub;
	if (sub >= a->score) return 0;
	l = a->qe - a->qb > a->re - a->rb? a->qe - a->qb : a->re - a->rb;
	identity = 1. - (double)(l * opt->a - a->score) / (opt->a + opt->b) / l;
	if (a->score == 0) mapq = 0;
	else if (opt->mapQ_coef_len > 0) {
		double tmp;
		tmp = l < opt->mapQ_coef_len? 1. : opt->mapQ_coef_fac / log(l);
		tmp *= identity * identity;
		mapq = (int)(6.02 * (a->score - sub) / opt->a * tmp * tmp + .499);
	} else {
		mapq = (int)(30.0 * (1. - (double)sub / a->score) * log(a->seedcov) + .499);
		mapq = identity < 0.95? (int)(mapq * identity * identity + .499) : mapq;
	}
	if (a->sub_n > 0) mapq -= (int)(4.343 * log(a->sub_n+1) + .499);
	if (mapq > 60) mapq = 60;
	if (mapq < 0) mapq = 0;
	mapq = (int)(mapq * (1. - a->frac_rep) + .499);
	return mapq;
}

static int infer_bw(int l1, int l2, int score, int a, int q, int r)
{
	int w;
	if (l1 == l2 && l1 * a - score < (q + r - a)<<1) return 0;
	w = ((double)((l1 < l2? l1 : l2) * a - score - q) / r + 2.);
	if (w < abs(l1 - l2)) w = abs(l1 - l2);
	return w;
}

/* fills the alignment part of one hit; cigar/md are malloc'd */
static void reg2aln(const ctx_t *cx, int l_query, u8 *query, const reg_t *ar, b200_hit_t *h, u32 **cigar_out, char **md_out)
{
	const b200_mem_opt_t *opt = cx->o; const b200_index_view_t *v = cx->v;
	int i, w2, tmp, qb, qe, NM = 0, score = 0, is_rev, last_sc = -(1<<30), n_cigar = 0;
	i64 pos, rb, re;
	u32 *cigar = 0; char *md = 0;
	qb = ar->qb, qe = ar->qe; rb = ar->rb, re = ar->re;
	h->mapq = ar->secondary < 0? approx_mapq_se(opt, ar) : 0;
	h->flag = ar->secondary >= 0? 0x100 : 0;
	tmp = infer_bw(qe - qb, re - rb, ar->truesc, opt->a, opt->o_del, opt->e_del);
	w2  = infer_bw(qe - qb, re - rb, ar->truesc, opt->a, opt->o_ins, opt->e_ins);
	w2 = w2 > tmp? w2 : tmp;
	if (w2 > opt->w) w2 = w2 < ar->w? w2 : ar->w;
	i = 0;
	do {
		free(cigar); free(md);
		w2 = w2 < opt->w<<2? w2 : opt->w<<2;
		cigar = gen_cigar2(cx, w2, qe - qb, &query[qb], rb, re, &score, &n_cigar, &NM, &md);
		if (score == last_sc || w2 == opt->w<<2) break;
		last_sc = score;
		w2 <<= 1;
	} while (++i < 3 && score < ar->truesc - opt->a);
	pos = depos(v, rb < v->l_pac? rb : re - 1, &is_rev);
	if (n_cigar > 0) {
		if ((cigar[0]&0xf) == 2) { pos += cigar[0]>>4; --n_cigar; memmove(cigar, cigar + 1, n_cigar * 4); }
		else if ((cigar[n_cigar-1]&0xf) == 2) --n_cigar;
	}
	if (qb != 0 || qe != l_query) {
		int clip5 = is_rev? l_query - qe : qb, clip3 = is_rev? qb : l_query - qe;
		cigar = realloc(cigar, 4 * (n_cigar + 2));
		if (clip5) { memmove(cigar+1, cigar, n_cigar * 4); cigar[0] = clip5<<4 | 3; ++n_cigar; }
		if (clip3) cigar[n_cigar++] = clip3<<4 | 3;
	}
	h->is_rev = is_rev; h->NM = NM; h->n_cigar = n_cigar;
	{ int rid = pos2rid(v, pos); h->pos = pos - v->contigs[rid].offset; if (rid != ar->rid) h->rid = -1000; }
	h->aln_sub = ar->sub > ar->csub? ar->sub : ar->csub;
	h->md_len = md? (int)strlen(md) : 0;
	*cigar_out = cigar; *md_out = md;
}

/* ------------------------------------------------------------------ driver: mem_align1 + mem_reg2aln for every read */

static const u8 nt4_tab(u8 c)
{
	switch (c) { case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2; case 'T': case 't': return 3; case '-': return 5; default: return c < 4? c : 4; }
}

typedef struct { i64 n_reads, n_hits, n_cigar, n_md; i64 *hit_off; b200_hit_t *hits; u32 *cigar; char *md; } ores_t;

void *oracle_align(const b200_index_view_t *v, const b200_mem_opt_t *opt, i64 n, const char *seqs, const i64 *off, const i64 *ids)
{
	ctx_t cx = { v, opt };
	ores_t *R = calloc(1, sizeof(ores_t));
	i64 r, mh = 1024, mc = 4096, mm = 16384;
	R->n_reads = n; R->hit_off = calloc(n + 1, 8);
	R->hits = malloc(mh * sizeof(b200_hit_t)); R->cigar = malloc(mc * 4); R->md = malloc(mm);
	for (r = 0; r < n; ++r) {
		int len = off[r+1] - off[r], i, n_chn;
		u8 *seq = malloc(len + 1);
		chain_t *chn;
		reg_v regs = {0,0,0};
		for (i = 0; i < len; ++i) seq[i] = nt4_tab(seqs[off[r] + i]);
		chn = mem_chain(&cx, len, seq, &n_chn);                      /* mem_align1_core (bwa/bwamem.c:1081-1117) */
		n_chn = chain_flt(opt, n_chn, chn);
		/* mem_flt_chained_seeds (bwa/bwamem.c:624-641) returns at once for reads below ~730 bp; longer reads are outside this oracle */
		for (i = 0; i < n_chn; ++i) { chain2aln(&cx, len, seq, &chn[i], &regs); free(chn[i].seeds); }
		free(chn);
		regs.n = sort_dedup_patch(&cx, seq, regs.n, regs.a);
		for (i = 0; i < (int)regs.n; ++i) if (regs.a[i].rid >= 0 && v->contigs[regs.a[i].rid].is_alt) regs.a[i].is_alt = 1;
		mark_primary_se(opt, regs.n, regs.a, ids[r]);                /* mem_align1 (bwa/bwamem_extra.c:112) */
		R->hit_off[r] = R->n_hits;
		for (i = 0; i < (int)regs.n; ++i) {                          /* BWAAligner::alignSequence loop (src/BWAAligner.cpp:117-128) */
			reg_t *a = &regs.a[i];
			b200_hit_t h; u32 *cg = 0; char *md = 0;
			memset(&h, 0, sizeof(h));
			h.rb = a->rb; h.re = a->re; h.qb = a->qb; h.qe = a->qe; h.rid = a->rid; h.score = a->score; h.truesc = a->truesc; h.sub = a->sub;
			h.alt_sc = a->alt_sc; h.csub = a->csub; h.sub_n = a->sub_n; h.w = a->w; h.seedcov = a->seedcov; h.secondary = a->secondary;
			h.secondary_all = a->secondary_all; h.seedlen0 = a->seedlen0; h.n_comp = a->n_comp; h.is_alt = a->is_alt; h.frac_rep = a->frac_rep; h.hash = a->hash;
			reg2aln(&cx, len, seq, a, &h, &cg, &md);
			if (R->n_hits == mh) { mh <<= 1; R->hits = realloc(R->hits, mh * sizeof(b200_hit_t)); }
			while (R->n_cigar + h.n_cigar > mc) { mc <<= 1; R->cigar = realloc(R->cigar, mc * 4); }
			while (R->n_md + h.md_len + 1 > mm) { mm <<= 1; R->md = realloc(R->md, mm); }
			h.cigar_off = R->n_cigar; h.md_off = R->n_md;
			if (h.n_cigar) memcpy(R->cigar + R->n_cigar, cg, 4 * h.n_cigar);
			if (md) memcpy(R->md + R->n_md, md, h.md_len);
			R->md[R->n_md + h.md_len] = 0;
			R->n_cigar += h.n_cigar; R->n_md += h.md_len + 1;
			R->hits[R->n_hits++] = h;
			free(cg); free(md);
		}
		free(regs.a); free(seq);
	}
	R->hit_off[n] = R->n_hits;
	return R;
}

void oracle_results_view(void *h, b200_results_view_t *v)
{
	ores_t *R = h;
	v->n_reads = R->n_reads; v->hit_off = R->hit_off; v->hits = R->hits; v->cigar = R->cigar; v->md = R->md;
	v->n_hits = R->n_hits; v->n_cigar = R->n_cigar; v->n_md = R->n_md;
}

void oracle_results_free(void *h) { ores_t *R = h; if (!R) return; free(R->hit_off); free(R->hits); free(R->cigar); free(R->md); free(R); }

/* ksw_extend2 batch on the restatement (config-3 tuples) */
void oracle_ksw_extend2_batch(i64 n, const b200_ext_job_t *jobs, const u8 *qp, const u8 *tp, const int8_t *mat, int o_del, int e_del, int o_ins, int e_ins, b200_ext_out_t *out)
{
	i64 i;
	for (i = 0; i < n; ++i) {
		const b200_ext_job_t *x = &jobs[i];
		b200_ext_out_t *o = &out[i];
		o->score = ksw_extend2(x->qlen, qp + x->q_off, x->tlen, tp + x->t_off, mat, o_del, e_del, o_ins, e_ins, x->w, x->end_bonus, x->zdrop, x->h0,
		                       &o->qle, &o->tle, &o->gtle, &o->gscore, &o->max_off);
	}
}
