/*
 * refdrv_bwa.c -- TEST INFRASTRUCTURE ONLY (oracle).
 *
 * Thin driver around the *unmodified* bwa sources of the reference mount
 * (/root/reference/bwa, or $SEQLIB_REF/bwa).  Nothing of the reference is
 * copied here: bwamem.c is pulled in by #include from the mount at compile
 * time (to reach its static mem_collect_intv), the other translation units
 * are compiled from where they lie by oracle/Makefile.  The resulting
 * oracle/_ref/libseqref_bwa.so is the strongest parity checker we have and
 * the "reference" CPU baseline of bench.py.  The product (seqlib_b200/) never
 * links, loads or calls this file.
 *
 * The only restated logic lives in refdrv_index_construct()/refdrv_index_write(),
 * which follow src/BWAIndex.cpp:83-180,183-341,360-406 (that file needs htslib
 * and cannot be compiled here).
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <pthread.h>
#include <time.h>

#include "bwamem.c" /* from -I$(REF)/bwa : the reference's own file, unmodified */
#include "../include/seqlib_b200.h"

extern int is_bwt(ubyte_t *T, int n);

typedef struct {
	bwaidx_t *idx;
	int borrowed; /* arrays belong to the caller */
} refidx_t;

static double now_s(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/* ---------------- index ---------------- */

#define PAC_SET(pac, l, c) ((pac)[(l)>>2] |= (c)<<((~(l)&3)<<1))
#define PAC_GET(pac, l) ((pac)[(l)>>2]>>((~(l)&3)<<1)&3)

/* src/BWAIndex.cpp:183-302 (seqlib_add1 + seqlib_make_pac): 2-bit pack with
 * lrand48()&3 for every non-ACGT base; the second call appends the reverse
 * complement of what it has just packed. */
static uint8_t *make_pac(int n, const char *const *seqs, int for_only, int64_t *l_out)
{
	int64_t tot = 0, l = 0, k;
	int i;
	for (i = 0; i < n; ++i) tot += strlen(seqs[i]);
	uint8_t *pac = calloc((size_t)(2 * tot + 3) / 4 + 8, 1);
	for (i = 0; i < n; ++i) {
		const char *s = seqs[i];
		for (k = 0; s[k]; ++k) {
			int c = nst_nt4_table[(unsigned char)s[k]];
			if (c >= 4) c = lrand48() & 3;
			PAC_SET(pac, l, c);
			++l;
		}
	}
	if (!for_only)
		for (k = l - 1; k >= 0; --k, ++l)
			PAC_SET(pac, l, 3 - PAC_GET(pac, k));
	*l_out = l;
	return pac;
}

void *refdrv_index_construct(int n, const char *const *names, const char *const *seqs)
{
	int64_t l_fwd, l_all, i;
	int j;
	uint8_t *fwd = make_pac(n, seqs, 1, &l_fwd);
	uint8_t *pac = make_pac(n, seqs, 0, &l_all);
	/* src/BWAIndex.cpp:305-341 seqlib_bwt_pac2bwt */
	bwt_t *bwt = calloc(1, sizeof(bwt_t));
	bwt->seq_len = l_all;
	bwt->bwt_size = (bwt->seq_len + 15) >> 4;
	ubyte_t *buf = calloc(bwt->seq_len + 1, 1);
	for (i = 0; i < l_all; ++i) {
		buf[i] = pac[i>>2] >> ((3 - (i&3)) << 1) & 3;
		++bwt->L2[1 + buf[i]];
	}
	for (j = 2; j <= 4; ++j) bwt->L2[j] += bwt->L2[j-1];
	bwt->primary = is_bwt(buf, bwt->seq_len);
	bwt->bwt = calloc(bwt->bwt_size, 4);
	for (i = 0; i < l_all; ++i)
		bwt->bwt[i>>4] |= buf[i] << ((15 - (i&15)) << 1);
	free(buf); free(pac);
	bwt_bwtupdate_core(bwt);
	bwt_cal_sa(bwt, 32);
	bwt_gen_cnt_table(bwt);
	/* src/BWAIndex.cpp:141-174 */
	bntseq_t *bns = calloc(1, sizeof(bntseq_t));
	bns->l_pac = l_fwd; bns->n_seqs = n; bns->seed = 11; bns->n_holes = 0;
	bns->anns = calloc(n, sizeof(bntann1_t));
	int64_t off = 0;
	for (j = 0; j < n; ++j) {
		bntann1_t *a = &bns->anns[j];
		a->offset = off; a->name = strdup(names[j]); a->anno = strdup("(null)");
		a->len = strlen(seqs[j]); off += a->len;
	}
	refidx_t *r = calloc(1, sizeof(refidx_t));
	r->idx = calloc(1, sizeof(bwaidx_t));
	r->idx->bwt = bwt; r->idx->bns = bns; r->idx->pac = fwd;
	return r;
}

void *refdrv_index_load(const char *prefix)
{
	bwaidx_t *idx = bwa_idx_load(prefix, BWA_IDX_ALL);
	if (!idx) return 0;
	refidx_t *r = calloc(1, sizeof(refidx_t));
	r->idx = idx;
	return r;
}

/* wrap arrays built elsewhere (the GPU index builder) in a bwaidx_t so the
 * reference code can run on the same 3 Gb index; arrays are borrowed. */
void *refdrv_index_from_view(const b200_index_view_t *v)
{
	int j;
	refidx_t *r = calloc(1, sizeof(refidx_t));
	r->borrowed = 1;
	r->idx = calloc(1, sizeof(bwaidx_t));
	bwt_t *bwt = calloc(1, sizeof(bwt_t));
	bwt->primary = v->primary; memcpy(bwt->L2, v->L2, sizeof(bwt->L2));
	bwt->seq_len = v->seq_len; bwt->bwt_size = v->bwt_size;
	bwt->bwt = (uint32_t*)v->bwt; bwt->sa_intv = v->sa_intv; bwt->n_sa = v->n_sa;
	bwt->sa = (bwtint_t*)v->sa;
	bwt_gen_cnt_table(bwt);
	bntseq_t *bns = calloc(1, sizeof(bntseq_t));
	bns->l_pac = v->l_pac; bns->n_seqs = v->n_seqs; bns->seed = 11;
	bns->anns = calloc(v->n_seqs, sizeof(bntann1_t));
	for (j = 0; j < v->n_seqs; ++j) {
		bns->anns[j].offset = v->contigs[j].offset; bns->anns[j].len = v->contigs[j].len;
		bns->anns[j].n_ambs = v->contigs[j].n_ambs; bns->anns[j].gi = v->contigs[j].gi;
		bns->anns[j].is_alt = v->contigs[j].is_alt;
		bns->anns[j].name = strdup(v->contigs[j].name);
		bns->anns[j].anno = strdup(v->contigs[j].anno? v->contigs[j].anno : "(null)");
	}
	r->idx->bwt = bwt; r->idx->bns = bns; r->idx->pac = (uint8_t*)v->pac;
	return r;
}

static b200_contig_t *g_view_contigs = 0; /* last view's contig table (test helper, not thread safe) */

void refdrv_index_view(void *h, b200_index_view_t *v)
{
	refidx_t *r = h;
	const bwt_t *b = r->idx->bwt;
	const bntseq_t *bns = r->idx->bns;
	int j;
	v->primary = b->primary; memcpy(v->L2, b->L2, sizeof(v->L2));
	v->seq_len = b->seq_len; v->bwt_size = b->bwt_size; v->bwt = b->bwt;
	v->sa_intv = b->sa_intv; v->n_sa = b->n_sa; v->sa = b->sa;
	v->l_pac = bns->l_pac; v->pac = r->idx->pac; v->n_seqs = bns->n_seqs;
	free(g_view_contigs);
	g_view_contigs = calloc(bns->n_seqs, sizeof(b200_contig_t));
	for (j = 0; j < bns->n_seqs; ++j) {
		g_view_contigs[j].offset = bns->anns[j].offset; g_view_contigs[j].len = bns->anns[j].len;
		g_view_contigs[j].n_ambs = bns->anns[j].n_ambs; g_view_contigs[j].gi = bns->anns[j].gi;
		g_view_contigs[j].is_alt = bns->anns[j].is_alt;
		g_view_contigs[j].name = bns->anns[j].name; g_view_contigs[j].anno = bns->anns[j].anno;
	}
	v->contigs = g_view_contigs;
}

/* src/BWAIndex.cpp:360-406 (WriteIndex + seqlib_write_pac_to_file) */
int refdrv_index_write(void *h, const char *prefix)
{
	refidx_t *r = h;
	char fn[4096];
	snprintf(fn, sizeof fn, "%s.bwt", prefix); bwt_dump_bwt(fn, r->idx->bwt);
	snprintf(fn, sizeof fn, "%s.sa", prefix); bwt_dump_sa(fn, r->idx->bwt);
	bns_dump(r->idx->bns, prefix);
	snprintf(fn, sizeof fn, "%s.pac", prefix);
	FILE *fp = fopen(fn, "wb");
	if (!fp) return -1;
	int64_t l_pac = r->idx->bns->l_pac;
	ubyte_t ct;
	fwrite(r->idx->pac, 1, (l_pac>>2) + ((l_pac&3) == 0? 0 : 1), fp);
	if (l_pac % 4 == 0) { ct = 0; fwrite(&ct, 1, 1, fp); }
	ct = l_pac % 4; fwrite(&ct, 1, 1, fp);
	fclose(fp);
	return 0;
}

void refdrv_index_destroy(void *h)
{
	refidx_t *r = h;
	if (!r) return;
	if (r->borrowed) {
		int j;
		for (j = 0; j < r->idx->bns->n_seqs; ++j) { free(r->idx->bns->anns[j].name); free(r->idx->bns->anns[j].anno); }
		free(r->idx->bns->anns); free(r->idx->bns); free(r->idx->bwt); free(r->idx);
	} else bwa_idx_destroy(r->idx);
	free(r);
}

/* ---------------- options ---------------- */

static mem_opt_t *opt_from(const b200_mem_opt_t *o)
{
	mem_opt_t *opt = mem_opt_init();
	if (o) {
		if (sizeof(mem_opt_t) != sizeof(b200_mem_opt_t)) { fprintf(stderr, "opt layout mismatch\n"); abort(); }
		memcpy(opt, o, sizeof(mem_opt_t));
	} else opt->flag |= MEM_F_SOFTCLIP; /* SeqLib/BWAAligner.h:17 */
	return opt;
}

void refdrv_opt_init(b200_mem_opt_t *o) /* mem_opt_init + SeqLib's MEM_F_SOFTCLIP */
{
	mem_opt_t *opt = mem_opt_init();
	opt->flag |= MEM_F_SOFTCLIP;
	memcpy(o, opt, sizeof(mem_opt_t));
	free(opt);
}

/* ---------------- alignment ---------------- */

typedef struct {
	int64_t n_reads;
	int64_t *hit_off;
	b200_hit_t *hits;
	uint32_t *cigar;
	char *md;
	int64_t n_hits, n_cigar, n_md;
	int64_t m_hits, m_cigar, m_md;
} refres_t;

typedef struct { /* per-read staging so threads do not contend */
	int n; b200_hit_t *h; uint32_t **cig; char **md;
} perread_t;

static void align_one(const mem_opt_t *opt, const bwaidx_t *idx, int l, const char *s, int64_t id, perread_t *pr)
{
	int i;
	char *seq = malloc(l + 1);
	memcpy(seq, s, l); seq[l] = 0;
	char *tmp = malloc(l + 1); memcpy(tmp, seq, l);
	mem_alnreg_v ar = mem_align1_core(opt, idx->bwt, idx->bns, idx->pac, l, tmp, 0);
	mem_mark_primary_se(opt, ar.n, ar.a, id); /* bwa/bwamem_extra.c:112 with the id made explicit */
	free(tmp);
	pr->n = ar.n;
	pr->h = calloc(ar.n? ar.n : 1, sizeof(b200_hit_t));
	pr->cig = calloc(ar.n? ar.n : 1, sizeof(uint32_t*));
	pr->md = calloc(ar.n? ar.n : 1, sizeof(char*));
	for (i = 0; i < (int)ar.n; ++i) {
		mem_alnreg_t *r = &ar.a[i];
		mem_aln_t a = mem_reg2aln(opt, idx->bns, idx->pac, l, seq, r);
		b200_hit_t *h = &pr->h[i];
		h->rb = r->rb; h->re = r->re; h->qb = r->qb; h->qe = r->qe; h->rid = r->rid;
		h->score = r->score; h->truesc = r->truesc; h->sub = r->sub; h->alt_sc = r->alt_sc;
		h->csub = r->csub; h->sub_n = r->sub_n; h->w = r->w; h->seedcov = r->seedcov;
		h->secondary = r->secondary; h->secondary_all = r->secondary_all; h->seedlen0 = r->seedlen0;
		h->n_comp = r->n_comp; h->is_alt = r->is_alt; h->frac_rep = r->frac_rep; h->hash = r->hash;
		h->pos = a.pos; h->flag = a.flag; h->is_rev = a.is_rev; h->mapq = a.mapq; h->NM = a.NM;
		h->aln_sub = a.sub; h->n_cigar = a.n_cigar;
		pr->cig[i] = a.cigar;
		if (a.cigar) { pr->md[i] = (char*)(a.cigar + a.n_cigar); h->md_len = strlen(pr->md[i]); }
		if (a.rid != r->rid) h->rid = -1000; /* cannot happen (assert in the reference) */
	}
	free(ar.a); free(seq);
}

typedef struct {
	const mem_opt_t *opt; const bwaidx_t *idx; const char *seqs; const int64_t *off; const int64_t *ids;
	perread_t *pr; int64_t n; int64_t next; int chunk; pthread_mutex_t mu;
} job_t;

static void *worker(void *d)
{
	job_t *j = d;
	for (;;) {
		pthread_mutex_lock(&j->mu);
		int64_t b = j->next; j->next += j->chunk;
		pthread_mutex_unlock(&j->mu);
		if (b >= j->n) break;
		int64_t e = b + j->chunk < j->n? b + j->chunk : j->n, i;
		for (i = b; i < e; ++i)
			align_one(j->opt, j->idx, (int)(j->off[i+1] - j->off[i]), j->seqs + j->off[i], j->ids[i], &j->pr[i]);
	}
	return 0;
}

/* mem_align1 + mem_reg2aln per read (what BWAAligner::alignSequence runs,
 * src/BWAAligner.cpp:104-128).  ids == NULL: lrand48() per read in order. */
void *refdrv_align(void *h, const b200_mem_opt_t *o, int64_t n, const char *seqs, const int64_t *off,
                   const int64_t *ids, int n_threads, double *seconds)
{
	refidx_t *r = h;
	mem_opt_t *opt = opt_from(o);
	int64_t i, *myids = 0;
	int t;
	if (!ids) {
		myids = malloc(sizeof(int64_t) * (n? n : 1));
		for (i = 0; i < n; ++i) myids[i] = lrand48();
		ids = myids;
	}
	perread_t *pr = calloc(n? n : 1, sizeof(perread_t));
	job_t j = { opt, r->idx, seqs, off, ids, pr, n, 0, 64 };
	pthread_mutex_init(&j.mu, 0);
	double t0 = now_s();
	if (n_threads <= 1) worker(&j);
	else {
		pthread_t *th = calloc(n_threads, sizeof(pthread_t));
		for (t = 0; t < n_threads; ++t) pthread_create(&th[t], 0, worker, &j);
		for (t = 0; t < n_threads; ++t) pthread_join(th[t], 0);
		free(th);
	}
	if (seconds) *seconds = now_s() - t0;
	refres_t *res = calloc(1, sizeof(refres_t));
	res->n_reads = n;
	res->hit_off = calloc(n + 1, sizeof(int64_t));
	for (i = 0; i < n; ++i) res->hit_off[i+1] = res->hit_off[i] + pr[i].n;
	res->n_hits = res->hit_off[n];
	res->hits = calloc(res->n_hits? res->n_hits : 1, sizeof(b200_hit_t));
	for (i = 0; i < n; ++i) {
		int k;
		for (k = 0; k < pr[i].n; ++k) { res->n_cigar += pr[i].h[k].n_cigar; res->n_md += pr[i].h[k].md_len + 1; }
	}
	res->cigar = calloc(res->n_cigar? res->n_cigar : 1, 4);
	res->md = calloc(res->n_md? res->n_md : 1, 1);
	int64_t pc = 0, pm = 0;
	for (i = 0; i < n; ++i) {
		int k;
		for (k = 0; k < pr[i].n; ++k) {
			b200_hit_t *hh = &res->hits[res->hit_off[i] + k];
			*hh = pr[i].h[k];
			hh->cigar_off = pc; hh->md_off = pm;
			if (pr[i].cig[k]) {
				memcpy(res->cigar + pc, pr[i].cig[k], 4 * hh->n_cigar); pc += hh->n_cigar;
				memcpy(res->md + pm, pr[i].md[k], hh->md_len + 1); pm += hh->md_len + 1;
				free(pr[i].cig[k]);
			} else { res->md[pm++] = 0; }
		}
		free(pr[i].h); free(pr[i].cig); free(pr[i].md);
	}
	res->n_md = pm;
	free(pr); free(myids); free(opt);
	return res;
}

void refdrv_results_view(void *h, b200_results_view_t *v)
{
	refres_t *r = h;
	v->n_reads = r->n_reads; v->hit_off = r->hit_off; v->hits = r->hits; v->cigar = r->cigar; v->md = r->md;
	v->n_hits = r->n_hits; v->n_cigar = r->n_cigar; v->n_md = r->n_md;
}

void refdrv_results_free(void *h)
{
	refres_t *r = h;
	if (!r) return;
	free(r->hit_off); free(r->hits); free(r->cigar); free(r->md); free(r);
}

/* The reference's own batched schedule (bwa/bwamem.c:1235-1264): kt_for over
 * mem_align1_core then mark-primary + SAM text, opt->n_threads threads.
 * Timing baseline only (tie-break id = read index).  Returns seconds. */
double refdrv_process_seqs(void *h, const b200_mem_opt_t *o, int64_t n, const char *seqs, const int64_t *off, int n_threads)
{
	refidx_t *r = h;
	mem_opt_t *opt = opt_from(o);
	int64_t i;
	opt->n_threads = n_threads;
	bseq1_t *bs = calloc(n? n : 1, sizeof(bseq1_t));
	for (i = 0; i < n; ++i) {
		int l = off[i+1] - off[i];
		bs[i].l_seq = l; bs[i].id = i;
		bs[i].seq = malloc(l + 1); memcpy(bs[i].seq, seqs + off[i], l); bs[i].seq[l] = 0;
		bs[i].name = strdup("r"); bs[i].comment = 0; bs[i].qual = 0; bs[i].sam = 0;
	}
	int save = bwa_verbose; bwa_verbose = 1;
	double t0 = now_s();
	mem_process_seqs(opt, r->idx->bwt, r->idx->bns, r->idx->pac, 0, n, bs, 0);
	double dt = now_s() - t0;
	bwa_verbose = save;
	for (i = 0; i < n; ++i) { free(bs[i].seq); free(bs[i].name); free(bs[i].sam); }
	free(bs); free(opt);
	return dt;
}

/* ---------------- stage dumps ---------------- */

/* mem_collect_intv (bwa/bwamem.c:140-188) per read */
int refdrv_collect_intv(void *h, const b200_mem_opt_t *o, int64_t n, const char *seqs, const int64_t *off,
                        int64_t **intv_off, b200_intv_t **intv)
{
	refidx_t *r = h;
	mem_opt_t *opt = opt_from(o);
	int64_t i, m = 1024, tot = 0;
	size_t k;
	b200_intv_t *out = malloc(m * sizeof(b200_intv_t));
	int64_t *ooff = calloc(n + 1, sizeof(int64_t));
	smem_aux_t *aux = smem_aux_init();
	for (i = 0; i < n; ++i) {
		int l = off[i+1] - off[i], q;
		uint8_t *s = malloc(l + 1);
		for (q = 0; q < l; ++q) s[q] = nst_nt4_table[(unsigned char)seqs[off[i] + q]];
		ooff[i] = tot;
		if (l >= opt->min_seed_len) { /* mem_chain returns before collecting otherwise, bwamem.c:286 */
			mem_collect_intv(opt, r->idx->bwt, l, s, aux);
			for (k = 0; k < aux->mem.n; ++k) {
				if (tot == m) { m <<= 1; out = realloc(out, m * sizeof(b200_intv_t)); }
				out[tot].x0 = aux->mem.a[k].x[0]; out[tot].x1 = aux->mem.a[k].x[1];
				out[tot].x2 = aux->mem.a[k].x[2]; out[tot].info = aux->mem.a[k].info; ++tot;
			}
		}
		free(s);
	}
	ooff[n] = tot;
	smem_aux_destroy(aux); free(opt);
	*intv_off = ooff; *intv = out;
	return 0;
}

/* chains after mem_chain + mem_chain_flt (bwa/bwamem.c:277-411), flattened:
 * per chain 6 int64 {pos, rid, weight, kept, n_seeds, first}, per seed 4 int64 {rbeg,qbeg,len,score} */
int refdrv_chains(void *h, const b200_mem_opt_t *o, int64_t n, const char *seqs, const int64_t *off,
                  int64_t **chn_off, int64_t **chn, int64_t **seed_off, int64_t **seeds)
{
	refidx_t *r = h;
	mem_opt_t *opt = opt_from(o);
	int64_t i, mc = 1024, ms = 4096, nc = 0, ns = 0;
	int64_t *C = malloc(mc * 6 * 8), *S = malloc(ms * 4 * 8);
	int64_t *coff = calloc(n + 1, 8), *soff = malloc(mc * 8 + 8);
	for (i = 0; i < n; ++i) {
		int l = off[i+1] - off[i], q, k, j;
		uint8_t *s = malloc(l + 1);
		for (q = 0; q < l; ++q) s[q] = nst_nt4_table[(unsigned char)seqs[off[i] + q]];
		coff[i] = nc;
		mem_chain_v chn_v = mem_chain(opt, r->idx->bwt, r->idx->bns, l, s, 0);
		chn_v.n = mem_chain_flt(opt, chn_v.n, chn_v.a);
		for (k = 0; k < (int)chn_v.n; ++k) {
			mem_chain_t *c = &chn_v.a[k];
			if (nc == mc) { mc <<= 1; C = realloc(C, mc * 6 * 8); soff = realloc(soff, mc * 8 + 8); }
			C[nc*6+0] = c->pos; C[nc*6+1] = c->rid; C[nc*6+2] = c->w; C[nc*6+3] = c->kept; C[nc*6+4] = c->n; C[nc*6+5] = c->first;
			soff[nc] = ns;
			for (j = 0; j < c->n; ++j) {
				if (ns == ms) { ms <<= 1; S = realloc(S, ms * 4 * 8); }
				S[ns*4+0] = c->seeds[j].rbeg; S[ns*4+1] = c->seeds[j].qbeg; S[ns*4+2] = c->seeds[j].len; S[ns*4+3] = c->seeds[j].score; ++ns;
			}
			++nc;
			free(c->seeds);
		}
		free(chn_v.a); free(s);
	}
	coff[n] = nc; soff[nc] = ns;
	free(opt);
	*chn_off = coff; *chn = C; *seed_off = soff; *seeds = S;
	return 0;
}

/* regions straight out of mem_chain2aln, before mem_sort_dedup_patch (bwa/bwamem.c:1095-1102) */
int refdrv_regs_raw(void *h, const b200_mem_opt_t *o, int64_t n, const char *seqs, const int64_t *off,
                    int64_t **reg_off, b200_hit_t **regs)
{
	refidx_t *r = h;
	mem_opt_t *opt = opt_from(o);
	int64_t i, m = 1024, tot = 0;
	b200_hit_t *R = calloc(m, sizeof(b200_hit_t));
	int64_t *roff = calloc(n + 1, 8);
	for (i = 0; i < n; ++i) {
		int l = off[i+1] - off[i], q, k;
		uint8_t *s = malloc(l + 1);
		for (q = 0; q < l; ++q) s[q] = nst_nt4_table[(unsigned char)seqs[off[i] + q]];
		roff[i] = tot;
		mem_chain_v chn_v = mem_chain(opt, r->idx->bwt, r->idx->bns, l, s, 0);
		chn_v.n = mem_chain_flt(opt, chn_v.n, chn_v.a);
		mem_flt_chained_seeds(opt, r->idx->bns, r->idx->pac, l, s, chn_v.n, chn_v.a);
		mem_alnreg_v regs_v; kv_init(regs_v);
		for (k = 0; k < (int)chn_v.n; ++k) {
			mem_chain2aln(opt, r->idx->bns, r->idx->pac, l, s, &chn_v.a[k], &regs_v);
			free(chn_v.a[k].seeds);
		}
		free(chn_v.a);
		for (k = 0; k < (int)regs_v.n; ++k) {
			mem_alnreg_t *a = &regs_v.a[k];
			if (tot == m) { m <<= 1; R = realloc(R, m * sizeof(b200_hit_t)); memset(R + tot, 0, (m - tot) * sizeof(b200_hit_t)); }
			b200_hit_t *hh = &R[tot++];
			hh->rb = a->rb; hh->re = a->re; hh->qb = a->qb; hh->qe = a->qe; hh->rid = a->rid;
			hh->score = a->score; hh->truesc = a->truesc; hh->w = a->w; hh->seedcov = a->seedcov;
			hh->seedlen0 = a->seedlen0; hh->frac_rep = a->frac_rep;
		}
		free(regs_v.a); free(s);
	}
	roff[n] = tot;
	free(opt);
	*reg_off = roff; *regs = R;
	return 0;
}

/* ---------------- ksw_extend2 batch ---------------- */

typedef struct {
	int64_t n, next; const b200_ext_job_t *jobs; const uint8_t *qp, *tp; const int8_t *mat;
	int o_del, e_del, o_ins, e_ins; b200_ext_out_t *out; pthread_mutex_t mu;
} extjob_t;

static void *ext_worker(void *d)
{
	extjob_t *j = d;
	for (;;) {
		pthread_mutex_lock(&j->mu);
		int64_t b = j->next; j->next += 256;
		pthread_mutex_unlock(&j->mu);
		if (b >= j->n) break;
		int64_t e = b + 256 < j->n? b + 256 : j->n, i;
		for (i = b; i < e; ++i) {
			const b200_ext_job_t *x = &j->jobs[i];
			b200_ext_out_t *o = &j->out[i];
			o->score = ksw_extend2(x->qlen, j->qp + x->q_off, x->tlen, j->tp + x->t_off, 5, j->mat,
			                       j->o_del, j->e_del, j->o_ins, j->e_ins, x->w, x->end_bonus, x->zdrop, x->h0,
			                       &o->qle, &o->tle, &o->gtle, &o->gscore, &o->max_off);
		}
	}
	return 0;
}

double refdrv_ksw_extend2_batch(int64_t n, const b200_ext_job_t *jobs, const uint8_t *qpool, const uint8_t *tpool,
                                const int8_t mat[25], int o_del, int e_del, int o_ins, int e_ins,
                                b200_ext_out_t *out, int n_threads)
{
	extjob_t j = { n, 0, jobs, qpool, tpool, mat, o_del, e_del, o_ins, e_ins, out };
	int t;
	pthread_mutex_init(&j.mu, 0);
	double t0 = now_s();
	if (n_threads <= 1) ext_worker(&j);
	else {
		pthread_t *th = calloc(n_threads, sizeof(pthread_t));
		for (t = 0; t < n_threads; ++t) pthread_create(&th[t], 0, ext_worker, &j);
		for (t = 0; t < n_threads; ++t) pthread_join(th[t], 0);
		free(th);
	}
	return now_s() - t0;
}

void refdrv_free(void *p) { free(p); }
void refdrv_srand48(long seed) { srand48(seed); }
long refdrv_lrand48(void) { return lrand48(); }

/* SAM text of single-end reads: per read mem_align1 (with the tie-break id made explicit) then mem_reg2sam
 * (bwa/bwamem.c:1034-1086), i.e. mem_gen_alt (bwa/bwamem_extra.c:125-173) + mem_aln2sam (bwa/bwamem.c:851-976).
 * quals / comments may be NULL.  *sam is malloc'd (refdrv_free). */
int refdrv_sam(void *h, const b200_mem_opt_t *o, int64_t n, const char *seqs, const int64_t *off, const int64_t *ids,
               const char *names, const int64_t *name_off, const char *quals, const int64_t *qual_off,
               const char *comments, const int64_t *comment_off, char **sam, int64_t *sam_len)
{
	refidx_t *r = h;
	mem_opt_t *opt = opt_from(o);
	int64_t i, cap = 1 << 16, len = 0;
	char *out = malloc(cap);
	for (i = 0; i < n; ++i) {
		int l = (int)(off[i+1] - off[i]), k;
		bseq1_t s; memset(&s, 0, sizeof(s));
		s.l_seq = l;
		s.seq = malloc(l + 1); memcpy(s.seq, seqs + off[i], l); s.seq[l] = 0;
		s.name = strndup(names + name_off[i], name_off[i+1] - name_off[i]);
		if (quals && qual_off[i+1] > qual_off[i]) s.qual = strndup(quals + qual_off[i], qual_off[i+1] - qual_off[i]);
		if (comments && comment_off[i+1] > comment_off[i]) s.comment = strndup(comments + comment_off[i], comment_off[i+1] - comment_off[i]);
		mem_alnreg_v ar = mem_align1_core(opt, r->idx->bwt, r->idx->bns, r->idx->pac, l, s.seq, 0); /* converts s.seq to codes in place */
		mem_mark_primary_se(opt, ar.n, ar.a, ids? ids[i] : lrand48());
		mem_reg2sam(opt, r->idx->bns, r->idx->pac, &s, &ar, 0, 0);
		k = strlen(s.sam);
		if (len + k + 1 > cap) { while (len + k + 1 > cap) cap *= 2; out = realloc(out, cap); }
		memcpy(out + len, s.sam, k); len += k;
		free(s.sam); free(s.seq); free(s.name); free(s.qual); free(s.comment); free(ar.a);
	}
	out[len] = 0;
	*sam = out; *sam_len = len;
	free(opt);
	return 0;
}
