/*
 * refdrv_fml.c -- TEST INFRASTRUCTURE ONLY (oracle).
 *
 * Driver around the unmodified fermi-lite sources of the reference mount
 * ($(REF)/fermi-lite, compiled in place by oracle/Makefile).  Exposes
 * fml_correct / fml_fltuniq / fml_assemble on flat buffers so tests and the
 * CPU baseline of bench.py can call them through ctypes.  Never linked into
 * the product.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <time.h>
#include "fml.h"
#include "htab.h"
#include "kmer.h"
#include "rld0.h"
#include "mag.h"
#include "kstring.h"

extern unsigned char seq_nt6_table[256];
struct bfc_ch_s *fml_count(int n, const fseq1_t *seq, int k, int q, int l_pre, int n_threads);

static double now_s(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static fseq1_t *mk_seqs(int n, const char *seqs, const char *quals, const int64_t *off)
{
	int i;
	fseq1_t *s = calloc(n? n : 1, sizeof(fseq1_t));
	for (i = 0; i < n; ++i) {
		int l = off[i+1] - off[i];
		s[i].l_seq = l;
		s[i].seq = malloc(l + 1); memcpy(s[i].seq, seqs + off[i], l); s[i].seq[l] = 0;
		if (quals) { s[i].qual = malloc(l + 1); memcpy(s[i].qual, quals + off[i], l); s[i].qual[l] = 0; }
	}
	return s;
}

void refdrv_fml_opt_init(fml_opt_t *opt) { fml_opt_init(opt); }
int refdrv_fml_opt_size(void) { return sizeof(fml_opt_t); }

/* fml_correct (fermi-lite/bfc.c:568) after optional fml_opt_adjust; reads are
 * rewritten in place into seqs_out/quals_out (same offsets). flt_uniq=1 runs
 * fml_fltuniq instead (lengths may shrink: new lengths in len_out, 0 = dropped). */
float refdrv_fml_correct(const fml_opt_t *opt0, int adjust, int flt_uniq, int n, const char *seqs, const char *quals,
                         const int64_t *off, char *seqs_out, char *quals_out, int32_t *len_out, double *seconds)
{
	fml_opt_t opt = *opt0;
	fseq1_t *s = mk_seqs(n, seqs, quals, off);
	int i;
	float kcov;
	if (adjust) fml_opt_adjust(&opt, n, s);
	double t0 = now_s();
	kcov = flt_uniq? fml_fltuniq(&opt, n, s) : fml_correct(&opt, n, s);
	if (seconds) *seconds = now_s() - t0;
	for (i = 0; i < n; ++i) {
		len_out[i] = s[i].l_seq;
		if (s[i].l_seq > 0 && s[i].seq) {
			memcpy(seqs_out + off[i], s[i].seq, s[i].l_seq);
			if (s[i].qual && quals_out) memcpy(quals_out + off[i], s[i].qual, s[i].l_seq);
			free(s[i].seq); free(s[i].qual);
		}
	}
	free(s);
	return kcov;
}

/* fml_assemble (fermi-lite/misc.c:280-302).  Returns the unitigs flattened:
 * utg_off[n_utg+1] into one char pool (sequence) and one cov pool, nsr[]. */
int refdrv_fml_assemble(const fml_opt_t *opt, int n, const char *seqs, const char *quals, const int64_t *off,
                        int64_t **utg_off, char **utg_seq, char **utg_cov, int32_t **utg_nsr,
                        int32_t **utg_novlp /* 2 per unitig */, int32_t **utg_ovlp /* len, from, id, to per overlap */, double *seconds)
{
	fseq1_t *s = mk_seqs(n, seqs, quals, off);
	int n_utg = 0, i;
	double t0 = now_s();
	fml_utg_t *u = fml_assemble(opt, n, s, &n_utg);
	if (seconds) *seconds = now_s() - t0;
	int64_t tot = 0;
	int64_t *uo = calloc(n_utg + 1, 8);
	for (i = 0; i < n_utg; ++i) { uo[i] = tot; tot += u[i].len; }
	uo[n_utg] = tot;
	char *us = malloc(tot + 1), *uc = malloc(tot + 1);
	int32_t *nsr = calloc(n_utg? n_utg : 1, 4);
	for (i = 0; i < n_utg; ++i) {
		memcpy(us + uo[i], u[i].seq, u[i].len);
		memcpy(uc + uo[i], u[i].cov, u[i].len);
		nsr[i] = u[i].nsr;
	}
	{
		int64_t no = 0, k = 0; int j;
		int32_t *nv = calloc(n_utg ? 2 * n_utg : 1, 4), *ov;
		for (i = 0; i < n_utg; ++i) { nv[2*i] = u[i].n_ovlp[0]; nv[2*i+1] = u[i].n_ovlp[1]; no += u[i].n_ovlp[0] + u[i].n_ovlp[1]; }
		ov = calloc(no ? 4 * no : 1, 4);
		for (i = 0; i < n_utg; ++i)
			for (j = 0; j < u[i].n_ovlp[0] + u[i].n_ovlp[1]; ++j, ++k) {
				ov[4*k] = u[i].ovlp[j].len; ov[4*k+1] = u[i].ovlp[j].from; ov[4*k+2] = u[i].ovlp[j].id; ov[4*k+3] = u[i].ovlp[j].to;
			}
		*utg_novlp = nv; *utg_ovlp = ov;
	}
	fml_utg_destroy(n_utg, u);
	/* in this fork fml_assemble does not free the reads (fermi-lite/misc.c:85-102) */
	for (i = 0; i < n; ++i) { if (s[i].l_seq > 0) { free(s[i].seq); free(s[i].qual); } }
	free(s);
	*utg_off = uo; *utg_seq = us; *utg_cov = uc; *utg_nsr = nsr;
	return n_utg;
}

static void free_seqs(int n, fseq1_t *s)
{
	int i;
	for (i = 0; i < n; ++i) { free(s[i].seq); free(s[i].qual); }
	free(s);
}

/* fml_count (fermi-lite/bfc.c:86-99) + bfc_ch_hist (fermi-lite/htab.c:104-127): hist[0..255] totals,
 * hist[256..319] high-quality counts; returns the mode. */
int refdrv_fml_count_hist(int n, const char *seqs, const char *quals, const int64_t *off, int k, int q, int l_pre,
                          uint64_t *hist, int64_t *n_distinct)
{
	fseq1_t *s = mk_seqs(n, seqs, quals, off);
	bfc_ch_t *ch = fml_count(n, s, k, q, l_pre, 1);
	int mode = bfc_ch_hist(ch, hist, hist + 256);
	if (n_distinct) *n_distinct = (int64_t)bfc_ch_count(ch);
	bfc_ch_destroy(ch);
	free_seqs(n, s);
	return mode;
}

/* bfc_ch_kmer_occ (fermi-lite/htab.c:85-93) for n_q ASCII k-mers against the table of the given reads. */
void refdrv_fml_kmer_occ(int n, const char *seqs, const char *quals, const int64_t *off, int k, int q, int l_pre,
                         int64_t n_q, const char *kmers, int32_t *occ)
{
	fseq1_t *s = mk_seqs(n, seqs, quals, off);
	bfc_ch_t *ch = fml_count(n, s, k, q, l_pre, 1);
	int64_t i; int j;
	for (i = 0; i < n_q; ++i) {
		bfc_kmer_t x = {{0, 0, 0, 0}};
		int ok = 1;
		for (j = 0; j < k; ++j) {
			int c = seq_nt6_table[(uint8_t)kmers[i * k + j]] - 1;
			if (c > 3) { ok = 0; break; }
			bfc_kmer_append(k, x.x, c);
		}
		occ[i] = ok ? bfc_ch_kmer_occ(ch, &x) : -1;
	}
	bfc_ch_destroy(ch);
	free_seqs(n, s);
}

/* fml_seq2fmi (fermi-lite/misc.c:65-128) on the given reads, then the whole BWT decoded run by run (rld_dec):
 * *bwt = malloc'd array of symbols 0..5 ($ACGTN), cnt[0..6] = e->cnt (cumulative), returns the BWT length.
 * n_q rank queries: for each q[i], sym[i] = rld_rank1a(e, q[i], &ranks[6*i]). */
int64_t refdrv_fml_bwt(int n, const char *seqs, const int64_t *off, uint8_t **bwt, uint64_t *cnt, uint64_t *mcnt,
                       int64_t n_q, const uint64_t *q, uint64_t *ranks, int32_t *sym)
{
	fml_opt_t opt;
	fseq1_t *s = mk_seqs(n, seqs, 0, off);
	rld_t *e;
	rlditr_t itr;
	int64_t tot, l, z = 0, i;
	int c = 0;
	fml_opt_init(&opt);
	e = fml_seq2fmi(&opt, n, s);
	free_seqs(n, s);
	*bwt = 0;
	if (e == 0) return 0;
	tot = e->mcnt[0];
	*bwt = malloc(tot + 1);
	rld_itr_init(e, &itr, 0);
	while ((l = rld_dec(e, &itr, &c, 0)) > 0) { memset(*bwt + z, c, l); z += l; }
	for (i = 0; i <= 6; ++i) cnt[i] = e->cnt[i], mcnt[i] = e->mcnt[i];
	for (i = 0; i < n_q; ++i) sym[i] = rld_rank1a(e, q[i], ranks + 6 * i);
	rld_destroy(e);
	return z;
}

static char *mag_text(const mag_t *g, int64_t *len)
{
	kstring_t out = {0, 0, 0}, all = {0, 0, 0};
	size_t i;
	for (i = 0; i < g->v.n; ++i) {
		if (g->v.a[i].len < 0) continue;
		mag_v_write(&g->v.a[i], &out);
		if (g->v.a[i].len > 0) kputsn(out.s, out.l, &all);
	}
	free(out.s);
	if (all.s == 0) { all.s = calloc(1, 1); }
	*len = all.l;
	return all.s;
}

/* The assembly half of fml_assemble (fermi-lite/misc.c:291-300) on reads that are already corrected / filtered:
 * fml_seq2fmi, fml_fmi2mag, [fml_mag_clean with min_ensr/min_insr derived from kcov], dumped in mag_g_print's text
 * format (fermi-lite/mag.c:151-176).  stage 0: graph straight out of fml_fmi2mag; stage 1: after fml_mag_clean. */
char *refdrv_fml_mag_text(const fml_opt_t *opt0, int stage, float kcov, int n, const char *seqs, const int64_t *off,
                          int64_t *text_len, float *rdist, double *seconds)
{
	fml_opt_t opt = *opt0;
	fseq1_t *s = mk_seqs(n, seqs, 0, off);
	rld_t *e;
	mag_t *g;
	char *txt;
	double t0 = now_s();
	fml_opt_adjust(&opt, n, s);
	e = fml_seq2fmi(&opt, n, s);
	free_seqs(n, s);
	*text_len = 0;
	if (e == 0) return 0;
	g = fml_fmi2mag(&opt, e);
	if (rdist) *rdist = g->rdist;
	if (stage >= 100) { /* debugging aid: the first (stage - 100) passes of fml_mag_clean / mag_g_clean (fermi-lite/mag.c:559-583) */
		extern int mag_g_rm_vint(mag_t *g, int min_len, int min_nsr, int min_ovlp);
		magopt_t o_; const magopt_t *o = &o_;
		int left = stage - 100, j;
		opt.mag_opt.min_ensr = opt.mag_opt.min_ensr > kcov * .1 ? opt.mag_opt.min_ensr : (int)(kcov * .1 + .499);
		opt.mag_opt.min_ensr = opt.mag_opt.min_ensr < opt0->max_cnt ? opt.mag_opt.min_ensr : opt0->max_cnt;
		opt.mag_opt.min_ensr = opt.mag_opt.min_ensr > opt0->min_cnt ? opt.mag_opt.min_ensr : opt0->min_cnt;
		opt.mag_opt.min_insr = opt.mag_opt.min_ensr - 1;
		o_ = opt.mag_opt; o_.min_merge_len = opt.min_merge_len;
#define STEP(x) do { if (left-- > 0) { x; } } while (0)
		STEP(mag_g_merge(g, 1, opt.min_merge_len));
		for (j = 2; j <= o->min_ensr; ++j) STEP(mag_g_rm_vext(g, o->min_elen, j));
		STEP(mag_g_merge(g, 0, o->min_merge_len));
		STEP(mag_g_rm_edge(g, g->min_ovlp, o->min_dratio1, o->min_elen, o->min_ensr));
		STEP(mag_g_merge(g, 1, o->min_merge_len));
		for (j = 2; j <= o->min_ensr; ++j) STEP(mag_g_rm_vext(g, o->min_elen, j));
		STEP(mag_g_merge(g, 0, o->min_merge_len));
		STEP(mag_g_pop_open(g, o->min_elen));
		STEP(mag_g_pop_simple(g, o->max_bcov, o->max_bfrac, o->min_merge_len, o->max_bdiff, 0));
		STEP(mag_g_rm_vint(g, o->min_elen, o->min_insr, g->min_ovlp));
		STEP(mag_g_rm_edge(g, g->min_ovlp, o->min_dratio1, o->min_elen, o->min_ensr));
		STEP(mag_g_merge(g, 1, o->min_merge_len));
		STEP(mag_g_rm_vext(g, o->min_elen, o->min_ensr));
		STEP(mag_g_merge(g, 0, o->min_merge_len));
		STEP(mag_g_pop_open(g, o->min_elen));
		STEP(mag_g_rm_vext(g, o->min_elen, o->min_ensr));
		STEP(mag_g_merge(g, 0, o->min_merge_len));
#undef STEP
	} else if (stage >= 1) {
		opt.mag_opt.min_ensr = opt.mag_opt.min_ensr > kcov * .1 ? opt.mag_opt.min_ensr : (int)(kcov * .1 + .499);
		opt.mag_opt.min_ensr = opt.mag_opt.min_ensr < opt0->max_cnt ? opt.mag_opt.min_ensr : opt0->max_cnt;
		opt.mag_opt.min_ensr = opt.mag_opt.min_ensr > opt0->min_cnt ? opt.mag_opt.min_ensr : opt0->min_cnt;
		opt.mag_opt.min_insr = opt.mag_opt.min_ensr - 1;
		fml_mag_clean(&opt, g);
	}
	if (seconds) *seconds = now_s() - t0;
	txt = mag_text(g, text_len);
	fml_mag_destroy(g);
	return txt;
}

void refdrv_fml_free(void *p) { free(p); }
