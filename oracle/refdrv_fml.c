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
                        int64_t **utg_off, char **utg_seq, char **utg_cov, int32_t **utg_nsr, double *seconds)
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

void refdrv_fml_free(void *p) { free(p); }
