// seedsw.cuh -- the seed-SW filter of long reads: mem_flt_chained_seeds / mem_seed_sw (bwa/bwamem.c:597-641).
//
// Only active when 5.5 ln(l_query) <= 0.05 l_query (reads longer than ~730 bp, i.e. contigs): every seed shorter than
// MEM_SHORT_LEN is re-scored by a local alignment of the seed +- 50 bp (ksw_align2 -> ksw_i16, bwa/ksw.c:255-343) and
// dropped from its chain when the score stays below min_HSP_score.  The score also orders the seeds in mem_chain2aln, so
// it has to be the number the reference computes: the striped SSE2 kernel (8 x int16 lanes, lazy-F loop, E updated from the
// pre-lazy H) is replayed lane by lane instead of being replaced by a textbook recurrence.
#pragma once
#include "fmindex.cuh"

namespace b200 {

enum { MEM_SHORT_EXT = 50, MEM_SHORT_LEN = 200, KSW_I16_MAX_SLEN = (MEM_SHORT_LEN + 7) / 8 };

struct V8s { i16 v[8]; };

HD i16 sat_add16(int a, int b) { int x = a + b; return (i16)(x > 32767 ? 32767 : x < -32768 ? -32768 : x); }
HD i16 subs_u16(i16 a, int b) { unsigned x = (unsigned)(u16)a; return (i16)(x > (unsigned)b ? x - (unsigned)b : 0u); }

// ksw_i16 with xtra = KSW_XSTART (no early stop, no second-best bookkeeping that could change the score); returns r.score.
// query: nt4 codes (0..4); the target is text[rb, rb + tlen) of the doubled reference.
HD int ksw_i16_score(const DevIndex &ix, int qlen, const u8 *query, int tlen, i64 rb, const i8 *mat,
                     int o_del, int e_del, int o_ins, int e_ins)
{
    const int slen = (qlen + 7) >> 3;
    if (slen <= 0 || slen > KSW_I16_MAX_SLEN) return 0;
    V8s qp[5 * KSW_I16_MAX_SLEN], H0[KSW_I16_MAX_SLEN], H1[KSW_I16_MAX_SLEN], E[KSW_I16_MAX_SLEN];
    for (int a = 0; a < 5; ++a)                      // ksw_qinit, size 2 (bwa/ksw.c:100-108)
        for (int i = 0; i < slen; ++i)
            for (int k = i, lane = 0; lane < 8; k += slen, ++lane)
                qp[a * slen + i].v[lane] = (i16)(k >= qlen ? 0 : mat[a * 5 + query[k]]);
    for (int i = 0; i < slen; ++i)
        for (int l = 0; l < 8; ++l) { H0[i].v[l] = 0; E[i].v[l] = 0; H1[i].v[l] = 0; }
    V8s *h0 = H0, *h1 = H1;
    const int oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
    int gmax = 0;
    for (int i = 0; i < tlen; ++i) {
        const V8s *S = qp + text_base(ix, rb + i) * slen;
        V8s h, f, mx;
        for (int l = 0; l < 8; ++l) { f.v[l] = 0; mx.v[l] = 0; }
        h.v[0] = 0;
        for (int l = 1; l < 8; ++l) h.v[l] = h0[slen - 1].v[l - 1];
        for (int j = 0; j < slen; ++j) {
            V8s e = E[j];
            for (int l = 0; l < 8; ++l) {
                i16 x = sat_add16(h.v[l], S[j].v[l]);
                if (x < e.v[l]) x = e.v[l];
                if (x < f.v[l]) x = f.v[l];
                if (mx.v[l] < x) mx.v[l] = x;
                h1[j].v[l] = x;
                i16 ee = subs_u16(e.v[l], e_del), t = subs_u16(x, oe_del);
                E[j].v[l] = ee > t ? ee : t;
                i16 ff = subs_u16(f.v[l], e_ins); t = subs_u16(x, oe_ins);
                f.v[l] = ff > t ? ff : t;
                h.v[l] = h0[j].v[l];
            }
        }
        for (int k = 0; k < 16; ++k) {
            bool done = false;
            for (int l = 7; l > 0; --l) f.v[l] = f.v[l - 1];
            f.v[0] = 0;
            for (int j = 0; j < slen; ++j) {
                bool any = false;
                for (int l = 0; l < 8; ++l) {
                    i16 x = h1[j].v[l];
                    if (x < f.v[l]) x = f.v[l];
                    h1[j].v[l] = x;
                    x = subs_u16(x, oe_ins);
                    f.v[l] = subs_u16(f.v[l], e_ins);
                    if (f.v[l] > x) any = true;
                }
                if (!any) { done = true; break; }
            }
            if (done) break;
        }
        int imax = mx.v[0];
        for (int l = 1; l < 8; ++l) if (mx.v[l] > imax) imax = mx.v[l];
        if (imax > gmax) gmax = imax;
        V8s *t2 = h0; h0 = h1; h1 = t2;
    }
    return gmax;
}

// mem_seed_sw (bwa/bwamem.c:597-622)
HD int seed_sw(const DevIndex &ix, const Opt &opt, int l_query, const u8 *query, i64 s_rbeg, int s_qbeg, int s_len)
{
    if (s_len >= MEM_SHORT_LEN) return -1;
    const i64 l_pac = ix.l_pac;
    int qb = s_qbeg, qe = s_qbeg + s_len;
    i64 rb = s_rbeg, re = s_rbeg + s_len, mid = (rb + re) >> 1;
    qb -= MEM_SHORT_EXT; qb = qb > 0 ? qb : 0;
    qe += MEM_SHORT_EXT; qe = qe < l_query ? qe : l_query;
    rb -= MEM_SHORT_EXT; rb = rb > 0 ? rb : 0;
    re += MEM_SHORT_EXT; re = re < l_pac << 1 ? re : l_pac << 1;
    if (rb < l_pac && l_pac < re) {
        if (mid < l_pac) re = l_pac;
        else rb = l_pac;
    }
    if (qe - qb >= MEM_SHORT_LEN || re - rb >= MEM_SHORT_LEN) return -1;
    {   // bns_fetch_seq (bwa/bntseq.c:426-451): clip the window to the contig that holds `mid`
        int is_rev;
        int rid = pos2rid(ix, depos(ix, mid, &is_rev));
        i64 far_beg = ix.contig_off[rid], far_end = ix.contig_off[rid + 1];
        if (is_rev) { i64 t = far_beg; far_beg = (l_pac << 1) - far_end; far_end = (l_pac << 1) - t; }
        rb = rb > far_beg ? rb : far_beg;
        re = re < far_end ? re : far_end;
    }
    return ksw_i16_score(ix, qe - qb, query + qb, (int)(re - rb), rb, opt.mat, opt.o_del, opt.e_del, opt.o_ins, opt.e_ins);
}

} // namespace b200
