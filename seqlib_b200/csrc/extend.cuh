// extend.cuh -- chain -> alignment regions for one read.
//   chain2aln <- mem_chain2aln (bwa/bwamem.c:658-812), cal_max_gap (:647-654)
// The reference window [rmax0, rmax1) is never unpacked: ksw reads the 2-bit
// text in place (forwards for the right extension, backwards for the left).
#pragma once
#include "common.cuh"
#include "fmindex.cuh"
#include "ksw.cuh"
#include "sort.cuh"

namespace b200 {

HD int cal_max_gap(const Opt &opt, int qlen)
{
    int l_del = (int)((double)(qlen * opt.a - opt.o_del) / opt.e_del + 1.);
    int l_ins = (int)((double)(qlen * opt.a - opt.o_ins) / opt.e_ins + 1.);
    int l = l_del > l_ins ? l_del : l_ins;
    l = l > 1 ? l : 1;
    return l < opt.w << 1 ? l : opt.w << 1;
}

struct RegSink { Reg *a; int n, cap; bool overflow; };

struct U64Less { HD bool operator()(u64 a, u64 b) const { return a < b; } };

#define B200_MAX_BAND_TRY 2

// cs[0..cn): the chain's seeds in chain order; srt: cn u64 of scratch; eh: l_query+1 cells.
template <class Ctr>
HD void chain2aln(const DevIndex &ix, const Opt &opt, int l_query, const u8 *query,
                  const Seed *cs, int cn, int c_rid, float c_frac_rep,
                  RegSink &av, u64 *srt, EH *eh, Ctr &ctr)
{
    int i, k, max_off[2], aw[2];
    i64 l_pac = ix.l_pac, rmax[2], tmp, max = 0;
    if (cn == 0) return;
    rmax[0] = l_pac << 1; rmax[1] = 0;
    for (i = 0; i < cn; ++i) {
        const Seed &t = cs[i];
        i64 b = t.rbeg - (t.qbeg + cal_max_gap(opt, t.qbeg));
        i64 e = t.rbeg + t.len + ((l_query - t.qbeg - t.len) + cal_max_gap(opt, l_query - t.qbeg - t.len));
        rmax[0] = rmax[0] < b ? rmax[0] : b;
        rmax[1] = rmax[1] > e ? rmax[1] : e;
        if (t.len > max) max = t.len;
    }
    rmax[0] = rmax[0] > 0 ? rmax[0] : 0;
    rmax[1] = rmax[1] < l_pac << 1 ? rmax[1] : l_pac << 1;
    if (rmax[0] < l_pac && l_pac < rmax[1]) {
        if (cs[0].rbeg < l_pac) rmax[1] = l_pac;
        else rmax[0] = l_pac;
    }
    {   // bns_fetch_seq (bwa/bntseq.c:426-451): clip the window to the contig of the first seed
        int is_rev;
        int rid = pos2rid(ix, depos(ix, cs[0].rbeg, &is_rev));
        i64 far_beg = ix.contig_off[rid], far_end = ix.contig_off[rid + 1];
        if (is_rev) { i64 t2 = far_beg; far_beg = (l_pac << 1) - far_end; far_end = (l_pac << 1) - t2; }
        rmax[0] = rmax[0] > far_beg ? rmax[0] : far_beg;
        rmax[1] = rmax[1] < far_end ? rmax[1] : far_end;
        ctr.ref_bytes += (unsigned long long)((rmax[1] - rmax[0] + 3) >> 2);
    }
    for (i = 0; i < cn; ++i) srt[i] = (u64)cs[i].score << 32 | (u64)i;
    introsort((size_t)cn, srt, U64Less());

    for (k = cn - 1; k >= 0; --k) {
        const Seed *s = &cs[(u32)srt[k]];
        for (i = 0; i < av.n; ++i) {
            const Reg *p = &av.a[i];
            i64 rd; int qd, w, max_gap;
            if (s->rbeg < p->rb || s->rbeg + s->len > p->re || s->qbeg < p->qb || s->qbeg + s->len > p->qe) continue;
            if (s->len - p->seedlen0 > .1 * l_query) continue;
            qd = s->qbeg - p->qb; rd = s->rbeg - p->rb;
            max_gap = cal_max_gap(opt, qd < rd ? qd : (int)rd);
            w = max_gap < p->w ? max_gap : p->w;
            if (qd - rd < w && rd - qd < w) break;
            qd = p->qe - (s->qbeg + s->len); rd = p->re - (s->rbeg + s->len);
            max_gap = cal_max_gap(opt, qd < rd ? qd : (int)rd);
            w = max_gap < p->w ? max_gap : p->w;
            if (qd - rd < w && rd - qd < w) break;
        }
        if (i < av.n) {
            for (i = k + 1; i < cn; ++i) {
                if (srt[i] == 0) continue;
                const Seed *t = &cs[(u32)srt[i]];
                if (t->len < s->len * .95) continue;
                if (s->qbeg <= t->qbeg && s->qbeg + s->len - t->qbeg >= s->len >> 2 && t->qbeg - s->qbeg != t->rbeg - s->rbeg) break;
                if (t->qbeg <= s->qbeg && t->qbeg + t->len - s->qbeg >= s->len >> 2 && s->qbeg - t->qbeg != s->rbeg - t->rbeg) break;
            }
            if (i == cn) { srt[k] = 0; continue; }
        }
        if (av.n >= av.cap) { av.overflow = true; return; }
        Reg *a = &av.a[av.n++];
        {   // memset(a, 0, sizeof(mem_alnreg_t))
            a->rb = a->re = 0; a->qb = a->qe = a->rid = a->score = a->truesc = a->sub = a->alt_sc = a->csub = a->sub_n = 0;
            a->w = a->seedcov = a->secondary = a->secondary_all = a->seedlen0 = a->n_comp = a->is_alt = 0;
            a->frac_rep = 0; a->pad_ = 0; a->hash = 0;
        }
        a->w = aw[0] = aw[1] = opt.w;
        a->score = a->truesc = -1;
        a->rid = c_rid;
        if (s->qbeg) {                                   // left extension on reversed sequences
            int qle = 0, tle = 0, gtle = 0, gscore = 0;
            tmp = s->rbeg - rmax[0];
            BytesSeq qs; qs.p = query + s->qbeg - 1; qs.step = -1;
            TextSeq rs; rs.ix = &ix; rs.pos = s->rbeg - 1; rs.step = -1;
            for (i = 0; i < B200_MAX_BAND_TRY; ++i) {
                int prev = a->score;
                aw[0] = opt.w << i;
                ExtResult r = extend2(s->qbeg, qs, (int)tmp, rs, opt.mat, opt.o_del, opt.e_del, opt.o_ins, opt.e_ins,
                                      aw[0], opt.pen_clip5, opt.zdrop, s->len * opt.a, eh, ctr);
                a->score = r.score; qle = r.qle; tle = r.tle; gtle = r.gtle; gscore = r.gscore; max_off[0] = r.max_off;
                if (a->score == prev || max_off[0] < (aw[0] >> 1) + (aw[0] >> 2)) break;
            }
            if (gscore <= 0 || gscore <= a->score - opt.pen_clip5) {
                a->qb = s->qbeg - qle; a->rb = s->rbeg - tle;
                a->truesc = a->score;
            } else {
                a->qb = 0; a->rb = s->rbeg - gtle;
                a->truesc = gscore;
            }
        } else { a->score = a->truesc = s->len * opt.a; a->qb = 0; a->rb = s->rbeg; }

        if (s->qbeg + s->len != l_query) {               // right extension
            int qle = 0, tle = 0, qe, gtle = 0, gscore = 0, sc0 = a->score;
            i64 re;
            qe = s->qbeg + s->len;
            re = s->rbeg + s->len - rmax[0];
            BytesSeq qs; qs.p = query + qe; qs.step = 1;
            TextSeq rs; rs.ix = &ix; rs.pos = rmax[0] + re; rs.step = 1;
            for (i = 0; i < B200_MAX_BAND_TRY; ++i) {
                int prev = a->score;
                aw[1] = opt.w << i;
                ExtResult r = extend2(l_query - qe, qs, (int)(rmax[1] - rmax[0] - re), rs, opt.mat, opt.o_del, opt.e_del, opt.o_ins, opt.e_ins,
                                      aw[1], opt.pen_clip3, opt.zdrop, sc0, eh, ctr);
                a->score = r.score; qle = r.qle; tle = r.tle; gtle = r.gtle; gscore = r.gscore; max_off[1] = r.max_off;
                if (a->score == prev || max_off[1] < (aw[1] >> 1) + (aw[1] >> 2)) break;
            }
            if (gscore <= 0 || gscore <= a->score - opt.pen_clip3) {
                a->qe = qe + qle; a->re = rmax[0] + re + tle;
                a->truesc += a->score - sc0;
            } else {
                a->qe = l_query; a->re = rmax[0] + re + gtle;
                a->truesc += gscore - sc0;
            }
        } else { a->qe = l_query; a->re = s->rbeg + s->len; }

        for (i = 0, a->seedcov = 0; i < cn; ++i) {
            const Seed &t = cs[i];
            if (t.qbeg >= a->qb && t.qbeg + t.len <= a->qe && t.rbeg >= a->rb && t.rbeg + t.len <= a->re)
                a->seedcov += t.len;
        }
        a->w = aw[0] > aw[1] ? aw[0] : aw[1];
        a->seedlen0 = s->len;
        a->frac_rep = c_frac_rep;
    }
}

} // namespace b200
