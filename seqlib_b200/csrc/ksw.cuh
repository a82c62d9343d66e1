// ksw.cuh -- banded affine-gap DP, scalar per-thread formulation.
//   extend2 <- ksw_extend2 (bwa/ksw.c:416-515)
//   global2 <- ksw_global2 (bwa/ksw.c:540-642)
// Sequences are read through small accessor objects so the kernels can walk
// the 2-bit reference text and the read (forwards or backwards) in place
// instead of materialising reversed copies like mem_chain2aln does
// (bwa/bwamem.c:744-749).  The per-column state eh[] is persistent across rows
// exactly like the reference's array: cells outside the current band keep
// their last value and are re-read when the band grows again.
#pragma once
#include "common.cuh"
#include "fmindex.cuh"

namespace b200 {

struct EH { i32 h, e; };

struct ExtResult { int score, qle, tle, gtle, gscore, max_off; };

// byte-array sequence, optional reverse walk
struct BytesSeq {
    const u8 *p; int step;          // element i = p[i*step]
    HD int operator[](int i) const { return p[(i64)i * step]; }
};
// reference text window starting at text position `pos`, optional reverse walk
struct TextSeq {
    const DevIndex *ix; i64 pos; int step;
    HD int operator[](int i) const { return text_base(*ix, pos + (i64)i * step); }
};

// the same window with the current 32-base word kept in registers: one load per 32 rows instead of one per row
struct TextSeqC {
    const DevIndex *ix; i64 pos; int step;
    mutable u64 w; mutable i64 wi;
    HD TextSeqC(const DevIndex *ix_, i64 pos_, int step_) : ix(ix_), pos(pos_), step(step_), w(0), wi(-1) {}
    HD int operator[](int i) const
    {
        i64 p = pos + (i64)i * step;
        if ((p >> 5) != wi) {
            wi = p >> 5;
#if defined(__CUDA_ARCH__)
            w = __ldg(ix->text + wi);
#else
            w = ix->text[wi];
#endif
        }
        return (int)((w >> (2 * (p & 31))) & 3);
    }
};

template <class QSeq, class TSeq, class Ctr>
HD ExtResult extend2(int qlen, const QSeq &query, int tlen, const TSeq &target, const i8 *mat,
                     int o_del, int e_del, int o_ins, int e_ins, int w, int end_bonus, int zdrop, int h0,
                     EH *eh, Ctr &ctr)
{
    ExtResult R;
    int i, j, k, oe_del = o_del + e_del, oe_ins = o_ins + e_ins, beg, end, max, max_i, max_j, max_ins, max_del, max_ie, gscore, max_off;
    for (j = 0; j <= qlen; ++j) eh[j].h = eh[j].e = 0;
    eh[0].h = h0; eh[1].h = h0 > oe_ins ? h0 - oe_ins : 0;
    for (j = 2; j <= qlen && eh[j - 1].h > e_ins; ++j) eh[j].h = eh[j - 1].h - e_ins;
    for (i = 0, max = 0; i < 25; ++i) max = max > mat[i] ? max : mat[i];
    max_ins = (int)((double)(qlen * max + end_bonus - o_ins) / e_ins + 1.);
    max_ins = max_ins > 1 ? max_ins : 1;
    w = w < max_ins ? w : max_ins;
    max_del = (int)((double)(qlen * max + end_bonus - o_del) / e_del + 1.);
    max_del = max_del > 1 ? max_del : 1;
    w = w < max_del ? w : max_del;
    max = h0; max_i = max_j = -1; max_ie = -1; gscore = -1; max_off = 0;
    beg = 0; end = qlen;
    unsigned long long cells = 0;
    for (i = 0; i < tlen; ++i) {
        int t, f = 0, h1, m = 0, mj = -1;
        const i8 *q = mat + target[i] * 5;
        if (beg < i - w) beg = i - w;
        if (end > i + w + 1) end = i + w + 1;
        if (end > qlen) end = qlen;
        if (beg == 0) { h1 = h0 - (o_del + e_del * (i + 1)); if (h1 < 0) h1 = 0; }
        else h1 = 0;
        cells += end > beg ? end - beg : 0;
        for (j = beg; j < end; ++j) {
            EH *p = &eh[j];
            int h, M = p->h, e = p->e;
            p->h = h1;
            M = M ? M + q[query[j]] : 0;
            h = M > e ? M : e;
            h = h > f ? h : f;
            h1 = h;
            mj = m > h ? mj : j;
            m = m > h ? m : h;
            t = M - oe_del; t = t > 0 ? t : 0;
            e -= e_del; e = e > t ? e : t;
            p->e = e;
            t = M - oe_ins; t = t > 0 ? t : 0;
            f -= e_ins; f = f > t ? f : t;
        }
        eh[end].h = h1; eh[end].e = 0;
        if (j == qlen) {
            max_ie = gscore > h1 ? max_ie : i;
            gscore = gscore > h1 ? gscore : h1;
        }
        if (m == 0) break;
        if (m > max) {
            max = m; max_i = i; max_j = mj;
            k = mj - i; k = k < 0 ? -k : k;
            max_off = max_off > k ? max_off : k;
        } else if (zdrop > 0) {
            if (i - max_i > mj - max_j) {
                if (max - m - ((i - max_i) - (mj - max_j)) * e_del > zdrop) break;
            } else {
                if (max - m - ((mj - max_j) - (i - max_i)) * e_ins > zdrop) break;
            }
        }
        for (j = beg; j < end && eh[j].h == 0 && eh[j].e == 0; ++j) {}
        beg = j;
        for (j = end; j >= beg && eh[j].h == 0 && eh[j].e == 0; --j) {}
        end = j + 2 < qlen ? j + 2 : qlen;
    }
    ctr.sw_cells += cells; ctr.n_ext++;
    R.score = max; R.qle = max_j + 1; R.tle = max_i + 1; R.gtle = max_ie + 1; R.gscore = gscore; R.max_off = max_off;
    return R;
}

#define KSW_MINUS_INF (-0x40000000)

// Global alignment.  z == NULL: score only.  Otherwise z holds n_col*tlen
// direction bytes and the CIGAR (BAM words, op 0=M 1=I 2=D) is produced into
// cigar[] (capacity cap_cigar; *n_cigar = -1 on overflow).
template <class QSeq, class TSeq, class Ctr>
HD int global2(int qlen, const QSeq &query, int tlen, const TSeq &target, const i8 *mat,
               int o_del, int e_del, int o_ins, int e_ins, int w,
               EH *eh, u8 *z, u32 *cigar, int cap_cigar, int *n_cigar_, Ctr &ctr)
{
    int i, j, k, oe_del = o_del + e_del, oe_ins = o_ins + e_ins, score, n_col;
    if (n_cigar_) *n_cigar_ = 0;
    n_col = qlen < 2 * w + 1 ? qlen : 2 * w + 1;
    eh[0].h = 0; eh[0].e = KSW_MINUS_INF;
    for (j = 1; j <= qlen && j <= w; ++j) { eh[j].h = -(o_ins + e_ins * j); eh[j].e = KSW_MINUS_INF; }
    for (; j <= qlen; ++j) eh[j].h = eh[j].e = KSW_MINUS_INF;
    unsigned long long cells = 0;
    for (i = 0; i < tlen; ++i) {
        i32 f = KSW_MINUS_INF, h1, beg, end, t;
        const i8 *q = mat + target[i] * 5;
        beg = i > w ? i - w : 0;
        end = i + w + 1 < qlen ? i + w + 1 : qlen;
        h1 = beg == 0 ? -(o_del + e_del * (i + 1)) : KSW_MINUS_INF;
        u8 *zi = z ? z + (i64)i * n_col : 0;
        cells += end > beg ? end - beg : 0;
        for (j = beg; j < end; ++j) {
            EH *p = &eh[j];
            i32 h, m = p->h, e = p->e;
            u8 d;
            p->h = h1;
            m += q[query[j]];
            d = m >= e ? 0 : 1;
            h = m >= e ? m : e;
            d = h >= f ? d : 2;
            h = h >= f ? h : f;
            h1 = h;
            t = m - oe_del;
            e -= e_del;
            d |= e > t ? 1 << 2 : 0;
            e = e > t ? e : t;
            p->e = e;
            t = m - oe_ins;
            f -= e_ins;
            d |= f > t ? 2 << 4 : 0;
            f = f > t ? f : t;
            if (zi) zi[j - beg] = d;
        }
        eh[end].h = h1; eh[end].e = KSW_MINUS_INF;
    }
    score = eh[qlen].h;
    ctr.sw_cells += cells; ctr.n_global++;
    if (z && n_cigar_) {
        int n = 0, which = 0;
        bool ovf = false;
        i = tlen - 1; k = (i + w + 1 < qlen ? i + w + 1 : qlen) - 1;
        // ops are produced back to front; push_cigar (bwa/ksw.c:528-538) merges equal neighbours
#define PUSH_OP(op_, len_) do { \
        if (n == 0 || (int)(cigar[n - 1] & 0xf) != (op_)) { if (n < cap_cigar) cigar[n++] = (u32)(len_) << 4 | (op_); else ovf = true; } \
        else cigar[n - 1] += (u32)(len_) << 4; } while (0)
        while (i >= 0 && k >= 0) {
            which = z[(i64)i * n_col + (k - (i > w ? i - w : 0))] >> (which << 1) & 3;
            if (which == 0) { PUSH_OP(0, 1); --i; --k; }
            else if (which == 1) { PUSH_OP(2, 1); --i; }
            else { PUSH_OP(1, 1); --k; }
            if (ovf) break;
        }
        if (!ovf && i >= 0) PUSH_OP(2, i + 1);
        if (!ovf && k >= 0) PUSH_OP(1, k + 1);
#undef PUSH_OP
        if (ovf) { *n_cigar_ = -1; return score; }
        for (i = 0; i < n >> 1; ++i) swap_(cigar[i], cigar[n - 1 - i]);
        *n_cigar_ = n;
    }
    return score;
}

} // namespace b200
