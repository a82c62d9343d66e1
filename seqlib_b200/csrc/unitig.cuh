// unitig.cuh -- per-string overlap facts of fermi-lite's unitig construction (fermi-lite/unitig.c), HD code.
//
//   utg_retrieve          <- fm6_retrieve        (fermi-lite/unitig.c:32-59)
//   utg_overlap_intv      <- overlap_intv        (fermi-lite/unitig.c:86-112)
//   utg_get_nei           <- fm6_get_nei         (fermi-lite/unitig.c:140-226)
//   utg_check_left(_simple) <- check_left(_simple) (fermi-lite/unitig.c:233-272)
//   utg_node              <- the index work of unitig1 / unitig_unidir (fermi-lite/unitig.c:274-366) for ONE string
//
// The reference walks unitigs seed by seed, and its three shared bitmaps (used / bend / visited) make the walk order
// dependent.  All FM-index work inside a walk, however, depends only on the string at the current end of the unitig:
// fm6_get_nei looks at s[beg..) = the last read, check_left at the last read plus the bases fm6_get_nei appended.  So the
// device computes, for EVERY string of the index independently (one string per thread), everything a walk could ask about
// it -- sequence, duplicate/contained status, right neighbours with overlap lengths, the bases an unambiguous extension
// appends, the verdict of check_left for that extension, and the `used` marks each of those steps would set -- and the
// order-dependent part (utg_walk.h, host) only chases these records through the bitmaps in the reference's seed order.
#pragma once
#include "fmd.cuh"
#include "sort.cuh"

namespace b200 {

struct UtgMark { u64 x0, x1, x2; };          // one set_bits(used, intv) call

struct UtgNei { u64 x0, x1, x2; i64 ovlp; };

struct UtgNode {             // 128 bytes = two cache lines: the host walk touches one record per absorbed read
    u64 x0, x1, x2;          // interval of $s$: x0 = first rank among the copies of s, x1 = same for revcomp(s), x2 = copies
    u64 ret_k;               // fm6_retrieve's return value
    i32 len;                 // length of s
    i32 flags;               // UTG_* below
    i32 n_nei;               // right neighbours found by fm6_get_nei (0 when it returns -1)
    i32 rbeg;                // fm6_get_nei's return value for beg = 0 (start of the neighbour inside s), -1: no overlap
    i32 ext_len;             // bases appended to s when n_nei == 1
    i32 cl;                  // check_left for that extension: 0 / -1 (only meaningful when n_nei == 1)
    i32 n_mark_r, n_mark_c;  // marks set by fm6_get_nei, marks set by check_left
    u64 seq_off;             // pools (UtgPools): s at seq[seq_off, +len), appended bases follow at seq[seq_off + len, +ext_len)
    u64 nei_off;             // nei[nei_off, +n_nei)  (x[0], x[1], overlap length)
    u64 mark_off;            // mark[mark_off, +n_mark_r) then +n_mark_c
    UtgNei nei0;             // copy of nei[nei_off] (valid when n_nei >= 1) and of the first appended bases, so that the common
    u8 ext8[8];              // step of a walk (one neighbour, a few new bases, no marks) needs no second random access
};

enum { UTG_DUP = 1, UTG_CONTAINED = 2, UTG_SHORT = 4, UTG_NO_OVLP = 8, UTG_OVERFLOW = 16 };

// What the forward (fm6_get_nei) or backward (check_left_simple) extension of ONE interval yields.  The extensions of all
// intervals of a round are independent of each other; the two-phase variants below compute them first (one thread, or the
// lanes of a group: the Coop policies) and then consume them in interval order with the reference's sequential logic.
struct UtgExt {
    u64 c0[3];               // ok[0] (x0, x1, x2)
    u64 s0[3];               // rld_extend0(ok[0]) when has0
    u64 ch[4][3];            // ok[1..4]; check_left_simple: ch[0] = ok[s[i]]
    u32 has0;                // ok[0].x[2] != 0 in a round after the first: s0 is valid
    u32 child;               // bit c-1: ok[c].x[2] != 0 and its sentinel extension is non-empty
};

struct UtgScratch {
    UtgExt *ext;                                  // cap entries, or null (one-phase code only)
    FmdIntv *a[2], *nei; i32 *cat; int cap;      // interval vectors (prev / curr / neighbours) and the category array
    int n[2], n_nei;
    u8 *s, *str; int s_cap;                      // the growing unitig string and check_left's reversed copy
    UtgMark *mark; int mark_cap, n_mark;
    bool overflow;
};

HD size_t utg_scratch_bytes(int cap, int s_cap, int mark_cap, bool with_ext = false)
{
    return (size_t)cap * (3 * sizeof(FmdIntv) + sizeof(i32) + (with_ext ? sizeof(UtgExt) : 0)) + (size_t)mark_cap * sizeof(UtgMark)
         + 2 * (size_t)((s_cap + 15) & ~15) + 64;
}

HD void utg_scratch_bind(UtgScratch &S, u8 *p, int cap, int s_cap, int mark_cap, bool with_ext = false)
{
    S.a[0] = (FmdIntv *)p; S.a[1] = S.a[0] + cap; S.nei = S.a[1] + cap;
    p += (size_t)cap * 3 * sizeof(FmdIntv);
    S.ext = nullptr;
    if (with_ext) { S.ext = (UtgExt *)p; p += (size_t)cap * sizeof(UtgExt); }
    S.mark = (UtgMark *)p; p += (size_t)mark_cap * sizeof(UtgMark);
    S.cat = (i32 *)p; p += (size_t)cap * sizeof(i32);
    p = (u8 *)(((uintptr_t)p + 15) & ~(uintptr_t)15);
    S.s = p; S.str = p + ((s_cap + 15) & ~15);
    S.cap = cap; S.s_cap = s_cap; S.mark_cap = mark_cap;
    S.n[0] = S.n[1] = S.n_nei = S.n_mark = 0; S.overflow = false;
}

HD void utg_mark(UtgScratch &S, const FmdIntv &p)
{
    if (S.n_mark >= S.mark_cap) { S.overflow = true; return; }
    UtgMark m; m.x0 = p.x[0]; m.x1 = p.x[1]; m.x2 = p.x[2];
    S.mark[S.n_mark++] = m;
}

// fm6_retrieve: the string whose sentinel row is x (symbols 1..4, REVERSED as the reference leaves it before seq_reverse).
// The backward extensions it performs are exactly those of overlap_intv(s, at5 = 0) that fm6_is_contained / fm6_get_nei run
// next on the same string, so the overlap candidates (suffix intervals preceded by a sentinel, depth >= min) are collected
// here: p[0..pn) in discovery order (shortest suffix first) with info = number of bases of the suffix; the caller turns that
// into positions and reverses the list.  Intervals of size 1 take the reference's shortcut (no extension): a size-1 interval
// whose row is preceded by a base has no sentinel extension, so nothing is missed.
HD u64 utg_retrieve(const FmdIndex &e, u64 x, u8 *s, int s_cap, int &l_out, FmdIntv &k2, int &contained, bool &overflow,
                    int min, FmdIntv *p, int &pn, int cap)
{
    u64 k = x, ok[6];
    FmdIntv ok2[6];
    int l = 0;
    contained = 0; pn = 0;
    k2.x[0] = k2.x[1] = k2.x[2] = 0; k2.info = 0;
    for (;;) {
        int c = fmd_rank1a(e, k + 1, ok);
        k = e.cnt[c] + ok[c] - 1;
        if (c == 0) break;
        if (l > 0) {
            if (k2.x[2] == 1) k2.x[0] = k;
            else {
                fmd_extend(e, k2, ok2, 1);
                if (l >= min && ok2[0].x[2]) {            // overlap_intv: depth = l, the suffix of l bases ends a read on its left
                    if (pn >= cap) { overflow = true; break; }
                    FmdIntv t = k2; t.info = (u64)l;
                    p[pn++] = t;
                }
                k2 = ok2[c];
            }
        } else fmd_set_intv(e, c, k2);
        if (l >= s_cap) { overflow = true; break; }
        s[l++] = (u8)c;
    }
    l_out = l;
    if (overflow || l == 0) return k;
    if (k2.x[2] != 1) {
        fmd_extend(e, k2, ok2, 1);
        if (ok2[0].x[2] != k2.x[2]) contained |= 1;
        k2 = ok2[0];
    } else k2.x[0] = k;
    fmd_extend(e, k2, ok2, 0);
    if (ok2[0].x[2] != k2.x[2]) contained |= 2;
    k2 = ok2[0];
    return k;
}

// overlap_intv: intervals of the suffixes (at5 = 0) / prefixes (at5 = 1) of seq that end a read, smallest first
HD FmdIntv utg_overlap_intv(const FmdIndex &e, int len, const u8 *seq, int min, int j, int at5, FmdIntv *p, int &pn, int cap,
                            int inc_sentinel, bool &overflow)
{
    int c, depth, dir, end;
    FmdIntv ik, ok[6];
    pn = 0;
    dir = at5 ? 1 : -1;
    end = at5 ? len : -1;
    c = seq[j];
    fmd_set_intv(e, c, ik);
    for (depth = 1, j += dir; j != end; j += dir, ++depth) {
        c = at5 ? fmd_comp(seq[j]) : seq[j];
        fmd_extend(e, ik, ok, !at5);
        if (!ok[c].x[2]) break;
        if (depth >= min && ok[0].x[2]) {
            if (pn >= cap) { overflow = true; break; }
            if (inc_sentinel) { ok[0].info = (u64)(i64)(j - dir); p[pn++] = ok[0]; }
            else { ik.info = (u64)(i64)(j - dir); p[pn++] = ik; }
        }
        ik = ok[c];
    }
    for (int a = 0, b = pn - 1; a < b; ++a, --b) { FmdIntv t = p[a]; p[a] = p[b]; p[b] = t; }
    return ik;
}

struct UtgInfoLess { HD bool operator()(const FmdIntv &a, const FmdIntv &b) const { return a.info < b.info; } };

// forward extension of one interval as fm6_get_nei consumes it (fermi-lite/unitig.c:162-186)
HD void utg_ext_forward(const FmdIndex &e, const FmdIntv &p, bool later_round, UtgExt &R)
{
    FmdIntv ok[6], ok0;
    fmd_extend(e, p, ok, 0);
    R.c0[0] = ok[0].x[0]; R.c0[1] = ok[0].x[1]; R.c0[2] = ok[0].x[2];
    R.has0 = 0; R.child = 0;
    if (ok[0].x[2] && later_round) {
        fmd_extend0(e, ok[0], ok0, 1);
        R.s0[0] = ok0.x[0]; R.s0[1] = ok0.x[1]; R.s0[2] = ok0.x[2];
        R.has0 = 1;
    }
    for (int c = 1; c < 5; ++c) {
        R.ch[c - 1][0] = ok[c].x[0]; R.ch[c - 1][1] = ok[c].x[1]; R.ch[c - 1][2] = ok[c].x[2];
        if (ok[c].x[2]) {
            fmd_extend0(e, ok[c], ok0, 1);
            if (ok0.x[2]) R.child |= 1u << (c - 1);
        }
    }
}

// backward extension of one interval as check_left_simple consumes it (fermi-lite/unitig.c:241-247)
HD void utg_ext_backward(const FmdIndex &e, const FmdIntv &p, int sym, UtgExt &R)
{
    FmdIntv ok[6];
    fmd_extend(e, p, ok, 1);
    R.c0[0] = ok[0].x[0]; R.c0[1] = ok[0].x[1]; R.c0[2] = ok[0].x[2];
    R.ch[0][0] = ok[sym].x[0]; R.ch[0][1] = ok[sym].x[1]; R.ch[0][2] = ok[sym].x[2];
}

// Coop policies.  UtgNoCoop: the reference-shaped one-phase loops (one thread per string, large inputs).  UtgScalarCoop: the
// two-phase loops with one thread doing every extension itself (host emulation of the logic the group kernel runs).  The
// device's group policy (fml.cu) shares the extensions of a round out over the lanes of a group.
struct UtgNoCoop { static constexpr bool kTwoPhase = false; };
struct UtgScalarCoop {
    static constexpr bool kTwoPhase = true;
    HD void forward(const FmdIndex &e, const FmdIntv *prev, int pn, const i32 *cat, bool later_round, UtgExt *R) const
    {
        for (int j = 0; j < pn; ++j) if (cat[j] >= 0) utg_ext_forward(e, prev[j], later_round, R[j]);
    }
    HD void backward(const FmdIndex &e, const FmdIntv *prev, int pn, int sym, UtgExt *R) const
    {
        for (int j = 0; j < pn; ++j) utg_ext_backward(e, prev[j], sym, R[j]);
    }
};

// fm6_get_nei with `used` always present (marks recorded in S.mark).  s / l: the string, extended in place.
// prev = S.a[0] on entry (may be pre-filled).  Returns rbeg or -1.
template <class Coop>
HD int utg_get_nei(const FmdIndex &e, int min_match, int beg, u8 *s, int &l, UtgScratch &S, const Coop &coop)
{
    int ori_l = l, j, i, c, rbeg, is_forked = 0;
    int pi = 0, ci = 1;              // indices of prev / curr in S.a
    FmdIntv ok[6], ok0;
    S.n[ci] = S.n_nei = 0;
    if (S.n[pi] == 0) {
        utg_overlap_intv(e, l - beg, s + beg, min_match, l - beg - 1, 0, S.a[pi], S.n[pi], S.cap, 0, S.overflow);
        if (S.overflow) return -1;
        if (S.n[pi] == 0) return -1;
        for (j = 0; j < S.n[pi]; ++j) S.a[pi][j].info += (u64)beg;
    }
    for (j = 0; j < S.n[pi]; ++j) S.cat[j] = 0;
    while (S.n[pi]) {
        FmdIntv *prev = S.a[pi], *curr = S.a[ci];
        int pn = S.n[pi], cn = 0;
        if constexpr (Coop::kTwoPhase) {
        coop.forward(e, prev, pn, S.cat, ori_l != l, S.ext);        // every extension this round can need
        for (j = 0; j < pn; ++j) {
            FmdIntv *p = &prev[j];
            if (S.cat[j] < 0) continue;
            const UtgExt &R = S.ext[j];
            if (R.has0) {                       // ok[0].x[2] && ori_l != l
                ok0.x[0] = R.s0[0]; ok0.x[1] = R.s0[1]; ok0.x[2] = R.s0[2];
                if (ok0.x[2]) {
                    if (R.c0[2] == p->x[2] && p->x[2] == ok0.x[2]) {
                        int cat0 = S.cat[j];
                        ok0.info = (u64)(i64)(ori_l - (i64)(p->info & 0xffffffffu));
                        for (i = j; i < pn && S.cat[i] == cat0; ++i) S.cat[i] = -1;
                        if (S.n_nei >= S.cap) { S.overflow = true; return -1; }
                        S.nei[S.n_nei++] = ok0;
                        continue;
                    } else utg_mark(S, ok0);
                }
            }
            if (S.cat[j] < 0) continue;
            for (c = 1; c < 5; ++c)
                if (R.child >> (c - 1) & 1) {
                    FmdIntv t;
                    t.x[0] = R.ch[c - 1][0]; t.x[1] = R.ch[c - 1][1]; t.x[2] = R.ch[c - 1][2];
                    t.info = (p->info & 0xfffffff0ffffffffull) | (u64)c << 32;
                    if (cn >= S.cap) { S.overflow = true; return -1; }
                    curr[cn++] = t;
                }
        }
        } else {
        for (j = 0; j < pn; ++j) {
            FmdIntv *p = &prev[j];
            if (S.cat[j] < 0) continue;
            fmd_extend(e, *p, ok, 0);
            if (ok[0].x[2] && ori_l != l) {
                fmd_extend0(e, ok[0], ok0, 1);
                if (ok0.x[2]) {
                    if (ok[0].x[2] == p->x[2] && p->x[2] == ok0.x[2]) {
                        int cat0 = S.cat[j];
                        ok0.info = (u64)(i64)(ori_l - (i64)(p->info & 0xffffffffu));
                        for (i = j; i < pn && S.cat[i] == cat0; ++i) S.cat[i] = -1;
                        if (S.n_nei >= S.cap) { S.overflow = true; return -1; }
                        S.nei[S.n_nei++] = ok0;
                        continue;
                    } else utg_mark(S, ok0);
                }
            }
            if (S.cat[j] < 0) continue;
            for (c = 1; c < 5; ++c)
                if (ok[c].x[2]) {
                    fmd_extend0(e, ok[c], ok0, 1);
                    if (ok0.x[2]) {
                        ok[c].info = (p->info & 0xfffffff0ffffffffull) | (u64)c << 32;
                        if (cn >= S.cap) { S.overflow = true; return -1; }
                        curr[cn++] = ok[c];
                    }
                }
        }
        }
        S.n[ci] = cn;
        if (cn) {
            u32 last, cat0;
            c = (int)(curr[0].info >> 32 & 0xf);
            if (l >= S.s_cap) { S.overflow = true; return -1; }
            s[l++] = (u8)fmd_comp(c);
            introsort((size_t)cn, curr, UtgInfoLess());
            last = (u32)(curr[0].info >> 32);
            S.cat[0] = 0;
            curr[0].info &= 0xffffffffull;
            for (j = 1, cat0 = 0; j < cn; ++j) {
                if ((u32)(curr[j].info >> 32) != last) { last = (u32)(curr[j].info >> 32); cat0 = (u32)j; }
                S.cat[j] = (i32)cat0;
                curr[j].info = (curr[j].info & 0xffffffffull) | (u64)cat0 << 36;
            }
            if (cat0 != 0) is_forked = 1;
        }
        int t = ci; ci = pi; pi = t;
    }
    // the reference leaves the vectors swapped an arbitrary number of times; callers only rely on both being empty or reset
    S.n[0] = S.n[1] = 0;
    if (S.n_nei == 0) return -1;
    rbeg = ori_l - (int)(u32)S.nei[0].info;
    if (S.n_nei == 1 && is_forked) {
        fmd_set_intv(e, 0, ok0);
        for (i = rbeg; i < ori_l; ++i) {
            fmd_extend(e, ok0, ok, 0);
            ok0 = ok[fmd_comp(s[i])];
        }
        for (i = ori_l; i < l; ++i) {
            int c0 = -1;
            fmd_extend(e, ok0, ok, 0);
            for (c = 1, j = 0; c < 5; ++c)
                if (ok[c].x[2] && ok[c].x[0] <= S.nei[0].x[0] && ok[c].x[0] + ok[c].x[2] >= S.nei[0].x[0] + S.nei[0].x[2]) { ++j; c0 = c; }
            if (j == 0 && ok[0].x[2]) break;
            if (c0 < 0) break;       // the reference asserts j == 1 here
            s[i] = (u8)fmd_comp(c0);
            ok0 = ok[c0];
        }
        l = i;
    }
    if (S.n_nei > 1) l = ori_l;
    return rbeg;
}

// check_left_simple
template <class Coop>
HD int utg_check_left_simple(const FmdIndex &e, int min_match, int beg, int rbeg, const u8 *s, int l, UtgScratch &S, const Coop &coop)
{
    FmdIntv ok[6];
    int pi = 0, ci = 1, i, j;
    utg_overlap_intv(e, l, s, min_match, rbeg, 1, S.a[pi], S.n[pi], S.cap, 1, S.overflow);
    if (S.overflow) return -1;
    for (i = rbeg - 1; i >= beg; --i) {
        FmdIntv *prev = S.a[pi], *curr = S.a[ci];
        int cn = 0;
        if constexpr (Coop::kTwoPhase) {
        coop.backward(e, prev, S.n[pi], (int)s[i], S.ext);
        for (j = 0; j < S.n[pi]; ++j) {
            FmdIntv *p = &prev[j];
            const UtgExt &R = S.ext[j];
            if (R.c0[2]) { FmdIntv t; t.x[0] = R.c0[0]; t.x[1] = R.c0[1]; t.x[2] = R.c0[2]; t.info = 0; utg_mark(S, t); }
            if (R.c0[2] + R.ch[0][2] != p->x[2]) { S.n[ci] = cn; return -1; }
            FmdIntv t; t.x[0] = R.ch[0][0]; t.x[1] = R.ch[0][1]; t.x[2] = R.ch[0][2]; t.info = 0;
            curr[cn++] = t;
        }
        } else {
        for (j = 0; j < S.n[pi]; ++j) {
            FmdIntv *p = &prev[j];
            fmd_extend(e, *p, ok, 1);
            if (ok[0].x[2]) utg_mark(S, ok[0]);
            if (ok[0].x[2] + ok[(int)s[i]].x[2] != p->x[2]) { S.n[ci] = cn; return -1; }
            curr[cn++] = ok[(int)s[i]];
        }
        }
        S.n[ci] = cn;
        int t = ci; ci = pi; pi = t;
    }
    return 0;
}

// check_left; the caller has exactly one neighbour in S.nei[0]
template <class Coop>
HD int utg_check_left(const FmdIndex &e, int min_match, int beg, int rbeg, const u8 *s, int l, UtgScratch &S, const Coop &coop)
{
    int i, ret;
    FmdIntv tmp;
    ret = utg_check_left_simple(e, min_match, beg, rbeg, s, l, S, coop);
    if (S.overflow) return -1;
    if (ret == 0) return 0;
    tmp = S.nei[0];
    S.n[0] = S.n[1] = S.n_nei = 0;
    int sl = 0;
    if (l - rbeg + 1 > S.s_cap) { S.overflow = true; return -1; }
    for (i = l - 1; i >= rbeg; --i) S.str[sl++] = (u8)fmd_comp(s[i]);
    utg_get_nei(e, min_match, 0, S.str, sl, S, coop);
    if (S.overflow) return -1;
    ret = S.n_nei > 1 ? -1 : 0;
    S.n_nei = 1; S.nei[0] = tmp;
    return ret;
}

// Everything unitig1 / unitig_unidir can ask the index about string x.  seq_out (>= 2 * longest read + 2 bytes) receives
// the string followed by the appended bases; nei_out / mark_out are the thread's staging areas (S.nei / S.mark).
template <class Coop>
HD void utg_node(const FmdIndex &e, int min_match, u64 x, UtgScratch &S, UtgNode &N, const Coop &coop)
{
    FmdIntv intv0;
    int contained = 0, l = 0;
    N.flags = 0; N.n_nei = 0; N.rbeg = -1; N.ext_len = 0; N.cl = 0; N.n_mark_r = N.n_mark_c = 0;
    N.nei0.x0 = N.nei0.x1 = N.nei0.x2 = 0; N.nei0.ovlp = 0;
    for (int a = 0; a < 8; ++a) N.ext8[a] = 0;
    S.n[0] = S.n[1] = S.n_nei = S.n_mark = 0; S.overflow = false;
    N.ret_k = utg_retrieve(e, x, S.s, S.s_cap / 2, l, intv0, contained, S.overflow, min_match, S.a[0], S.n[0], S.cap);
    // seq_reverse
    for (int a = 0, b = l - 1; a < b; ++a, --b) { u8 t = S.s[a]; S.s[a] = S.s[b]; S.s[b] = t; }
    // overlap candidates: suffix of d bases starts at l - d; longest suffix (smallest interval) first, as kv_reverse leaves them
    for (int a = 0; a < S.n[0]; ++a) S.a[0][a].info = (u64)(l - (int)S.a[0][a].info);
    for (int a = 0, b = S.n[0] - 1; a < b; ++a, --b) { FmdIntv t = S.a[0][a]; S.a[0][a] = S.a[0][b]; S.a[0][b] = t; }
    N.len = l;
    N.x0 = intv0.x[0]; N.x1 = intv0.x[1]; N.x2 = intv0.x[2];
    if (S.overflow) { N.flags |= UTG_OVERFLOW; return; }
    // a copy that is not the first of its group never seeds a unitig, but walks reach the group through its first rank,
    // whichever copy sits there: the overlap record is computed for every copy
    if (intv0.x[2] > 1 && N.ret_k != intv0.x[0]) N.flags |= UTG_DUP;
    if (contained) { N.flags |= UTG_CONTAINED; return; }
    if (l <= min_match) { N.flags |= UTG_SHORT; return; }
    // fm6_is_contained only pre-computes the overlap list (its verdict is not used by unitig1): utg_retrieve has left that
    // list in S.a[0].  An empty list means "no overlap" (fm6_get_nei would rebuild the same empty list and return -1).
    int rbeg = S.n[0] ? utg_get_nei(e, min_match, 0, S.s, l, S, coop) : -1;
    if (S.overflow) { N.flags |= UTG_OVERFLOW; return; }
    N.n_mark_r = S.n_mark;
    N.rbeg = rbeg; N.n_nei = rbeg < 0 ? 0 : S.n_nei;
    if (rbeg < 0) { N.flags |= UTG_NO_OVLP; l = N.len; }
    N.ext_len = l - N.len;
    if (N.n_nei >= 1) { N.nei0.x0 = S.nei[0].x[0]; N.nei0.x1 = S.nei[0].x[1]; N.nei0.x2 = S.nei[0].x[2]; N.nei0.ovlp = (i64)S.nei[0].info; }
    for (int a = 0; a < 8; ++a) N.ext8[a] = a < N.ext_len ? S.s[N.len + a] : (u8)0;
    if (N.n_nei == 1) {
        N.cl = utg_check_left(e, min_match, 0, rbeg, S.s, l, S, coop);
        if (S.overflow) { N.flags |= UTG_OVERFLOW; return; }
        N.n_mark_c = S.n_mark - N.n_mark_r;
    }
}

} // namespace b200
