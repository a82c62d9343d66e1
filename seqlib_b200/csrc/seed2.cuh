// seed2.cuh -- mem_collect_intv (bwa/bwamem.c:140-188) shaped for the GPU: one extension site.
//
// seed.cuh states SMEM seeding as the reference does (bwt_smem1a / bwt_seed_strategy1, bwa/bwt.c:289-379): nested
// loops with bwt_extend in three places, two work lists of 32-byte intervals per read in scratch memory.  Profiled on
// a B200 that shape is bound by instruction issue (a third of the lanes active) and by scratch traffic, not by the
// Occ gathers.  Here the same computation is a small machine that yields every extension:
//
//     request(a, o, s, c)  ->  extend_lean  ->  consume(na, no, ns)
//
// so a warp has ONE place where the two dependent 32-byte Occ gathers and the popcounts happen, reached by all lanes
// together.  The work list is a single array of 16-byte packed intervals (coordinates < 2^36, end < 2^16) that the
// forward sweep fills from the top down -- it is born reversed, as the reference wants it after its in-place
// reversal -- and that the backward sweep compacts in place (row i+1 never has more survivors than row i read so
// far).  The caller keeps that list in shared memory.  SMEMs shorter than min_seed_len are dropped when they are
// produced (the reference drops them after each bwt_smem1 call; the "is this SMEM contained in the previous one"
// test uses the unfiltered start, kept in a register).  Intervals leave in production order; the final order is
// the sort by (start, end), done by the caller -- equal keys mean the same substring and therefore identical
// intervals, so the reference's unstable introsort cannot order them differently in any visible way.
//
// Restrictions (the caller routes everything else to seed.cuh): no base > 3 in the read, len < 65536,
// seq_len < 2^36, and the list capacity.
//
// Prefix-interval tables (SeedTab).  The bi-interval of a string is a pure function of the index, so for every
// string of length j <= K the result of "extend to that string" is tabulated once per index: level j holds 4^j packed
// 16-byte intervals, keyed by the string itself (base i in bits 2i..).  Every extension whose RESULT is a string of
// length <= K -- the first K steps of each forward sweep, the short entries of every backward row, the first K steps
// of every bwt_seed_strategy1 start -- becomes one 16-byte gather (L2-resident for the low levels) instead of two
// dependent-free 32-byte Occ gathers.  The tables are produced by the same extend_lean code, level j from level
// j-1 (k_seedtab_level, engine.cu), so they hold exactly what the iterated bwt_extend (bwa/bwt.c:262-275) computes,
// and interval lists stay identical to the reference's.  With 180 GB of HBM, K = 15 (23 GB) fits next to a 3 Gb index.
#pragma once
#include "common.cuh"
#include "fmindex.cuh"
#include "seed.cuh"

namespace b200 {

struct PIntv { u32 w0, w1, w2, w3; };   // x0, x1, x2 low words; w3 = x0>>32 | (x1>>32)<<4 | (x2>>32)<<8 | end<<16

HD PIntv pintv_pack(u64 x0, u64 x1, u64 x2, u32 end)
{
    PIntv p;
    p.w0 = (u32)x0; p.w1 = (u32)x1; p.w2 = (u32)x2;
    p.w3 = (u32)(x0 >> 32) | (u32)(x1 >> 32) << 4 | (u32)(x2 >> 32) << 8 | end << 16;
    return p;
}
HD void pintv_unpack(const PIntv &p, u64 &x0, u64 &x1, u64 &x2, u32 &end)
{
    x0 = (u64)p.w0 | (u64)(p.w3 & 15u) << 32;
    x1 = (u64)p.w1 | (u64)((p.w3 >> 4) & 15u) << 32;
    x2 = (u64)p.w2 | (u64)((p.w3 >> 8) & 15u) << 32;
    end = p.w3 >> 16;
}

struct SeedTab { const PIntv *base; int K; };    // K == 0: no tables
HD u64 seedtab_level_off(int j) { return ((1ull << (2 * j)) - 4) / 3; }      // entries of levels 1..j-1 (a multiple of 4)
HD u64 seedtab_entries(int K) { return K > 0 ? seedtab_level_off(K + 1) : 0; }

// bwt_extend (bwa/bwt.c:262-275) for the one child the callers use, on raw coordinates:
//   a = the coordinate the Occ ranks are taken on (x[!is_back]), o = the other one, s = interval size.
// Ranks a-1 and a-1+s are never -1 here (every interval starts at L2[c]+1 >= 1).
// the arithmetic of extend_lean on two loaded blocks (b1 holds rank kk, b2 rank ll)
HD void extend_blocks(const DevIndex &ix, const OccLoad &b1, const OccLoad &b2, u64 kk, u64 ll, u32 dk, u32 dl, u64 o, int c,
                      u64 &na, u64 &no, u64 &ns)
{
    const int rk = (int)(kk & 63), rl = (int)(ll & 63);          // ranks inside the block, minus one
    const u64 mk = (2ull << rk) - 1, ml = (2ull << rl) - 1;      // rk == 63: 2<<63 wraps to 0, minus 1 = all ones
    const u64 klo = b1.s0 & mk, khi = b1.s1 & mk, llo = b2.s0 & ml, lhi = b2.s1 & ml;
    const u32 kpl = popc64(klo), kph = popc64(khi), kpt = popc64(klo & khi);
    const u32 lpl = popc64(llo), lph = popc64(lhi), lpt = popc64(llo & lhi);
    // counts of C, G, T up to k and up to l (A only where needed)
    const u32 tk1 = b1.c1 + (kpl - kpt), tk2 = b1.c2 + (kph - kpt), tk3 = b1.c3 + kpt;
    const u32 tl1 = b2.c1 + (lpl - lpt), tl2 = b2.c2 + (lph - lpt), tl3 = b2.c3 + lpt;
    const u32 s1 = tl1 - tk1, s2 = tl2 - tk2, s3 = tl3 - tk3;
    const u32 tk0 = b1.c0 + ((u32)rk + 1 - kpl - kph + kpt);
    const u32 tl0 = b2.c0 + ((u32)rl + 1 - lpl - lph + lpt);
    const u32 s0 = tl0 - tk0;
    const u32 tkc = c == 0 ? tk0 : c == 1 ? tk1 : c == 2 ? tk2 : tk3;
    const u32 sc = c == 0 ? s0 : c == 1 ? s1 : c == 2 ? s2 : s3;
    const u64 gt = c == 0 ? (u64)s1 + s2 + s3 : c == 1 ? (u64)s2 + s3 : c == 2 ? (u64)s3 : 0ull;   // children > c come first on the other side
    na = ix.L2[c] + 1 + tkc;
    no = o + (dl - dk) + gt;        // dl - dk = 1 iff the '$' rank lies inside [a, a+s)
    ns = sc;
}

template <class Ctr>
HD void extend_lean(const DevIndex &ix, u64 a, u64 o, u64 s, int c, u64 &na, u64 &no, u64 &ns, Ctr &ctr)
{
    const u64 k = a - 1, l = k + s;
    const u32 dk = k >= ix.primary, dl = l >= ix.primary;
    const u64 kk = k - dk, ll = l - dl;
    const u64 bk = kk >> 6, bl = ll >> 6;
    // both gathers are issued before either is used, unconditionally: when k and l share a block the second request
    // merges with the first in L1, whereas a predicated second load would have to wait for the first to land
    OccLoad b1 = load_block(ix, bk);
    OccLoad b2 = load_block(ix, bl);
    ctr.occ_blocks += bl != bk ? 2 : 1;
    extend_blocks(ix, b1, b2, kk, ll, dk, dl, o, c, na, no, ns);
}

// One extension through the tables when the result string is short enough (tl = its length, key = the string), else
// through the Occ blocks.  Written so that a warp issues ONE pair of 256-bit loads whatever its lanes need: a table
// lane loads the 32-byte sector that holds its 16-byte entry (twice -- the second request merges in L1).
// fwd: the caller extends forward (result coordinates swap roles, see SeedMachine::request).
template <class Ctr>
HD void extend_or_lookup(const DevIndex &ix, const SeedTab &tab, int tl, u32 key, bool fwd, u64 a, u64 o, u64 s, int c,
                         u64 &na, u64 &no, u64 &ns, Ctr &ctr)
{
    const u64 k = a - 1, l = k + s;
    const u32 dk = k >= ix.primary, dl = l >= ix.primary;
    const u64 kk = k - dk, ll = l - dl;
    u64 bk = kk >> 6, bl = ll >> 6;
    const OccBlock *p1 = ix.occ + bk, *p2 = ix.occ + bl;
    u32 half = 0;
    if (tl) {
        const PIntv *e = tab.base + seedtab_level_off(tl) + key;
        half = (u32)(((uintptr_t)e >> 4) & 1);
        p1 = p2 = (const OccBlock *)((uintptr_t)e & ~(uintptr_t)31);
        if (tl <= 10) ctr.tab_lo++; else ctr.tab_hi++;
    } else ctr.occ_blocks += bl != bk ? 2 : 1;
    const OccLoad b1 = load_block_at(p1), b2 = load_block_at(p2);
    extend_blocks(ix, b1, b2, kk, ll, dk, dl, o, c, na, no, ns);
    if (tl) {
        PIntv t;
        if (half) { t.w0 = (u32)b1.s0; t.w1 = (u32)(b1.s0 >> 32); t.w2 = (u32)b1.s1; t.w3 = (u32)(b1.s1 >> 32); }
        else { t.w0 = b1.c0; t.w1 = b1.c1; t.w2 = b1.c2; t.w3 = b1.c3; }
        u64 t0, t1, t2; u32 e_;
        pintv_unpack(t, t0, t1, t2, e_);
        na = fwd ? t1 : t0; no = fwd ? t0 : t1; ns = t2;
    }
}

// One read's seeding as a resumable machine.  List: get(e) / set(e, PIntv) over `cap` entries; Query: operator[](i) in 0..3.
template <class List, class Query>
struct SeedMachine {
    enum { M_DONE = 0, M_FWD, M_BWD, M_P3, M_TASK, M_ENDFWD, M_LASTROW };
    int mode, pass, x, k2, old_n, sx, i, j, nprev, ncurr, top, ret, last_start, first, ovf;
    int len, cap, min_seed_len, split_len, split_width, max_intv3, min_intv, K, last_pass;
    u64 x0, x1, x2, lastcurr;     // ik of the forward sweeps / last size pushed in this backward row
    u64 p0, p1, p2;               // the list entry being extended backwards
    u32 iend, pend;
    bool have_p;                  // p0..p2 hold the entry's interval (false: only its end was read)
    List L; Query q; IntvSink out;

    // last_pass_ = 2: stop after the SMEM passes (bwt_smem1a calls, re-seeding); start3() then runs the third pass
    // (bwt_seed_strategy1) on its own -- k_seed2 / k_seed3 split them so that a warp never mixes the two kinds of lanes
    HD void init(const Opt &opt, int len_, int cap_, const List &L_, const Query &q_, const IntvSink &out_, int K_ = 0, int last_pass_ = 3)
    {
        len = len_; cap = cap_; L = L_; q = q_; out = out_; K = K_; last_pass = last_pass_;
        min_seed_len = opt.min_seed_len;
        split_len = (int)(opt.min_seed_len * opt.split_factor + .499);
        split_width = opt.split_width;
        max_intv3 = (int)opt.max_mem_intv;
        pass = 1; x = 0; k2 = 0; old_n = 0; ovf = 0; mode = M_DONE;
        out.n = 0; out.overflow = false;
    }

    HD void fail() { ovf = 1; mode = M_DONE; }

    HD void emit(u64 e0, u64 e1, u64 e2, u64 info)
    {
        Intv v; v.x0 = e0; v.x1 = e1; v.x2 = e2; v.info = info;
        out.push(v);
        if (out.overflow) fail();
    }

    HD void push_fwd()
    {
        if (top == 0) { fail(); return; }
        L.put(--top, x0, x1, x2, iend);
        ret = (int)iend;
    }

    // Everything between two extensions that is not the per-extension bookkeeping: ending a forward sweep, the row
    // i == -1 of a backward sweep, picking the next bwt_smem1 / bwt_seed_strategy1 call (bwa/bwamem.c:150-184).
    // A loop over transient modes instead of mutually recursive helpers, so that it inlines and the state stays in registers.
    HD void settle(const DevIndex &ix)
    {
        for (;;) {
            if (mode == M_ENDFWD) {
                nprev = cap - top; ncurr = 0; j = 0; first = 1; last_start = 0;
                i = sx - 1;
                mode = i < 0 ? M_LASTROW : M_BWD;
            } else if (mode == M_LASTROW) {
                // bwa/bwt.c:325-345 with c = -1: only the longest survivor can be reported, with start 0
                if (first || 0 < last_start) {
                    u64 e0, e1, e2; u32 e;
                    const bool ok = L.take(top, e0, e1, e2, e);
                    if ((int)e >= min_seed_len) { if (!ok) { fail(); return; } emit(e0, e1, e2, (u64)e); if (ovf) return; }
                }
                mode = M_TASK;
                if (pass == 1) x = ret;
            } else if (mode == M_TASK) {
                if (pass == 1) {
                    if (x >= len) { pass = 2; old_n = out.n; k2 = 0; }
                    else { sx = x; min_intv = 1; }
                }
                if (pass == 2) {
                    bool have = false;
                    while (k2 < old_n) {
                        const Intv p = out.a[k2++];
                        int start = (int)(p.info >> 32), end = (int)(i32)p.info;
                        if (end - start < split_len || p.x2 > (u64)split_width) continue;
                        sx = (start + end) >> 1; min_intv = (int)p.x2 + 1;
                        have = true;
                        break;
                    }
                    if (!have) {
                        pass = 3; x = 0;
                        if (max_intv3 <= 0) { mode = M_DONE; return; }
                    }
                }
                if (pass == 3) {
                    if (last_pass < 3 || x >= len) { mode = M_DONE; return; }
                    // the first min(K, min_seed_len) - 1 extensions of a bwt_seed_strategy1 start (bwa/bwt.c:355-379) can
                    // neither report nor stop (i - x < min_seed_len): one table lookup of q[x, x + jump) replaces them
                    const int jump = K < min_seed_len ? K : min_seed_len;
                    if (jump >= 2 && x + jump <= len) {
                        x0 = x1 = 1; x2 = 0;             // placeholders: the lookup result overwrites them
                        i = x + jump - 1;
                        mode = M_P3;
                        return;
                    }
                    Intv t; set_intv(ix, q[x], t);
                    x0 = t.x0; x1 = t.x1; x2 = t.x2;
                    i = x + 1;
                    mode = i < len ? M_P3 : M_DONE;      // a start on the last base extends nothing and reports nothing
                    return;
                }
                // begin the forward sweep of bwt_smem1a at sx
                Intv t; set_intv(ix, q[sx], t);
                x0 = t.x0; x1 = t.x1; x2 = t.x2; iend = (u32)(sx + 1);
                top = cap;
                i = sx + 1;
                mode = M_FWD;
                if (i >= len) { push_fwd(); if (ovf) return; mode = M_ENDFWD; }
            } else return;
        }
    }

    HD void start(const DevIndex &ix)
    {
        if (len < min_seed_len) { mode = M_DONE; return; }
        mode = M_TASK;
        settle(ix);
    }

    // the third pass alone, after init(): n0 intervals of the first two passes are already in the sink
    HD void start3(const DevIndex &ix, int n0)
    {
        out.n = n0;
        pass = 3; x = 0;
        if (len < min_seed_len || max_intv3 <= 0) { mode = M_DONE; return; }
        mode = M_TASK;
        settle(ix);
    }

    // The next extension and the string it produces: q[st, st + ln).  tl = ln when a table level holds it, else 0.
    // A backward step that a table answers needs only the END of its list entry (the string is q[i, end)): the entry's
    // interval is fetched lazily, when the entry dies as an SMEM -- lists that keep intervals outside shared memory
    // (k_seed2's HybridList) then touch them in a third of the steps only.
    HD void request(u64 &a, u64 &o, u64 &s, int &c, int &tl, u32 &key, bool &fwd)
    {
        fwd = mode != M_BWD;
        if (mode == M_BWD) {
            pend = L.end(top + j);
            const int ln = (int)pend - i;
            tl = ln <= K ? ln : 0;
            c = q[i];
            if (tl) { have_p = false; a = 1; o = 1; s = 0; key = q.key(i, ln); return; }
            have_p = true;
            if (!L.take(top + j, p0, p1, p2, pend)) { p0 = p1 = 1; p2 = 0; fail(); }      // the list did not keep this interval: spill path
            a = p0; o = p1; s = p2; key = 0u;
            return;
        }
        a = x1; o = x0; s = x2; c = 3 - q[i];
        const int st = mode == M_FWD ? sx : x;
        const int ln = i + 1 - st;
        tl = ln <= K ? ln : 0;
        key = tl ? q.key(st, ln) : 0u;
    }

    HD void consume(const DevIndex &ix, u64 na, u64 no, u64 ns)
    {
        if (mode == M_BWD) {
            if (ns < (u64)min_intv) {
                if (ncurr == 0 && (first || i + 1 < last_start)) {
                    first = 0; last_start = i + 1;
                    if ((int)pend - (i + 1) >= min_seed_len) {
                        if (!have_p) { u32 e_; have_p = true; if (!L.take(top + j, p0, p1, p2, e_)) { fail(); return; } }
                        emit(p0, p1, p2, (u64)(i + 1) << 32 | pend);
                        if (ovf) return;
                    }
                }
            } else if (ncurr == 0 || ns != lastcurr) {
                L.put(top + ncurr, na, no, ns, pend);
                ++ncurr; lastcurr = ns;
            }
            if (++j == nprev) {
                if (ncurr == 0) { mode = M_TASK; if (pass == 1) x = ret; }
                else {
                    nprev = ncurr; ncurr = 0; j = 0;
                    if (--i < 0) mode = M_LASTROW;
                }
            }
        } else if (mode == M_FWD) {
            bool stop = false;
            if (ns != x2) {
                push_fwd();
                if (ovf) return;
                stop = ns < (u64)min_intv;
            }
            if (!stop) {
                x1 = na; x0 = no; x2 = ns; iend = (u32)(i + 1);
                if (++i == len) { push_fwd(); if (ovf) return; stop = true; }
            }
            if (stop) mode = M_ENDFWD;
        } else {        // M_P3
            if (ns < (u64)max_intv3 && i - x >= min_seed_len) {
                if (ns > 0) { emit(no, na, ns, (u64)x << 32 | (u64)(i + 1)); if (ovf) return; }
                x = i + 1;
                mode = M_TASK;
            } else {
                x1 = na; x0 = no; x2 = ns;
                if (++i >= len) mode = M_DONE;
            }
        }
        if (mode > M_P3) settle(ix);
    }
};

// plain-array list / byte query: the host emulation and the debug path
struct ArrayList {
    PIntv *p;
    HD void put(int e, u64 x0, u64 x1, u64 x2, u32 end) { p[e] = pintv_pack(x0, x1, x2, end); }
    HD bool take(int e, u64 &x0, u64 &x1, u64 &x2, u32 &end) const { pintv_unpack(p[e], x0, x1, x2, end); return true; }
    HD u32 end(int e) const { return p[e].w3 >> 16; }
};
struct ByteQuery {
    const u8 *p;
    HD int operator[](int i) const { return p[i]; }
    HD u32 key(int st, int ln) const { u32 k = 0; for (int t = 0; t < ln; ++t) k |= (u32)p[st + t] << (2 * t); return k; }
};

HD bool seed2_eligible(const DevIndex &ix, int len, const u8 *seq)
{
    if (len >= 65536 || ix.seq_len >= (1ull << 36)) return false;
    for (int i = 0; i < len; ++i) if (seq[i] > 3) return false;
    return true;
}

HD void sort_intv_by_info(Intv *a, int n)
{
    for (int i = 1; i < n; ++i) {
        Intv v = a[i];
        int j = i;
        for (; j > 0 && v.info < a[j - 1].info; --j) a[j] = a[j - 1];
        a[j] = v;
    }
}

// scalar driver: same results as collect_intv() for eligible reads; returns false when the list capacity was exceeded
template <class Ctr>
HD bool collect_intv_v2(const DevIndex &ix, const Opt &opt, int len, const u8 *seq, IntvSink &out, PIntv *list, int cap, Ctr &ctr,
                        const SeedTab *tab = nullptr)
{
    ArrayList L; L.p = list;
    ByteQuery q; q.p = seq;
    SeedTab none; none.base = nullptr; none.K = 0;
    const SeedTab &T = tab ? *tab : none;
    SeedMachine<ArrayList, ByteQuery> m;
    for (int part = 0; part < 2; ++part) {             // the SMEM passes, then bwt_seed_strategy1: two machines, like k_seed2 + k_seed3
        if (part == 0) { m.init(opt, len, cap, L, q, out, T.K, 2); m.start(ix); }
        else { const int n0 = m.out.n; IntvSink o2 = m.out; m.init(opt, len, cap, L, q, o2, T.K, 3); m.start3(ix, n0); }
        while (m.mode != 0) {
            u64 a, o, s, na, no, ns; int c, tl; u32 key; bool fwd;
            m.request(a, o, s, c, tl, key, fwd);
            extend_or_lookup(ix, T, tl, key, fwd, a, o, s, c, na, no, ns, ctr);
            m.consume(ix, na, no, ns);
        }
        if (m.ovf) break;
    }
    out = m.out;
    if (m.ovf) return false;
    sort_intv_by_info(out.a, out.n);
    return true;
}

} // namespace b200
