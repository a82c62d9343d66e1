// seed.cuh -- SMEM / FM-index seeding for one read.
//   collect_intv  <-  mem_collect_intv   (bwa/bwamem.c:140-188)
//   smem1         <-  bwt_smem1a with max_intv = 0, i.e. bwt_smem1 (bwa/bwt.c:289-356)
//   seed_strategy1<-  bwt_seed_strategy1 (bwa/bwt.c:358-379)
// One thread runs one read; the two work lists of the bidirectional search
// live in per-thread scratch (len+1 entries each), results are appended to
// the read's interval slot in HBM.
#pragma once
#include "common.cuh"
#include "fmindex.cuh"
#include "sort.cuh"

namespace b200 {

struct IntvSink {       // bounded append-only output list
    Intv *a;
    int n, cap;
    bool overflow;
    HD void push(const Intv &v) { if (n < cap) a[n++] = v; else overflow = true; }
};

HD void reverse_(Intv *a, int n)
{
    for (int j = 0; j < n >> 1; ++j) swap_(a[j], a[n - 1 - j]);
}

// Appends the SMEMs through x to `out` (in start order, like the reference
// after its final reversal).  Returns the next x.  *n_new = how many were appended.
template <class Ctr>
HD int smem1(const DevIndex &ix, int len, const u8 *q, int x, int min_intv,
             IntvSink &out, int *n_new, Intv *prev, Intv *curr, Ctr &ctr)
{
    *n_new = 0;
    if (q[x] > 3) return x + 1;
    if (min_intv < 1) min_intv = 1;
    Intv ik, oc;
    int i, j, c, ncurr = 0, nprev;
    set_intv(ix, q[x], ik);
    ik.info = x + 1;
    oc.x0 = oc.x1 = oc.x2 = oc.info = 0;
    for (i = x + 1; i < len; ++i) {            // forward extension
        if (q[i] < 4) {
            c = 3 - q[i];
            extend_c(ix, ik, c, 0, oc, ctr);
            if (oc.x2 != ik.x2) {
                curr[ncurr++] = ik;
                if (oc.x2 < (u64)min_intv) break;
            }
            ik = oc; ik.info = i + 1;
        } else {
            curr[ncurr++] = ik;
            break;
        }
    }
    if (i == len) curr[ncurr++] = ik;
    reverse_(curr, ncurr);
    int ret = (int)curr[0].info;
    { Intv *t = curr; curr = prev; prev = t; }
    nprev = ncurr;
    int base = out.n;                          // reference builds `mem` back to front, then reverses
    for (i = x - 1; i >= -1; --i) {            // backward extension
        c = i < 0 ? -1 : q[i] < 4 ? q[i] : -1;
        ncurr = 0;
        for (j = 0; j < nprev; ++j) {
            Intv *p = &prev[j];
            if (c >= 0) extend_c(ix, *p, c, 1, oc, ctr);
            if (c < 0 || oc.x2 < (u64)min_intv) {
                if (ncurr == 0) {
                    if (out.n == base || (u64)(i + 1) < (out.a[out.n - 1].info >> 32)) {
                        ik = *p; ik.info |= (u64)(i + 1) << 32;
                        out.push(ik);
                        if (out.overflow) return ret;
                    }
                }
            } else if (ncurr == 0 || oc.x2 != curr[ncurr - 1].x2) {
                oc.info = p->info;
                curr[ncurr++] = oc;
            }
        }
        if (ncurr == 0) break;
        { Intv *t = curr; curr = prev; prev = t; }
        nprev = ncurr;
    }
    reverse_(out.a + base, out.n - base);
    *n_new = out.n - base;
    return ret;
}

template <class Ctr>
HD int seed_strategy1(const DevIndex &ix, int len, const u8 *q, int x, int min_len, int max_intv, Intv *mem, Ctr &ctr)
{
    Intv ik, oc;
    mem->x0 = mem->x1 = mem->x2 = mem->info = 0;
    if (q[x] > 3) return x + 1;
    set_intv(ix, q[x], ik);
    for (int i = x + 1; i < len; ++i) {
        if (q[i] < 4) {
            int c = 3 - q[i];
            extend_c(ix, ik, c, 0, oc, ctr);
            if (oc.x2 < (u64)max_intv && i - x >= min_len) {
                *mem = oc;
                mem->info = (u64)x << 32 | (u64)(i + 1);
                return i + 1;
            }
            ik = oc;
        } else return i + 1;
    }
    return len;
}

// keep only the entries of out[from, out.n) whose length is >= min_seed_len
HD void keep_long_(IntvSink &out, int from, int min_seed_len)
{
    int k = from;
    for (int i = from; i < out.n; ++i) {
        const Intv &p = out.a[i];
        int slen = (int)((u32)p.info - (u32)(p.info >> 32));
        if (slen >= min_seed_len) out.a[k++] = p;
    }
    out.n = k;
}

struct IntvLess { HD bool operator()(const Intv &a, const Intv &b) const { return a.info < b.info; } };

template <class Ctr>
HD void collect_intv(const DevIndex &ix, const Opt &opt, int len, const u8 *seq,
                     IntvSink &out, Intv *prev, Intv *curr, Ctr &ctr)
{
    int x = 0, n_new;
    int split_len = (int)(opt.min_seed_len * opt.split_factor + .499);
    out.n = 0;
    while (x < len) {                                   // pass 1: all SMEMs
        if (seq[x] < 4) {
            int from = out.n;
            x = smem1(ix, len, seq, x, 1, out, &n_new, prev, curr, ctr);
            if (out.overflow) return;
            keep_long_(out, from, opt.min_seed_len);
        } else ++x;
    }
    int old_n = out.n;                                  // pass 2: re-seed inside long SMEMs
    for (int k = 0; k < old_n; ++k) {
        Intv p = out.a[k];
        int start = (int)(p.info >> 32), end = (int)(i32)p.info;
        if (end - start < split_len || p.x2 > (u64)opt.split_width) continue;
        int from = out.n;
        smem1(ix, len, seq, (start + end) >> 1, (int)p.x2 + 1, out, &n_new, prev, curr, ctr);
        if (out.overflow) return;
        keep_long_(out, from, opt.min_seed_len);
    }
    if (opt.max_mem_intv > 0) {                         // pass 3: LAST-like
        x = 0;
        while (x < len) {
            if (seq[x] < 4) {
                Intv m;
                x = seed_strategy1(ix, len, seq, x, opt.min_seed_len, (int)opt.max_mem_intv, &m, ctr);
                if (m.x2 > 0) { out.push(m); if (out.overflow) return; }
            } else ++x;
        }
    }
    introsort((size_t)out.n, out.a, IntvLess());
}

} // namespace b200
