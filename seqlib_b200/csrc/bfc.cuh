// bfc.cuh -- k-mer counting keys, count-table lookups and per-read error correction of fermi-lite's BFC stage.
//
//   kmer_append / kmer_change / kmer_hash   <- bfc_kmer_append, bfc_kmer_change, bfc_hash_64, bfc_kmer_hash (fermi-lite/kmer.h:10-88)
//   kmer_key                                 <- get_subhash (fermi-lite/htab.c:45-58): WHICH k-mers share a counter
//   CountTable::get                          <- bfc_ch_get / bfc_ch_kmer_occ (fermi-lite/htab.c:85-102)
//   ec_kcov, ec_best_island, ec_greedy_k,
//   ec1dir, ec1                              <- bfc_ec_kcov, bfc_ec_best_island, bfc_ec_greedy_k, bfc_ec1dir, bfc_ec1 (fermi-lite/bfc.c:142-466)
//   max_streak / fltuniq1                    <- max_streak and the flt_uniq branch of worker_ec (fermi-lite/bfc.c:469-506)
//
// The reference keeps 2^l_pre khash tables keyed by (2k - l_pre) bits with the 14 count bits packed below the key.
// Only the equivalence "same sub-table and same stored key" and the saturating counts are observable, so the GPU
// table is a single open-addressing array of 16-byte slots keyed by that pair (KmerKey); counts come from a radix
// sort + run-length reduction (saturating adds commute), see fml.cu.  Everything here is HD code: the kernels call it
// per read, tests/hostsim runs it on the CPU against the reference library.
#pragma once
#include "common.cuh"

namespace b200 {

// seq_nt6_table (fermi-lite/misc.c:12-29) minus one: A,C,G,T -> 0..3, N/other -> 4
HD int nt6m1(unsigned char c)
{
    switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return 4;
    }
}

struct Kmer4 { u64 x[4]; };

HD void kmer_clear(Kmer4 &z) { z.x[0] = z.x[1] = z.x[2] = z.x[3] = 0; }

HD void kmer_append(int k, u64 x[4], int c)
{
    u64 mask = (1ull << k) - 1;
    x[0] = (x[0] << 1 | (u64)(c & 1)) & mask;
    x[1] = (x[1] << 1 | (u64)(c >> 1)) & mask;
    x[2] = x[2] >> 1 | (1ull ^ (u64)(c & 1)) << (k - 1);
    x[3] = x[3] >> 1 | (1ull ^ (u64)(c >> 1)) << (k - 1);
}

HD void kmer_change(int k, u64 x[4], int d, int c)
{
    u64 t = ~(1ull << d);
    x[0] = (u64)(c & 1) << d | (x[0] & t);
    x[1] = (u64)(c >> 1) << d | (x[1] & t);
    t = ~(1ull << (k - 1 - d));
    x[2] = (u64)(1 ^ (c & 1)) << (k - 1 - d) | (x[2] & t);
    x[3] = (u64)(1 ^ (c >> 1)) << (k - 1 - d) | (x[3] & t);
}

HD u64 bfc_hash64(u64 key, u64 mask)
{
    key = (~key + (key << 21)) & mask;
    key = key ^ key >> 24;
    key = ((key + (key << 3)) + (key << 8)) & mask;
    key = key ^ key >> 14;
    key = ((key + (key << 2)) + (key << 4)) & mask;
    key = key ^ key >> 28;
    key = (key + (key << 31)) & mask;
    return key;
}

// y[0], y[1] as bfc_kmer_hash leaves them in h[] (the pair handed to bfc_ch_insert / bfc_ch_get)
HD void kmer_hash(int k, const u64 x[4], u64 y[2])
{
    int t = k >> 1, u = ((x[1] >> t & 1) > (x[3] >> t & 1));
    u64 mask = (1ull << k) - 1;
    u64 h0 = bfc_hash64((x[u << 1 | 0] + x[u << 1 | 1]) & mask, mask);
    u64 h1 = bfc_hash64(h0 ^ x[u << 1 | 1], mask);
    y[0] = (h0 + h1) & mask;
    y[1] = h1;
}

struct KmerKey { u64 lo, hi; };

// bfc_ch_init's clamping of l_pre (fermi-lite/htab.c:20-33)
HD int bfc_l_pre(int k, int l_pre)
{
    if (k * 2 - l_pre > 50) l_pre = k * 2 - 50;
    if (l_pre > 20) l_pre = 20;
    return l_pre;
}

// (sub-table, stored key) of get_subhash.  k <= 32: the pair is a bijection of z = y0 << k | y1.  k > 32: the key is
// folded to 50 bits exactly as `(...) << 14` folds it in the reference.
HD KmerKey kmer_key(int k, int l_pre, const u64 y[2])
{
    KmerKey r;
    if (k <= 32) { r.lo = y[0] << k | y[1]; r.hi = 0; }
    else {
        int t = k - l_pre;
        int shift = t + k < 50 ? k : 50 - t;
        u64 key = (y[0] & ((1ull << t) - 1)) << shift ^ y[1];
        r.lo = (key << 14) >> 14;
        r.hi = y[0] >> t;
    }
    return r;
}

// Open-addressing table over 16-byte slots: w0 = key.lo, w1 = key.hi << 16 | 1 << 15 | count (14 bits); w1 == 0: empty.
struct CountSlot { u64 w0, w1; };

HD u64 count_slot_hash(const KmerKey &key)
{
    u64 h = key.lo * 0x9E3779B97F4A7C15ull;
    h ^= h >> 29; h += key.hi * 0xD6E8FEB86659FD93ull;
    h ^= h >> 32; h *= 0xBF58476D1CE4E5B9ull;
    h ^= h >> 31;
    return h;
}

struct CountTable {
    const CountSlot *slots; u64 mask;     // capacity - 1 (power of two)
    int k, l_pre;
    HD int get(const KmerKey &key) const
    {
        u64 i = count_slot_hash(key) & mask;
        for (;;) {
#if defined(__CUDA_ARCH__)
            const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2 *>(slots + i));
            const u64 w0 = v.x, w1 = v.y;
#else
            const u64 w0 = slots[i].w0, w1 = slots[i].w1;
#endif
            if (w1 == 0) return -1;
            if (w0 == key.lo && (w1 >> 16) == key.hi) return (int)(w1 & 0x3fff);
            i = (i + 1) & mask;
        }
    }
    HD int kmer_occ(const Kmer4 &z) const
    {
        u64 y[2];
        kmer_hash(k, z.x, y);
        return get(kmer_key(k, l_pre, y));
    }
};

// worker_count (fermi-lite/bfc.c:66-84): one record per k-mer of the read, in read order.  rec_hi = key.hi with the
// "all k bases have quality >= q" flag in bit 31.  Returns the number of records (lo / hi may be null to only count).
HD int count_read_kmers(int k, int l_pre, int q, const char *seq, const char *qual, int len, u64 *lo, u32 *hi)
{
    Kmer4 x; kmer_clear(x);
    const u64 mask = (1ull << k) - 1;
    u64 qmer = 0;
    int n = 0, l = 0;
    for (int i = 0; i < len; ++i) {
        int c = nt6m1((unsigned char)seq[i]);
        if (c < 4) {
            kmer_append(k, x.x, c);
            qmer = (qmer << 1 | (u64)(qual == 0 || qual[i] - 33 >= q)) & mask;
            if (++l >= k) {
                if (lo) {
                    u64 y[2];
                    kmer_hash(k, x.x, y);
                    KmerKey key = kmer_key(k, l_pre, y);
                    lo[n] = key.lo; hi[n] = (u32)key.hi | (qmer == mask ? 0x80000000u : 0u);
                }
                ++n;
            }
        } else { l = 0; qmer = 0; kmer_clear(x); }
    }
    return n;
}

// ---------------------------------------------------------------------------------------------- error correction
struct BfcOpt {      // bfc_opt_t (fermi-lite/bfc.h:44-57) as bfc_opt_init + fml_correct_core fill it
    int k, q, l_pre, min_cov, max_end_ext, win_multi_ec;
    float min_trim_frac;
    int w_ec, w_ec_high, w_absent, w_absent_high, max_path_diff, max_heap;
};

HD void bfc_opt_defaults(BfcOpt &o)
{
    o.k = -1; o.q = 20; o.l_pre = -1; o.min_cov = 4; o.max_end_ext = 5; o.win_multi_ec = 10; o.min_trim_frac = .8f;
    o.w_ec = 1; o.w_ec_high = 7; o.w_absent = 3; o.w_absent_high = 1; o.max_path_diff = 15; o.max_heap = 100;
}

struct EcBase {      // ecbase_t (fermi-lite/bfc.h:86-91) packed into 4 bytes; lcov / hcov are 6-bit counters in the reference
    u8 b : 3, q : 1, ob : 3, oq : 1;
    u8 lcov, hcov;
    u8 solid_end : 1, high_end : 1;
};

enum { BFC_EC_HIST = 5, BFC_EC_HIST_HIGH = 2, BFC_MAX_PATHS = 4 };
enum { ECCODE_OK = 0, ECCODE_MISC = 1, ECCODE_MANY_N = 2, ECCODE_NO_SOLID = 3, ECCODE_UNCORR_N = 4, ECCODE_MANY_FAIL = 5, ECCODE_SCRATCH = 6 };

struct EcPenalty { u8 ec : 1, ec_high : 1, absent : 1, absent_high : 1, b : 4; };     // bfc_penalty_t (fermi-lite/bfc.h:37-39)

struct EcHeap {      // echeap1_t
    int tot_pen, i, k;
    i32 ecpos_high[BFC_EC_HIST_HIGH], ecpos[BFC_EC_HIST];
    Kmer4 x;
};

struct EcStack {     // ecstack1_t
    int parent, i, tot_pen;
    u8 b; EcPenalty pen; u16 cnt;
};

struct EcHeapKey { int tot_pen, slot; };      // what the binary heap moves around; the 72-byte entry stays put in `pool`

struct EcScratch {
    EcBase *seq; u8 *ec0, *ec1;     // len entries each; of the two corrected copies only the base is ever read back
    EcHeapKey *heap; EcHeap *pool; u8 *free_; int heap_cap, heap_n, n_free;
    EcStack *stack; int stack_cap, stack_n;
    int heap_hw, stack_hw;          // high-water marks (scratch sizing)
    bool overflow;
};

HD size_t ec_scratch_bytes(int maxlen, int heap_cap, int stack_cap)
{
    return (((size_t)maxlen * (sizeof(EcBase) + 2) + 15) & ~(size_t)15) + (size_t)(heap_cap + 2) * sizeof(EcHeap)
         + (size_t)(heap_cap + 2) * sizeof(EcHeapKey) + (((size_t)heap_cap + 2 + 15) & ~(size_t)15) + (size_t)stack_cap * sizeof(EcStack);
}

HD void ec_scratch_bind(EcScratch &s, u8 *p, int maxlen, int heap_cap, int stack_cap)
{
    s.seq = (EcBase *)p; s.ec0 = (u8 *)(s.seq + maxlen); s.ec1 = s.ec0 + maxlen;     // p is 16-byte aligned
    p += ((size_t)maxlen * (sizeof(EcBase) + 2) + 15) & ~(size_t)15;
    s.pool = (EcHeap *)p; s.heap_cap = heap_cap; s.heap_n = 0;
    p += (size_t)(heap_cap + 2) * sizeof(EcHeap);
    s.heap = (EcHeapKey *)p;
    p += (size_t)(heap_cap + 2) * sizeof(EcHeapKey);
    s.free_ = p; s.n_free = 0;
    p += ((size_t)heap_cap + 2 + 15) & ~(size_t)15;
    s.stack = (EcStack *)p; s.stack_cap = stack_cap; s.stack_n = 0;
    s.heap_hw = s.stack_hw = 0;
    s.overflow = false;
}

HD int weighted_penalty(const BfcOpt &o, const EcPenalty &p)
{
    return o.w_ec * p.ec + o.w_ec_high * p.ec_high + o.w_absent * p.absent + o.w_absent_high * p.absent_high;
}

// ks_heapup_ec / ks_heapdown_ec (fermi-lite/ksort.h:125-146) with heap_lt(a, b) = a.tot_pen > b.tot_pen
HD void ec_heapup(int n, EcHeapKey *l)
{
    int k = n - 1;
    EcHeapKey tmp = l[k];
    while (k) {
        int i = (k - 1) >> 1;
        if (tmp.tot_pen > l[i].tot_pen) break;
        l[k] = l[i]; k = i;
    }
    l[k] = tmp;
}

HD void ec_heapdown(int i, int n, EcHeapKey *l)
{
    int k = i;
    EcHeapKey tmp = l[i];
    while ((k = (k << 1) + 1) < n) {
        if (k != n - 1 && l[k].tot_pen > l[k + 1].tot_pen) ++k;
        if (l[k].tot_pen > tmp.tot_pen) break;
        l[i] = l[k]; i = k;
    }
    l[i] = tmp;
}

// bfc_seq_conv (fermi-lite/bfc.c:101-116)
HD void ec_seq_conv(const char *s, const char *q, int l, int qthres, EcBase *a)
{
    for (int i = 0; i < l; ++i) {
        EcBase c;
        c.b = c.ob = (u8)nt6m1((unsigned char)s[i]);
        c.q = c.oq = !q ? 1 : (q[i] - 33 >= qthres ? 1 : 0);
        if (c.b > 3) c.q = c.oq = 0;
        c.lcov = c.hcov = 0; c.solid_end = c.high_end = 0;
        a[i] = c;
    }
}

HD EcBase ecbase_comp(const EcBase &b)
{
    EcBase r = b;
    r.b = b.b < 4 ? 3 - b.b : 4;
    r.ob = b.ob < 4 ? 3 - b.ob : 4;
    return r;
}

HD void ec_bases_revcomp(u8 *a, int n)
{
    int i;
    for (i = 0; i < n >> 1; ++i) {
        u8 tmp = a[i] < 4 ? 3 - a[i] : 4;
        a[i] = a[n - 1 - i] < 4 ? 3 - a[n - 1 - i] : 4;
        a[n - 1 - i] = tmp;
    }
    if (n & 1) a[i] = a[i] < 4 ? 3 - a[i] : 4;
}

HD void ec_seq_revcomp(EcBase *a, int n)
{
    int i;
    for (i = 0; i < n >> 1; ++i) {
        EcBase tmp = ecbase_comp(a[i]);
        a[i] = ecbase_comp(a[n - 1 - i]);
        a[n - 1 - i] = tmp;
    }
    if (n & 1) a[i] = ecbase_comp(a[i]);
}

template <class Tab>
HD int ec_greedy_k(int k, int mode, const Kmer4 &x, const Tab &ch)
{
    int max = 0, max_ec = -1, max2 = 0;
    for (int i = 0; i < k; ++i) {
        int c = (int)(x.x[1] >> i & 1) << 1 | (int)(x.x[0] >> i & 1);
        for (int j = 0; j < 4; ++j) {
            if (j == c) continue;
            Kmer4 y = x;
            kmer_change(k, y.x, i, j);
            int ret = ch.kmer_occ(y);
            if (ret < 0) continue;
            if ((max & 0xff) < (ret & 0xff)) { max2 = max; max = ret; max_ec = i << 2 | j; }
            else if ((max2 & 0xff) < (ret & 0xff)) max2 = ret;
        }
    }
    return (max & 0xff) * 3 > mode && (max2 & 0xff) < 3 ? max_ec : -1;
}

HD int ec_first_kmer(int k, const EcBase *a, int n, int start, Kmer4 &x)
{
    int i, l;
    kmer_clear(x);
    for (i = start, l = 0; i < n; ++i) {
        if (a[i].b < 4) {
            kmer_append(k, x.x, a[i].b);
            if (++l == k) break;
        } else { l = 0; kmer_clear(x); }
    }
    return i;
}

// bfc_ec_kcov (fermi-lite/bfc.c:175-196).  The reference bumps lcov / hcov of all k positions under every solid k-mer; the
// same 6-bit counters come out of one backward sweep with running window sums (positions covered by the k-mers that END in
// [j, j + k - 1]).
template <class Tab>
HD void ec_kcov(int k, int min_occ, EcBase *a, int n, const Tab &ch)
{
    Kmer4 x; kmer_clear(x);
    int l = 0;
    for (int i = 0; i < n; ++i) {
        EcBase &c = a[i];
        c.high_end = c.solid_end = 0; c.lcov = c.hcov = 0;
        if (c.b < 4) {
            kmer_append(k, x.x, c.b);
            if (++l >= k) {
                int r = ch.kmer_occ(x);
                if (r >= 0) {
                    if ((r >> 8 & 0x3f) >= min_occ + 1) c.high_end = 1;
                    if ((r & 0xff) >= min_occ) c.solid_end = 1;
                }
            }
        } else { l = 0; kmer_clear(x); }
    }
    int ls = 0, hs = 0;
    for (int j = n - 1; j >= 0; --j) {
        if (a[j].solid_end) { ++ls; hs += a[j].high_end; }
        if (j + k < n && a[j + k].solid_end) { --ls; hs -= a[j + k].high_end; }
        a[j].lcov = (u8)(ls & 63); a[j].hcov = (u8)(hs & 63);
    }
}

HD u64 ec_best_island(int k, const EcBase *a, int n)
{
    int i, l, max, max_i;
    for (i = k - 1, max = l = 0, max_i = -1; i < n; ++i) {
        if (!a[i].solid_end) {
            if (l > max) { max = l; max_i = i; }
            l = 0;
        } else ++l;
    }
    if (l > max) { max = l; max_i = i; }
    return max > 0 ? (u64)(max_i - max - k + 1) << 32 | (u32)max_i : 0;
}

// buf_update (fermi-lite/bfc.c:231-263)
HD void ec_buf_update(const BfcOpt &o, EcScratch &e, const EcHeap &prev, const EcPenalty &pen, int cnt)
{
    if (e.stack_n >= e.stack_cap || e.heap_n >= e.heap_cap || e.n_free == 0) { e.overflow = true; return; }
    EcStack &q = e.stack[e.stack_n++];
    q.parent = prev.k; q.i = prev.i; q.b = pen.b; q.pen = pen;
    q.cnt = cnt > 0 ? (u16)(cnt & 0xff) : 0;
    q.tot_pen = prev.tot_pen + weighted_penalty(o, pen);
    const int slot = e.free_[--e.n_free];
    EcHeap &r = e.pool[slot];
    r.i = prev.i + 1;
    r.k = e.stack_n - 1;
    r.x = prev.x;
    if (pen.ec_high) {
        for (int t = BFC_EC_HIST_HIGH - 1; t >= 1; --t) r.ecpos_high[t] = prev.ecpos_high[t - 1];
        r.ecpos_high[0] = prev.i;
    } else for (int t = 0; t < BFC_EC_HIST_HIGH; ++t) r.ecpos_high[t] = prev.ecpos_high[t];
    if (pen.ec) {
        for (int t = BFC_EC_HIST - 1; t >= 1; --t) r.ecpos[t] = prev.ecpos[t - 1];
        r.ecpos[0] = prev.i;
    } else for (int t = 0; t < BFC_EC_HIST; ++t) r.ecpos[t] = prev.ecpos[t];
    r.tot_pen = q.tot_pen;
    kmer_append(o.k, r.x.x, pen.b);
    EcHeapKey hk; hk.tot_pen = r.tot_pen; hk.slot = slot;
    e.heap[e.heap_n++] = hk;
    ec_heapup(e.heap_n, e.heap);
    if (e.heap_n > e.heap_hw) e.heap_hw = e.heap_n;
    if (e.stack_n > e.stack_hw) e.stack_hw = e.stack_n;
}

// buf_backtrack (fermi-lite/bfc.c:265-278)
HD int ec_backtrack(const EcStack *s, int end, int n, u8 *path)
{
    int n_absent = 0;
    while (end >= 0) {
        int i = s[end].i;
        if (i < n) {
            path[i] = s[end].b;
            n_absent += s[end].pen.absent;
        }
        end = s[end].parent;
    }
    return n_absent;
}

// bfc_ec1dir (fermi-lite/bfc.c:280-399): best-first search over corrections of seq[start+k .. end) and a short extension
template <class Tab>
HD int ec1dir(const BfcOpt &o, const Tab &ch, EcScratch &e, const EcBase *seq, int n, u8 *ec, int start, int end)
{
    int i, l, rv = -1, path[BFC_MAX_PATHS], n_paths = 0, min_path = -1, min_path_pen = 0x7fffffff, n_failures = 0;
    e.heap_n = e.stack_n = 0;
    e.n_free = e.heap_cap + 2;
    for (i = 0; i < e.n_free; ++i) e.free_[i] = (u8)(e.n_free - 1 - i);      // slot 0 is handed out first
    int zslot = e.free_[--e.n_free];
    EcHeap &z0 = e.pool[zslot];
    EcHeap z;
    z.tot_pen = 0; z.i = 0; z.k = 0; kmer_clear(z.x);
    for (z.i = start, l = 0; z.i < end; ++z.i) {
        int c = seq[z.i].b;
        if (c < 4) {
            if (++l == o.k) break;
            kmer_append(o.k, z.x.x, c);
        } else { l = 0; kmer_clear(z.x); }
    }
    z.k = -1;
    for (i = 0; i < BFC_EC_HIST; ++i) z.ecpos[i] = -1;
    for (i = 0; i < BFC_EC_HIST_HIGH; ++i) z.ecpos_high[i] = -1;
    z0 = z;
    { EcHeapKey hk; hk.tot_pen = 0; hk.slot = zslot; e.heap[e.heap_n++] = hk; }
    zslot = -1;
    for (i = 0; i < n; ++i) ec[i] = seq[i].b;
    for (;;) {
        int stop = 0;
        if (zslot >= 0) e.free_[e.n_free++] = (u8)zslot;        // the entry popped last round is no longer referenced
        zslot = -1;
        if (e.heap_n == 0) { rv = -2; break; }
        zslot = e.heap[0].slot;
        z = e.pool[zslot];
        e.heap[0] = e.heap[--e.heap_n];
        ec_heapdown(0, e.heap_n, e.heap);
        if (min_path >= 0 && z.tot_pen > min_path_pen + o.max_path_diff) break;
        if (z.i - end > o.max_end_ext) stop = 1;
        if (!stop) {
            const EcBase *c = z.i < n ? &seq[z.i] : 0;
            int b, os = -1, fixed = 0, other_ext = 0, n_added = 0, added_cnt[4];
            EcPenalty added[4];
            if (z.i > end) fixed = 1;
            if (c && c->b < 4) {
                Kmer4 x = z.x;
                kmer_append(o.k, x.x, c->b);
                os = ch.kmer_occ(x);
                if (c->q && (os & 0xff) >= o.min_cov + 1 && c->lcov >= o.min_cov + 1) fixed = 1;
                else if (c->hcov > o.k * .75) fixed = 1;
            }
            for (b = 0; b < 4; ++b) {
                EcPenalty pen;
                if (fixed && c && b != c->b) continue;
                if (c == 0 || b != c->b) {
                    Kmer4 x = z.x;
                    pen.ec = 0; pen.ec_high = 0; pen.absent = 0; pen.absent_high = 0; pen.b = b;
                    if (c) {
                        if (c->q && z.ecpos_high[BFC_EC_HIST_HIGH - 1] >= 0 && z.i - z.ecpos_high[BFC_EC_HIST_HIGH - 1] < o.win_multi_ec) continue;
                        if (z.ecpos[BFC_EC_HIST - 1] >= 0 && z.i - z.ecpos[BFC_EC_HIST - 1] < o.win_multi_ec) continue;
                    }
                    kmer_append(o.k, x.x, b);
                    int s = ch.kmer_occ(x);
                    if (s < 0 || (s & 0xff) < o.min_cov) continue;
                    pen.ec = c && c->b < 4 ? 1 : 0;
                    pen.ec_high = pen.ec ? c->oq : 0;
                    pen.absent = 0;
                    pen.absent_high = ((s >> 8 & 0xff) < o.min_cov);
                    pen.b = b;
                    added_cnt[n_added] = s;
                    added[n_added++] = pen;
                    ++other_ext;
                } else {
                    pen.ec = pen.ec_high = 0;
                    pen.absent = (os < 0 || (os & 0xff) < o.min_cov);
                    pen.absent_high = (os < 0 || (os >> 8 & 0xff) < o.min_cov);
                    pen.b = b;
                    added_cnt[n_added] = os;
                    added[n_added++] = pen;
                }
            }
            if (fixed == 0 && other_ext == 0) ++n_failures;
            if (n_failures > n * 2) { rv = -3; break; }
            if (c || n_added == 1) {
                if (n_added > 1 && e.heap_n > o.max_heap) {
                    int min_b = -1, min = 0x7fffffff;
                    for (b = 0; b < n_added; ++b) {
                        int t = weighted_penalty(o, added[b]);
                        if (min > t) { min = t; min_b = b; }
                    }
                    ec_buf_update(o, e, z, added[min_b], added_cnt[min_b]);
                } else {
                    for (b = 0; b < n_added; ++b) ec_buf_update(o, e, z, added[b], added_cnt[b]);
                }
                if (e.overflow) return -9;
            } else {
                if (n_added == 0) e.stack[z.k].tot_pen += o.w_absent * (o.max_end_ext - (z.i - end));
                stop = 1;
            }
        }
        if (stop) {
            if (e.stack[z.k].tot_pen < min_path_pen) { min_path_pen = e.stack[z.k].tot_pen; min_path = n_paths; }
            path[n_paths++] = z.k;
            if (n_paths == BFC_MAX_PATHS) break;
        }
    }
    if (n_paths == 0) return rv;
    rv = ec_backtrack(e.stack, path[min_path], n, ec);
    for (i = 0; i < n; ++i)
        if (i < start + o.k || i >= end) ec[i] = 4;
    return rv;
}

// bfc_ec1 (fermi-lite/bfc.c:401-466): corrects seq / qual (len bytes, no terminator needed) in place; returns the ec_code
template <class Tab>
HD int ec1(const BfcOpt &o, const Tab &ch, int mode, char *seq, char *qual, int len, EcScratch &e)
{
    int i, start = 0, end = 0, n_n = 0, rv0, rv1;
    EcBase *a = e.seq;
    ec_seq_conv(seq, qual, len, o.q, a);
    for (i = 0; i < len; ++i) if (a[i].ob > 3) ++n_n;
    if (n_n > len * .05) return ECCODE_MANY_N;
    ec_kcov(o.k, o.min_cov, a, len, ch);
    u64 r = ec_best_island(o.k, a, len);
    if (r == 0) {
        Kmer4 x;
        int ec = -1;
        while ((end = ec_first_kmer(o.k, a, len, start, x)) < len) {
            ec = ec_greedy_k(o.k, mode, x, ch);
            if (ec >= 0) break;
            if (end + (o.k >> 1) >= len) break;
            start = end - (o.k >> 1);
        }
        if (ec >= 0) {
            a[end - (ec >> 2)].b = ec & 3;
            ++end; start = end - o.k;
        } else return ECCODE_NO_SOLID;
    } else { start = (int)(r >> 32); end = (int)(u32)r; }
    if ((rv0 = ec1dir(o, ch, e, a, len, e.ec0, start, len)) < 0)
        return rv0 == -9 ? ECCODE_SCRATCH : rv0 == -2 ? ECCODE_UNCORR_N : rv0 == -3 ? ECCODE_MANY_FAIL : ECCODE_MISC;
    ec_seq_revcomp(a, len);
    if ((rv1 = ec1dir(o, ch, e, a, len, e.ec1, len - end, len)) < 0)
        return rv1 == -9 ? ECCODE_SCRATCH : rv1 == -2 ? ECCODE_UNCORR_N : rv1 == -3 ? ECCODE_MANY_FAIL : ECCODE_MISC;
    ec_bases_revcomp(e.ec1, len);
    ec_seq_revcomp(a, len);
    for (i = 0; i < len; ++i) {
        EcBase &c = a[i];
        if (e.ec0[i] == e.ec1[i]) c.b = e.ec0[i] > 3 ? a[i].b : e.ec0[i];
        else if (e.ec1[i] > 3) c.b = e.ec0[i];
        else if (e.ec0[i] > 3) c.b = e.ec1[i];
        else c.b = a[i].ob;
    }
    for (i = 0; i < len; ++i) {
        int is_diff = !(a[i].b == a[i].ob);
        seq[i] = (is_diff ? "acgtn" : "ACGTN")[a[i].b];
        if (qual) qual[i] = is_diff ? (char)(34 + a[i].ob) : "+?"[a[i].q];
    }
    return ECCODE_OK;
}

// max_streak (fermi-lite/bfc.c:469-488)
template <class Tab>
HD u64 max_streak(int k, const Tab &ch, const char *seq, int len)
{
    int i, l;
    u64 max = 0, t = 0;
    Kmer4 x; kmer_clear(x);
    for (i = l = 0; i < len; ++i) {
        int c = nt6m1((unsigned char)seq[i]);
        if (c < 4) {
            kmer_append(k, x.x, c);
            if (++l >= k) {
                if (ch.kmer_occ(x) > 0) t += 1ull << 32;
                else t = i + 1;
            } else t = i + 1;
        } else { l = 0; kmer_clear(x); t = i + 1; }
        max = max > t ? max : t;
    }
    return max;
}

// the flt_uniq branch of worker_ec (fermi-lite/bfc.c:492-506): trims the read in place, returns the new length (0 = dropped)
template <class Tab>
HD int fltuniq1(const BfcOpt &o, const Tab &ch, char *seq, char *qual, int len)
{
    u64 max = max_streak(o.k, ch, seq, len);
    if (max >> 32 && (double)((max >> 32) + o.k - 1) / len > o.min_trim_frac) {
        int start = (int)(u32)max, end = start + (int)(max >> 32);
        start -= o.k - 1;
        int n = end - start;
        if (start > 0) {
            for (int i = 0; i < n; ++i) seq[i] = seq[start + i];
            if (qual) for (int i = 0; i < n; ++i) qual[i] = qual[start + i];
        }
        return n;
    }
    return 0;
}

} // namespace b200
