// seed2.cuh -- mem_collect_intv (bwa/bwamem.c:140-188) shaped for the GPU: one gather site.
//
// seed.cuh states SMEM seeding as the reference does (bwt_smem1a / bwt_seed_strategy1, bwa/bwt.c:289-379): nested
// loops with bwt_extend in three places, two work lists of 32-byte intervals per read in scratch memory.  Profiled on
// a B200 that shape is bound by instruction issue (a third of the lanes active) and by scratch traffic, not by the
// Occ gathers.  Here the same computation is a small machine that yields every memory access it depends on:
//
//     request(...)  ->  gather (two 256-bit loads)  ->  consume / consume_chain / consume_aux
//
// so a warp has ONE place where the dependent gathers happen, reached by all lanes together.  The work list of a
// bwt_smem1a call is a bit mask (strings shorter than K bases) plus a small ring of packed intervals in the caller's
// shared memory (see SeedMachine).  SMEMs shorter than min_seed_len are dropped when they are
// produced (the reference drops them after each bwt_smem1 call; the "is this SMEM contained in the previous one"
// test uses the unfiltered start, kept in a register).  Intervals leave in production order; the final order is
// the sort by (start, end), done by the caller -- equal keys mean the same substring and therefore identical
// intervals, so the reference's unstable introsort cannot order them differently in any visible way.
//
// Restrictions (the caller routes everything else to seed.cuh): no base > 3 in the read, len < 65536,
// seq_len < 2^36, and the list capacity.
//
// Prefix-chain table (SeedTab).  The bi-interval of a string is a pure function of the index, and so is the size of the
// interval of each of its prefixes.  For every K-mer W the table holds ONE 32-byte entry: the packed bi-interval of W,
// min(size, 255) of the interval of every proper prefix W[0, m), m = 1..K-1, and a mask whose bit m says "the size of
// the (m+1)-prefix differs from the size of the m-prefix".  That is everything bwt_smem1a (bwa/bwt.c:289-351) asks of
// strings of at most K bases: it compares sizes of nested strings for equality (the mask), compares sizes with small
// thresholds (min_intv <= split_width + 1, max_mem_intv; saturation at 255 is exact for thresholds <= 255), and it
// needs coordinates only of what it reports (>= min_seed_len > K bases) or extends beyond K bases (the K-mer itself).
// So ONE gather of the entry of q[sx, sx+K) replaces the first K - 1 extensions of a forward sweep, and ONE gather of
// the entry of q[i, i+K) answers row i of the backward sweep for every list entry that ends within K bases of i --
// the triangle of ~K^2/2 short extensions per bwt_smem1a call becomes ~K gathers.  List entries shorter than K bases
// carry only their end.  The entries are produced from the same extend_lean code (k_chain_build, engine.cu), so
// every number is exactly what the iterated bwt_extend (bwa/bwt.c:262-275) computes and interval lists stay
// identical to the reference's.  K = 15: 4^15 x 32 B = 34 GB next to the 52 GB image of a 3 Gb index.
// The caller switches the table off (K = 0, Occ blocks only) when min_seed_len <= K or a threshold exceeds 255.
//
// Text path.  Once a forward sweep of the first pass (min_intv == 1) holds an interval of size ONE, every further
// extension succeeds iff the next read base equals the text base behind that single occurrence (the FM-index is the index
// of the forward + reverse-complement text, ix.text).  So the machine reads the occurrence's position from the suffix
// array (one gather), compares the read with the text word by word (one or two sectors) and gets the end of the sweep --
// instead of one dependent Occ gather per base.  The same holds for the backward sweep: the size-one entry is the longest
// of the list, it stays size one until the text before the occurrence differs from the read, so the number of rows it
// survives is one more comparison; the rows themselves then need no gather for it.  Such an interval is reported with
// x0 = INTV_TEXT_FLAG | position: the chain stage, the only consumer of x0 (bwt_sa of it, bwa/bwamem.c:279), takes the
// position as it is.  x1 of a tracked interval is not maintained (nothing reads it); b200_debug_collect_intv runs
// with the text path off and returns the reference's coordinates.  Needs the full suffix array (sa_shift == 0) and a text
// that IS the text of the BWT (SeqLib's ConstructIndex builds the BWT over a second randomisation of N bases: the
// engine verifies BWT[k] == text[SA[k] - 1] for all k before it sets SeedTab::text).
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

struct alignas(32) ChainEnt { u32 w0, w1, w2, w3; u8 sz[16]; };   // w0..w3 as PIntv with w3 >> 16 = change mask; sz[m] = min(size of the m-prefix, 255)
struct SeedTab { const ChainEnt *base; int K; int text; };     // K == 0: no table; text != 0: the text path may be used (see SeedMachine)
HD u64 seedtab_level_off(int j) { return ((1ull << (2 * j)) - 4) / 3; }      // entries of levels 1..j-1 of the builder's scratch (a multiple of 4)
HD u64 seedtab_entries(int K) { return K > 0 ? seedtab_level_off(K + 1) : 0; }
HD bool seedtab_opt_ok(int K, int min_seed_len, int split_width, i64 max_mem_intv)
{
    return K >= 4 && K <= 16 && min_seed_len > K && split_width >= 0 && split_width < 254 && max_mem_intv <= 255;
}

// a loaded chain entry
struct ChainView {
    u64 x0, x1, x2, szlo, szhi; u32 mask;
    HD u32 size_sat(int m, int K) const       // min(size of the m-prefix, 255), 1 <= m <= K
    {
        if (m >= K) return x2 < 255 ? (u32)x2 : 255u;
        return (u32)(((m < 8 ? szlo : szhi) >> (8 * (m & 7))) & 255u);
    }
    HD bool same(int a, int b) const { return ((mask >> a) & ((1u << (b - a)) - 1u)) == 0; }   // sizes of the a- and b-prefix equal (a <= b <= K)
    // number of prefix lengths 1..K whose size is >= t (1 <= t <= 255; sizes never grow with the length, byte 0 and bytes >= K are 0)
    HD int lmax(u32 t, int K) const
    {
        int n = x2 >= (u64)t ? 1 : 0;
#if defined(__CUDA_ARCH__)
        const u32 tt = t * 0x01010101u;       // __vcmpgeu4: 0xff in every byte lane where a >= b
        n += (__popc(__vcmpgeu4((u32)szlo, tt)) + __popc(__vcmpgeu4((u32)(szlo >> 32), tt)) +
              __popc(__vcmpgeu4((u32)szhi, tt)) + __popc(__vcmpgeu4((u32)(szhi >> 32), tt))) >> 3;
#else
        for (int m = 1; m < K; ++m) n += size_sat(m, K) >= t;
#endif
        return n;
    }
};
HD ChainView chain_view(const OccLoad &b)
{
    ChainView v; PIntv t; t.w0 = b.c0; t.w1 = b.c1; t.w2 = b.c2; t.w3 = b.c3;
    u32 e_; pintv_unpack(t, v.x0, v.x1, v.x2, e_);
    v.mask = e_; v.szlo = b.s0; v.szhi = b.s1;
    return v;
}
// entry of a K-mer from the sizes of its prefixes (s[1..K]) and its interval
HD ChainEnt chain_make(int K, u64 x0, u64 x1, u64 x2, const u64 *s)
{
    ChainEnt e; u32 mask = 0;
    for (int m = 0; m < 16; ++m) e.sz[m] = 0;
    for (int m = 1; m < K; ++m) {
        e.sz[m] = (u8)(s[m] < 255 ? s[m] : 255);
        if (s[m + 1] != s[m]) mask |= 1u << m;
    }
    PIntv p = pintv_pack(x0, x1, x2, mask);
    e.w0 = p.w0; e.w1 = p.w1; e.w2 = p.w2; e.w3 = p.w3;
    return e;
}

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

// The one gather site: the two Occ blocks of an extension (kind 0), the 32-byte chain entry of a K-mer (kind 1; loaded twice,
// the second request merges in L1), or two sectors the machine will read next (kind 2: a suffix-array entry, text around a
// position -- the loads pull them into L1, the values are picked up by consume_aux), so that a warp issues ONE pair of
// 256-bit loads whatever its lanes need.
template <class Ctr>
HD void gather(const DevIndex &ix, const SeedTab &tab, int kind, u32 key, const void *ga, const void *gb, u64 a, u64 o, u64 s, int c,
               u64 &na, u64 &no, u64 &ns, ChainView &cv, Ctr &ctr)
{
    const u64 k = a - 1, l = k + s;
    const u32 dk = k >= ix.primary, dl = l >= ix.primary;
    const u64 kk = k - dk, ll = l - dl;
    u64 bk = kk >> 6, bl = ll >> 6;
    const OccBlock *p1 = ix.occ + bk, *p2 = ix.occ + bl;
    if (kind == 1) { p1 = p2 = (const OccBlock *)(tab.base + key); ctr.tab_hi++; }
    else if (kind == 2) { p1 = (const OccBlock *)ga; p2 = (const OccBlock *)gb; ctr.tab_lo++; }
    else ctr.occ_blocks += bl != bk ? 2 : 1;
    const OccLoad b1 = load_block_at(p1), b2 = load_block_at(p2);
    if (kind == 1) cv = chain_view(b1);
    else if (kind == 0) extend_blocks(ix, b1, b2, kk, ll, dk, dl, o, c, na, no, ns);
#if defined(__CUDA_ARCH__)
    else asm volatile("" :: "r"(b1.c0), "r"(b2.c0));      // keep the prefetching loads
#endif
}

HD u64 ld_u64(const u64 *p)
{
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}
// 16 bases of the text from position P on, base t in bits 2t.. (reads the word after P's as well: the text has a padding word)
HD u32 text_bits32(const DevIndex &ix, u64 P)
{
    const u64 w = ld_u64(ix.text + (P >> 5)), w2 = ld_u64(ix.text + (P >> 5) + 1);
    const int sh = (int)(P & 31) * 2;
    return (u32)(sh ? (w >> sh) | (w2 << (64 - sh)) : w);
}
HD int ctz32(u32 x)
{
#if defined(__CUDA_ARCH__)
    return __ffs((int)x) - 1;
#else
    return __builtin_ctz(x);
#endif
}
HD int clz32(u32 x)
{
#if defined(__CUDA_ARCH__)
    return __clz((int)x);
#else
    return __builtin_clz(x);
#endif
}
static const u64 INTV_TEXT_FLAG = 1ull << 63;      // Intv::x0 with this bit: the low bits are the text position of the (single) occurrence

// One read's seeding as a resumable machine.  Query: operator[](i) in 0..3, key(st, ln).
// Work list of a bwt_smem1a call (the reference's prev / curr arrays, bwa/bwt.c:303-349), in two parts:
//   * "long" entries -- strings of at least K bases, the only ones whose interval is ever needed -- in the caller's List
//     (get / put of packed intervals; `cap` slots, a power of two, used as a ring: the forward sweep fills it downwards, so
//     the list is born longest-first as the reference wants it after its in-place reversal; a backward row compacts it in
//     place and may append one entry, the string that just reached K bases);
//   * "short" entries -- strings of fewer than K bases that start at sx: only their ends matter, and those are a bit mask
//     (bit m: the entry that ends at sx + m).  A backward row turns that mask into the next one with a dozen bit operations
//     on the row's chain entry (sizes are monotone in the length, so "alive" is a prefix of the mask and "size differs from
//     the last kept entry" is "a change bit lies between this entry and the next longer one").
template <class List, class Query>
struct SeedMachine {
    enum { M_DONE = 0, M_FWD, M_BWD, M_P3, M_FWDC, M_FSA, M_FTX, M_BTX, M_TASK, M_ENDFWD, M_LASTROW };
    int mode, pass, x, k2, old_n, sx, i, j, nlong, ncurr, ret, last_start, first, ovf;
    int len, cap, min_seed_len, split_len, split_width, max_intv3, min_intv, K, last_pass;
    u32 top, smask;               // ring position of the longest long entry; short entries
    u64 x0, x1, x2, lastcurr;     // ik of the forward sweeps / last size kept in this backward row
    u64 p0, p1, p2;               // the list entry being extended backwards
    u32 iend, pend;
    u64 tpos; int tleft, textok; bool tracked;      // text path: position of q[sx] of the tracked (size-one, longest) entry, rows it survives
    List L; Query q; IntvSink out;

    // last_pass_ = 2: stop after the SMEM passes (bwt_smem1a calls, re-seeding); start3() then runs the third pass
    // (bwt_seed_strategy1) on its own -- k_seed2 / k_seed3 split them so that a warp never mixes the two kinds of lanes
    HD void init(const Opt &opt, int len_, int cap_, const List &L_, const Query &q_, const IntvSink &out_, int K_ = 0, int last_pass_ = 3, int textok_ = 0)
    {
        len = len_; cap = cap_; L = L_; q = q_; out = out_; K = K_; last_pass = last_pass_; textok = textok_; tracked = false; tleft = 0; tpos = 0;
        min_seed_len = opt.min_seed_len;
        split_len = (int)(opt.min_seed_len * opt.split_factor + .499);
        split_width = opt.split_width;
        max_intv3 = (int)opt.max_mem_intv;
        pass = 1; x = 0; k2 = 0; old_n = 0; ovf = 0; mode = M_DONE; nlong = 0; smask = 0; top = 1u << 20;
        out.n = 0; out.overflow = false;
    }

    HD int slot(u32 e) const { return (int)(e & (u32)(cap - 1)); }
    HD void fail() { ovf = 1; mode = M_DONE; }

    HD void emit(u64 e0, u64 e1, u64 e2, u64 info)
    {
        Intv v; v.x0 = e0; v.x1 = e1; v.x2 = e2; v.info = info;
        out.push(v);
        if (out.overflow) fail();
    }

    HD void push_fwd()             // the forward sweep lists ik (a long entry)
    {
        if (nlong == cap) { fail(); return; }
        L.put(slot(--top), x0, x1, x2, iend);
        ++nlong;
        ret = (int)iend;
    }

    // the end of a backward row (bwa/bwt.c:346-348)
    HD void row_end()
    {
        if (ncurr == 0 && smask == 0) { mode = M_TASK; if (pass == 1) x = ret; }
        else {
            nlong = ncurr; ncurr = 0; j = 0;
            if (--i < 0) mode = M_LASTROW;
        }
    }

    // Everything between two gathers that is not the per-extension bookkeeping: ending a forward sweep, the row
    // i == -1 of a backward sweep, picking the next bwt_smem1 / bwt_seed_strategy1 call (bwa/bwamem.c:150-184).
    // A loop over transient modes instead of mutually recursive helpers, so that it inlines and the state stays in registers.
    HD void settle(const DevIndex &ix)
    {
        for (;;) {
            if (mode == M_ENDFWD) {
                ncurr = 0; j = 0; first = 1; last_start = 0;
                i = sx - 1;
                mode = i < 0 ? M_LASTROW : tracked ? M_BTX : M_BWD;
            } else if (mode == M_LASTROW) {
                // bwa/bwt.c:325-345 with c = -1: only the longest survivor can be reported, with start 0 (a short entry never:
                // it has fewer than K < min_seed_len bases)
                if (nlong > 0 && (first || 0 < last_start)) {
                    u64 e0, e1, e2; u32 e;
                    if (tracked) {
                        e = L.end(slot(top));
                        if ((int)e >= min_seed_len) { emit(INTV_TEXT_FLAG | (tpos - (u64)sx), 0, 1, (u64)e); if (ovf) return; }
                    } else {
                        const bool ok = L.take(slot(top), e0, e1, e2, e);
                        if ((int)e >= min_seed_len) { if (!ok) { fail(); return; } emit(e0, e1, e2, (u64)e); if (ovf) return; }
                    }
                }
                tracked = false;
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
                    // the first K - 1 extensions of a bwt_seed_strategy1 start (bwa/bwt.c:355-379) can neither report nor
                    // stop (i - x < min_seed_len): the interval of q[x, x + K) in the chain entry replaces them
                    // (K < min_seed_len: with fewer than K bases left nothing can be reported any more)
                    if (K > 0) {
                        if (x + K > len) { mode = M_DONE; return; }
                        x0 = x1 = 1; x2 = 0;             // placeholders: the lookup result overwrites them
                        i = x + K - 1;
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
                tracked = false; nlong = 0; smask = 0;
                if (K > 0) { mode = M_FWDC; return; }       // through the chain entry of q[sx, sx + K)
                Intv t; set_intv(ix, q[sx], t);
                x0 = t.x0; x1 = t.x1; x2 = t.x2; iend = (u32)(sx + 1);
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

    // Row i of a backward sweep starts with the tracked entry: no gather, the text comparison already said how long it lives.
    HD void tracked_top()
    {
        if (tleft > 0) {
            if (nlong == 1 && smask == 0) {          // alone: all the rows until it dies or reaches the read's start
                const int d = tleft < i + 1 ? tleft : i + 1;
                tleft -= d; i -= d;
                if (i < 0) mode = M_LASTROW;
                return;
            }
            --tleft; ncurr = 1; lastcurr = 1; j = 1;      // kept in the first slot, where it is
        } else {
            tracked = false;
            const u32 e = L.end(slot(top));
            if (first || i + 1 < last_start) {
                first = 0; last_start = i + 1;
                if ((int)e - (i + 1) >= min_seed_len) {
                    emit(INTV_TEXT_FLAG | (tpos - (u64)(sx - (i + 1))), 0, 1, (u64)(i + 1) << 32 | e);
                    if (ovf) return;
                }
            }
            j = 1;
        }
        if (j == nlong && smask == 0) row_end();
    }

    // The next gather.  Returns its kind: 0 = an extension through the Occ blocks (coordinate a, other coordinate o, size s,
    // base c); 1 = the chain entry of the K-mer `key` -- the start of a forward sweep, the jump of a third-pass start, or the
    // short entries of a backward row; 2 = the sectors at ga / gb (text path); -1 = the read finished meanwhile.
    // the gather-free part of the next step: rows that begin with the tracked entry
    HD void pre(const DevIndex &ix)
    {
        while (mode == M_BWD && j == 0 && tracked) {
            tracked_top();
            if (mode > M_BTX) settle(ix);
        }
    }
    HD int request(const DevIndex &ix, u64 &a, u64 &o, u64 &s, int &c, u32 &key, const void *&ga, const void *&gb)
    {
        pre(ix);
        if (mode == M_DONE) return -1;
        key = 0u; a = 1; o = 1; s = 0; c = 0; ga = gb = ix.occ;
        if (mode == M_BWD) {
            if (j == nlong) { key = q.key(i, K); return 1; }          // the short entries: q[i, end) of at most K bases
            c = q[i];
            if (!L.take(slot(top + (u32)j), p0, p1, p2, pend)) { p0 = p1 = 1; p2 = 0; fail(); }      // the list did not keep this interval: spill path
            a = p0; o = p1; s = p2;
            return 0;
        }
        if (mode == M_FWDC) { key = q.key(sx, K); return 1; }
        if (mode >= M_FSA) {        // M_FSA, M_FTX, M_BTX
            const u8 *t0 = (const u8 *)ix.text;
            if (mode == M_FSA) ga = gb = (const u8 *)ix.sa + ((x0 * 8) & ~31ull);
            else if (mode == M_FTX) {
                const u64 sec = (tpos + (u64)(i - sx)) >> 7, last = (ix.seq_len + 31) >> 7;     // the text has (seq_len + 31) / 32 + 1 words
                ga = t0 + (sec << 5); gb = t0 + ((sec < last ? sec + 1 : sec) << 5);
            } else {
                const u64 sec = (tpos ? tpos - 1 : 0) >> 7;
                ga = t0 + (sec << 5); gb = t0 + ((sec ? sec - 1 : 0) << 5);
            }
            return 2;
        }
        a = x1; o = x0; s = x2; c = 3 - q[i];
        if (mode == M_P3 && K > 0 && i + 1 - x == K) { key = q.key(x, K); return 1; }
        return 0;
    }

    // text path: the prefetched sectors are in L1, the machine reads what it needs from them
    HD void consume_aux(const DevIndex &ix)
    {
        if (mode == M_FSA) { tpos = ld_u64(ix.sa + x0); tracked = true; mode = M_FTX; return; }
        if (mode == M_FTX) {
            // the interval of q[sx, i) has size one and its occurrence starts at tpos: q[i, len) against the text behind it
            // (bwa/bwt.c:303-320 with ok[c].x[2] in {0, 1}: the sweep ends at the first difference, the text's or the read's end)
            const u64 P = tpos + (u64)(i - sx);
            int n = len - i;
            if ((u64)n > ix.seq_len - P) n = (int)(ix.seq_len - P);
            int t = 0;
            while (t < n) {
                const int cnt = n - t < 16 ? n - t : 16;
                u32 d = text_bits32(ix, P + (u64)t) ^ q.key(i + t, 16);
                if (cnt < 16) d &= (1u << (2 * cnt)) - 1u;
                if (d) { t += ctz32(d) >> 1; break; }
                t += cnt;
            }
            i += t; iend = (u32)i; x2 = 1;
            push_fwd();
            if (ovf) return;
            mode = M_ENDFWD;
            settle(ix);
            return;
        }
        // M_BTX: q[sx-1], q[sx-2], ... against the text before the occurrence = the number of rows the tracked entry survives
        int n = sx;
        if ((u64)n > tpos) n = (int)tpos;
        int t = 0;
        while (t < n) {
            const int cnt = n - t < 16 ? n - t : 16;
            u32 d = text_bits32(ix, tpos - (u64)(t + cnt)) ^ q.key(sx - t - cnt, 16);
            d <<= 32 - 2 * cnt;                 // base cnt-1 of the chunk (the one next to what matched so far) on top
            if (d) { t += clz32(d) >> 1; break; }
            t += cnt;
        }
        tleft = t;
        mode = M_BWD;
    }

    // A chain entry arrives: the whole forward prefix (M_FWDC), the short entries of a backward row (M_BWD), or a third-pass jump.
    HD void consume_chain(const DevIndex &ix, const ChainView &E)
    {
        if (mode == M_P3) { consume(ix, E.x1, E.x0, E.x2); return; }
        const int lmax = E.lmax((u32)min_intv, K);               // prefixes of up to lmax bases have >= min_intv occurrences
        if (mode == M_FWDC) {
            // bwa/bwt.c:303-320 for the strings q[sx, sx + m), m = 1 .. avail: the m-prefix is listed when the (m+1)-prefix has
            // another size (change bit m), and the sweep stops there when that size is below min_intv, i.e. m >= lmax
            const int avail = len - sx < K ? len - sx : K;
            const u32 cm = E.mask & ((1u << avail) - 2u);        // change bits m = 1 .. avail-1
            const u32 st = cm & ~((1u << (lmax > 1 ? lmax : 1)) - 1u);      // change bits m >= max(lmax, 1): the first one stops the sweep
            if (st) {
                const int ms = ctz32(st);
                smask = cm & ((2u << ms) - 1u);
                ret = sx + ms;
                mode = M_ENDFWD;
            } else {
                smask = cm;
                if (sx + avail == len) {                       // ran into the read's end: the last string is listed
                    if (avail == K) { x0 = E.x0; x1 = E.x1; x2 = E.x2; iend = (u32)len; push_fwd(); if (ovf) return; }
                    else { smask |= 1u << avail; ret = len; }
                    mode = M_ENDFWD;
                } else {                                       // avail == K: on with the Occ blocks
                    x0 = E.x0; x1 = E.x1; x2 = E.x2; iend = (u32)(sx + K);
                    i = sx + K;
                    if (cm) ret = sx + 31 - clz32(cm);
                    mode = textok && x2 == 1 && min_intv == 1 ? M_FSA : M_FWD;
                }
            }
            if (mode == M_ENDFWD) settle(ix);
            return;
        }
        // M_BWD, row i: the short entry that ends at sx + m becomes q[i, sx + m) of m + d bases, d = sx - i (bwa/bwt.c:325-345).
        // Longest first: dead ones (more than lmax bases) come before alive ones.
        const int d = sx - i;
        const u32 A = lmax > d ? smask & ((2u << (lmax - d)) - 1u) : 0u;
        if ((smask & ~A) && ncurr == 0 && (first || i + 1 < last_start)) { first = 0; last_start = i + 1; }     // (too short to be reported)
        u32 kept = 0;
        if (A) {
            const int tm = 31 - clz32(A);                       // the longest alive short entry: compared with the last long entry kept
            const bool keep_top = ncurr == 0 || !(lastcurr == E.x2 && E.same(tm + d, K));
            // another one is kept iff its size differs from the next longer alive entry's: a change bit in between.  Every change
            // bit below the top entry marks the highest alive entry at or below it (occluded fill downwards, alive entries block).
            const u32 X = (E.mask >> d) & ((1u << tm) - 1u);
            u32 gen = X & ~A, pro = ~A;
            gen |= pro & (gen >> 1); pro &= pro >> 1;
            gen |= pro & (gen >> 2); pro &= pro >> 2;
            gen |= pro & (gen >> 4); pro &= pro >> 4;
            gen |= pro & (gen >> 8);
            kept = (A & (X | (gen >> 1))) & ~(1u << tm);
            if (keep_top) {
                if (tm + d == K) {                              // it reached K bases: a long entry from now on
                    if (ncurr == cap) { fail(); return; }
                    L.put(slot(top + (u32)ncurr), E.x0, E.x1, E.x2, (u32)(sx + tm));
                    ++ncurr;
                } else kept |= 1u << tm;
            }
        }
        smask = kept;
        row_end();
        if (mode > M_BTX) settle(ix);
    }

    HD void consume(const DevIndex &ix, u64 na, u64 no, u64 ns)
    {
        if (mode == M_BWD) {
            if (ns < (u64)min_intv) {
                if (ncurr == 0 && (first || i + 1 < last_start)) {
                    first = 0; last_start = i + 1;
                    if ((int)pend - (i + 1) >= min_seed_len) {
                        emit(p0, p1, p2, (u64)(i + 1) << 32 | pend);
                        if (ovf) return;
                    }
                }
            } else if (ncurr == 0 || ns != lastcurr) {
                L.put(slot(top + (u32)ncurr), na, no, ns, pend);
                ++ncurr; lastcurr = ns;
            }
            if (++j == nlong && smask == 0) row_end();
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
                else if (textok && ns == 1 && min_intv == 1) mode = M_FSA;
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
        if (mode > M_BTX) settle(ix);
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
    const u8 *p; int n;
    HD int operator[](int i) const { return p[i]; }
    // bases [st, st + ln), base t in bits 2t..; positions beyond the read count as base 0
    HD u32 key(int st, int ln) const { u32 k = 0; for (int t = 0; t < ln && st + t < n; ++t) k |= (u32)p[st + t] << (2 * t); return k; }
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
    ByteQuery q; q.p = seq; q.n = len;
    SeedTab none; none.base = nullptr; none.K = 0;
    const SeedTab &T = tab && seedtab_opt_ok(tab->K, opt.min_seed_len, opt.split_width, opt.max_mem_intv) ? *tab : none;
    SeedMachine<ArrayList, ByteQuery> m;
    for (int part = 0; part < 2; ++part) {             // the SMEM passes, then bwt_seed_strategy1: two machines, like k_seed2 + k_seed3
        const int textok = tab && tab->text && ix.sa_shift == 0;
        if (part == 0) { m.init(opt, len, cap, L, q, out, T.K, 2, textok); m.start(ix); }
        else { const int n0 = m.out.n; IntvSink o2 = m.out; m.init(opt, len, cap, L, q, o2, T.K, 3, textok); m.start3(ix, n0); }
        while (m.mode != 0) {
            u64 a, o, s, na = 0, no = 0, ns = 0; int c; u32 key; const void *ga, *gb; ChainView cv;
            const int kind = m.request(ix, a, o, s, c, key, ga, gb);
            if (kind < 0) break;
            gather(ix, T, kind, key, ga, gb, a, o, s, c, na, no, ns, cv, ctr);
            if (kind == 0) m.consume(ix, na, no, ns); else if (kind == 1) m.consume_chain(ix, cv); else m.consume_aux(ix);
        }
        if (m.ovf) break;
    }
    out = m.out;
    if (m.ovf) return false;
    sort_intv_by_info(out.a, out.n);
    return true;
}

} // namespace b200
