// fmindex.cuh -- FM-index primitives over the 32-byte Occ block layout.
// Semantics follow bwa (bwa/bwt.c:53-275): rank k counts BWT symbols in
// B0[0..k'] with k' = k - (k >= primary) because '$' is not stored.
#pragma once
#include "common.cuh"

namespace b200 {

struct OccLoad { u32 c0, c1, c2, c3; u64 s0, s1; };

HD OccLoad load_block_at(const OccBlock *p)
{
    OccLoad r;
#if defined(__CUDA_ARCH__)
    // one 256-bit load (LDG.E.256 on sm_100a): the block is one 32-byte sector, so one request per rank query
    u32 a0, a1, a2, a3, b0, b1, b2, b3;
    asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3), "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3) : "l"(p));
    r.c0 = a0; r.c1 = a1; r.c2 = a2; r.c3 = a3;
    r.s0 = (u64)b0 | (u64)b1 << 32; r.s1 = (u64)b2 | (u64)b3 << 32;
#else
    const OccBlock &b = *p;
    r.c0 = b.cnt[0]; r.c1 = b.cnt[1]; r.c2 = b.cnt[2]; r.c3 = b.cnt[3];
    r.s0 = b.sym[0]; r.s1 = b.sym[1];
#endif
    return r;
}
HD OccLoad load_block(const DevIndex &ix, u64 blk) { return load_block_at(ix.occ + blk); }

// counts of A,C,G,T in the first r (1..64) symbols of a loaded block, added to the block base.
// s0 / s1 are the low / high bit planes of the 64 symbols: three popcounts under one mask give all four counts.
HD void block_rank4(const OccLoad &b, int r, u64 cnt[4])
{
    u64 m = r >= 64 ? ~0ull : ((1ull << r) - 1);
    u64 lo = b.s0 & m, hi = b.s1 & m;
    int pl = popc64(lo), ph = popc64(hi), pt = popc64(lo & hi);
    cnt[0] = b.c0 + (u64)(r - pl - ph + pt); cnt[1] = b.c1 + (u64)(pl - pt); cnt[2] = b.c2 + (u64)(ph - pt); cnt[3] = b.c3 + (u64)pt;
}

// bwt_2occ4 (bwa/bwt.c:189-220) on ranks k <= l; k or l may be (u64)-1.
template <class Ctr>
HD void occ4_pair(const DevIndex &ix, u64 k, u64 l, u64 ck[4], u64 cl[4], Ctr &ctr)
{
    const u64 NEG1 = ~0ull;
    u64 kk = k, ll = l;
    if (k != NEG1) kk = k - (k >= ix.primary);
    if (l != NEG1) ll = l - (l >= ix.primary);
    if (k == NEG1) { ck[0] = ck[1] = ck[2] = ck[3] = 0; }
    if (l == NEG1) { cl[0] = cl[1] = cl[2] = cl[3] = 0; }
    u64 bk = kk >> 6, bl = ll >> 6;
    if (k != NEG1 && l != NEG1) {
        // both blocks are requested before either is used: two gathers in flight per lane instead of one
        OccLoad b1 = load_block(ix, bk);
        OccLoad b2 = bl != bk ? load_block(ix, bl) : b1;
        ctr.occ_blocks += bl != bk ? 2 : 1;
        block_rank4(b1, (int)(kk & 63) + 1, ck);
        block_rank4(b2, (int)(ll & 63) + 1, cl);
        return;
    }
    if (k != NEG1) { OccLoad b = load_block(ix, bk); ctr.occ_blocks++; block_rank4(b, (int)(kk & 63) + 1, ck); }
    if (l != NEG1) { OccLoad b = load_block(ix, bl); ctr.occ_blocks++; block_rank4(b, (int)(ll & 63) + 1, cl); }
}

// bwt_extend (bwa/bwt.c:262-275) restricted to the one child the callers use: ok[c].
// All four Occ differences are still needed (the reverse-side start of child c skips the children > c);
// everything stays in registers -- no array is indexed by the run-time base c.
template <class Ctr>
HD void extend_c(const DevIndex &ix, const Intv &ik, int c, int is_back, Intv &oc, Ctr &ctr)
{
    u64 tk[4], tl[4];
    u64 a = is_back ? ik.x0 : ik.x1;      // x[!is_back]
    u64 o = is_back ? ik.x1 : ik.x0;      // x[is_back]
    occ4_pair(ix, a - 1, a - 1 + ik.x2, tk, tl, ctr);
    u64 s0 = tl[0] - tk[0], s1 = tl[1] - tk[1], s2 = tl[2] - tk[2], s3 = tl[3] - tk[3];
    u64 no3 = o + (a <= ix.primary && a + ik.x2 - 1 >= ix.primary);
    u64 no2 = no3 + s3, no1 = no2 + s2, no0 = no1 + s1;
    u64 tkc = c == 0 ? tk[0] : c == 1 ? tk[1] : c == 2 ? tk[2] : tk[3];
    u64 sc = c == 0 ? s0 : c == 1 ? s1 : c == 2 ? s2 : s3;
    u64 noc = c == 0 ? no0 : c == 1 ? no1 : c == 2 ? no2 : no3;
    u64 nac = ix.L2[c] + 1 + tkc;
    oc.x2 = sc;
    if (is_back) { oc.x0 = nac; oc.x1 = noc; }
    else { oc.x1 = nac; oc.x0 = noc; }
}

// bwt_set_intv (bwa/bwt.h:82)
HD void set_intv(const DevIndex &ix, int c, Intv &ik)
{
    ik.x0 = ix.L2[c] + 1; ik.x2 = ix.L2[c + 1] - ix.L2[c]; ik.x1 = ix.L2[3 - c] + 1; ik.info = 0;
}

// bwt_B0 (bwa/bwt.h:80): symbol x of the $-less BWT
HD int bwt_sym(const OccLoad &b, int j) { return (int)((b.s0 >> j) & 1) | (int)((b.s1 >> j) & 1) << 1; }

// bwt_invPsi (bwa/bwt.c:53-59) with bwt_occ (bwa/bwt.c:107-129)
template <class Ctr>
HD u64 inv_psi(const DevIndex &ix, u64 k, Ctr &ctr)
{
    if (k == ix.primary) return 0;
    u64 x = k - (k > ix.primary);
    OccLoad b = load_block(ix, x >> 6);
    ctr.occ_blocks++;
    int c = bwt_sym(b, (int)(x & 63));
    // bwt_occ(k, c): k == seq_len cannot reach here with k <= seq_len handled below
    u64 n;
    if (k == ix.seq_len) n = ix.L2[c + 1] - ix.L2[c];
    else {
        u64 kk = k - (k >= ix.primary);     // same block as x (kk == x unless k == primary, excluded above)
        u64 cnt[4];
        if ((kk >> 6) != (x >> 6)) { b = load_block(ix, kk >> 6); ctr.occ_blocks++; }
        block_rank4(b, (int)(kk & 63) + 1, cnt);
        n = cnt[c];
    }
    return ix.L2[c] + n;
}

// bwt_sa (bwa/bwt.c:86-96)
template <class Ctr>
HD u64 sa_lookup(const DevIndex &ix, u64 k, Ctr &ctr)
{
    u64 steps = 0, mask = (1ull << ix.sa_shift) - 1;
    while (k & mask) { ++steps; k = inv_psi(ix, k, ctr); }
    ctr.sa_reads++;
#if defined(__CUDA_ARCH__)
    return steps + __ldg(ix.sa + (k >> ix.sa_shift));
#else
    return steps + ix.sa[k >> ix.sa_shift];
#endif
}

// base at position p (0 <= p < 2*l_pac) of the forward+reverse text: what
// bns_get_seq (bwa/bntseq.c:403-424) unpacks from pac.
HD int text_base(const DevIndex &ix, i64 p)
{
#if defined(__CUDA_ARCH__)
    u64 w = __ldg(ix.text + (p >> 5));
#else
    u64 w = ix.text[p >> 5];
#endif
    return (int)((w >> (2 * (p & 31))) & 3);
}

// bns_pos2rid (bwa/bntseq.c:354-368) on a forward-strand position
HD int pos2rid(const DevIndex &ix, i64 pos_f)
{
    if (pos_f >= ix.l_pac) return -1;
    int lo = 0, hi = ix.n_seqs;      // contig_off[lo] <= pos_f < contig_off[hi]
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (pos_f >= ix.contig_off[mid]) lo = mid; else hi = mid;
    }
    return lo;
}

// bns_depos (bwa/bntseq.h:87-90)
HD i64 depos(const DevIndex &ix, i64 pos, int *is_rev)
{
    *is_rev = pos >= ix.l_pac;
    return *is_rev ? (ix.l_pac << 1) - 1 - pos : pos;
}

// bns_intv2rid (bwa/bntseq.c:370-378)
HD int intv2rid(const DevIndex &ix, i64 rb, i64 re)
{
    int is_rev;
    if (rb < ix.l_pac && re > ix.l_pac) return -2;
    int rid_b = pos2rid(ix, depos(ix, rb, &is_rev));
    int rid_e = rb < re ? pos2rid(ix, depos(ix, re - 1, &is_rev)) : rid_b;
    return rid_b == rid_e ? rid_b : -1;
}

} // namespace b200
