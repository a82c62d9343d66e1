// fmd.cuh -- the FMD-index of a read set (fermi-lite's rld_t, fermi-lite/rld0.h:25-47) as fixed 64-byte rank blocks.
//
//   fmd_rank1a / fmd_rank2a / fmd_extend / fmd_extend0   <- rld_rank1a, rld_rank2a, rld_extend (fermi-lite/rld0.c:402-489),
//                                                           rld_extend0 (fermi-lite/unitig.c:21-30)
//
// The reference stores the BWT of {read, revcomp(read)} as delta-coded runs and ranks by decoding a block's bit stream.
// Only the BWT string itself is observable through rank queries, so the device keeps it as three bit planes per 128
// symbols plus the running A/C/G/T counts: one rank = one aligned 64-byte load + popcounts.  The alphabet is the
// reference's ($=0, A=1, C=2, G=3, T=4, N=5); reads containing N never enter the index (fermi-lite/misc.c:91-96), so
// symbol 5 has no occurrences.
#pragma once
#include "common.cuh"

namespace b200 {

struct FmdBlock {            // 128 BWT symbols
    u32 cnt[4];              // A, C, G, T in bwt[0, 128 * block)  ($ = position - sum)
    u64 p0[2], p1[2], p2[2]; // symbol j of half h: p2 bit = ($), else code - 1 = p1 bit << 1 | p0 bit
};

struct FmdIndex {
    const FmdBlock *blk;
    u64 n;                   // BWT length = mcnt[0]
    u64 cnt[7];              // cnt[c] = number of symbols < c  (rld_t.cnt after rld_enc_finish)
    u64 n_str;               // number of sentinels = mcnt[1]
};

struct FmdIntv { u64 x[3]; u64 info; };      // rldintv_t (fermi-lite/rld0.h:43-46)

// seq_nt6_table (fermi-lite/misc.c:12-29): A C G T -> 1..4, anything else 5
HD int fmd_nt6(char ch)
{
    switch (ch) {
    case 'A': case 'a': return 1;
    case 'C': case 'c': return 2;
    case 'G': case 'g': return 3;
    case 'T': case 't': return 4;
    default: return 5;
    }
}

HD int fmd_comp(int a) { return a >= 1 && a <= 4 ? 5 - a : a; }

HD void fmd_set_intv(const FmdIndex &e, int c, FmdIntv &ik)      // fm6_set_intv (fermi-lite/unitig.c:19)
{
    ik.x[0] = e.cnt[c]; ik.x[2] = e.cnt[c + 1] - e.cnt[c]; ik.x[1] = e.cnt[fmd_comp(c)]; ik.info = 0;
}

HD void fmd_load_block(const FmdIndex &e, u64 b, FmdBlock &B)
{
#if defined(__CUDA_ARCH__)
    const ulonglong2 *p = reinterpret_cast<const ulonglong2 *>(e.blk + b);
    ulonglong2 v0 = __ldg(p), v1 = __ldg(p + 1), v2 = __ldg(p + 2), v3 = __ldg(p + 3);
    B.cnt[0] = (u32)v0.x; B.cnt[1] = (u32)(v0.x >> 32); B.cnt[2] = (u32)v0.y; B.cnt[3] = (u32)(v0.y >> 32);
    B.p0[0] = v1.x; B.p0[1] = v1.y; B.p1[0] = v2.x; B.p1[1] = v2.y; B.p2[0] = v3.x; B.p2[1] = v3.y;
#else
    B = e.blk[b];
#endif
}

// counts of $ A C G T in the first o (1..128) symbols of block b added to the block's running counts
HD void fmd_block_rank(const FmdBlock &B, u64 b, int o, u64 ok[6])
{
    u64 m0 = o >= 64 ? ~0ull : (1ull << o) - 1;
    u64 m1 = o <= 64 ? 0ull : (o >= 128 ? ~0ull : (1ull << (o - 64)) - 1);
    u64 a0 = ~B.p2[0] & m0, a1 = ~B.p2[1] & m1;          // non-sentinel positions
    u64 cA = popc64(a0 & ~B.p1[0] & ~B.p0[0]) + popc64(a1 & ~B.p1[1] & ~B.p0[1]);
    u64 cC = popc64(a0 & ~B.p1[0] & B.p0[0]) + popc64(a1 & ~B.p1[1] & B.p0[1]);
    u64 cG = popc64(a0 & B.p1[0] & ~B.p0[0]) + popc64(a1 & B.p1[1] & ~B.p0[1]);
    u64 cT = popc64(a0 & B.p1[0] & B.p0[0]) + popc64(a1 & B.p1[1] & B.p0[1]);
    u64 before = (u64)B.cnt[0] + B.cnt[1] + B.cnt[2] + B.cnt[3];
    ok[1] = B.cnt[0] + cA; ok[2] = B.cnt[1] + cC; ok[3] = B.cnt[2] + cG; ok[4] = B.cnt[3] + cT;
    ok[0] = (b << 7) - before + (u64)(popc64(B.p2[0] & m0) + popc64(B.p2[1] & m1));
    ok[5] = 0;
}

HD int fmd_block_sym(const FmdBlock &B, int j)           // symbol j (0..127) of the block
{
    int h = j >> 6, s = j & 63;
    if (B.p2[h] >> s & 1) return 0;
    return 1 + (int)(B.p0[h] >> s & 1) + ((int)(B.p1[h] >> s & 1) << 1);
}

// rld_rank1a: ok[c] = occurrences of c in bwt[0, k); returns bwt[k - 1] (-1 when k == 0)
HD int fmd_rank1a(const FmdIndex &e, u64 k, u64 ok[6])
{
    if (k == 0) { for (int a = 0; a < 6; ++a) ok[a] = 0; return -1; }
    u64 b = (k - 1) >> 7; int o = (int)((k - 1) & 127) + 1;
    FmdBlock B;
    fmd_load_block(e, b, B);
    fmd_block_rank(B, b, o, ok);
    return fmd_block_sym(B, o - 1);
}

// rld_rank2a: both ends of an interval (k <= l); one block load when they share a block
HD void fmd_rank2a(const FmdIndex &e, u64 k, u64 l, u64 ok[6], u64 ol[6])
{
    if (k == 0) { for (int a = 0; a < 6; ++a) ok[a] = 0; fmd_rank1a(e, l, ol); return; }
    u64 bk = (k - 1) >> 7;
    FmdBlock B;
    fmd_load_block(e, bk, B);
    fmd_block_rank(B, bk, (int)((k - 1) & 127) + 1, ok);
    if (l == 0) { for (int a = 0; a < 6; ++a) ol[a] = 0; return; }
    u64 bl = (l - 1) >> 7;
    if (bl != bk) fmd_load_block(e, bl, B);
    fmd_block_rank(B, bl, (int)((l - 1) & 127) + 1, ol);
}

// rld_extend (fermi-lite/rld0.c:472-489)
HD void fmd_extend(const FmdIndex &e, const FmdIntv &ik, FmdIntv ok[6], int is_back)
{
    u64 tk[6], tl[6];
    fmd_rank2a(e, ik.x[!is_back], ik.x[!is_back] + ik.x[2], tk, tl);
    for (int i = 0; i < 6; ++i) {
        ok[i].x[!is_back] = e.cnt[i] + tk[i];
        ok[i].x[2] = (tl[i] -= tk[i]);
    }
    ok[0].x[is_back] = ik.x[is_back];
    ok[4].x[is_back] = ok[0].x[is_back] + tl[0];
    ok[3].x[is_back] = ok[4].x[is_back] + tl[4];
    ok[2].x[is_back] = ok[3].x[is_back] + tl[3];
    ok[1].x[is_back] = ok[2].x[is_back] + tl[2];
    ok[5].x[is_back] = ok[1].x[is_back] + tl[1];
}

// number of sentinels in bwt[0, k): the only count rld_extend0 consumes (one block load, two popcounts)
HD u64 fmd_rank_sentinel(const FmdIndex &e, u64 k)
{
    if (k == 0) return 0;
    u64 b = (k - 1) >> 7; int o = (int)((k - 1) & 127) + 1;
#if defined(__CUDA_ARCH__)
    const ulonglong2 *p = reinterpret_cast<const ulonglong2 *>(e.blk + b);
    ulonglong2 v0 = __ldg(p), v3 = __ldg(p + 3);
    u64 before = (u64)(u32)v0.x + (u32)(v0.x >> 32) + (u32)v0.y + (u32)(v0.y >> 32);
    u64 q0 = v3.x, q1 = v3.y;
#else
    const FmdBlock &B = e.blk[b];
    u64 before = (u64)B.cnt[0] + B.cnt[1] + B.cnt[2] + B.cnt[3];
    u64 q0 = B.p2[0], q1 = B.p2[1];
#endif
    u64 m0 = o >= 64 ? ~0ull : (1ull << o) - 1;
    u64 m1 = o <= 64 ? 0ull : (o >= 128 ? ~0ull : (1ull << (o - 64)) - 1);
    return (b << 7) - before + (u64)(popc64(q0 & m0) + popc64(q1 & m1));
}

// rld_extend0 (fermi-lite/unitig.c:21-30): only the sentinel branch of an extension
HD void fmd_extend0(const FmdIndex &e, const FmdIntv &ik, FmdIntv &ok0, int is_back)
{
    u64 tk = fmd_rank_sentinel(e, ik.x[!is_back]), tl = fmd_rank_sentinel(e, ik.x[!is_back] + ik.x[2]);
    ok0.x[!is_back] = tk;
    ok0.x[is_back] = ik.x[is_back];
    ok0.x[2] = tl - tk;
}

} // namespace b200
