// ksw_wave.cuh -- ksw_extend2 (bwa/ksw.c:416-515) as an anti-diagonal wavefront over packed 16-bit lanes.
//
// G lanes of a warp share one extension.  Every lane owns TWO consecutive target rows, held in the two 16-bit halves
// of its registers (VIADD.16x2 / VIMNMX3.S16x2 / VIADDMNMX.S16x2.RELU on sm_100a): at step s lane g computes cell
// (row 2g, column j) in the low half and cell (row 2g+1, column j-1) in the high half -- two cells of one
// anti-diagonal per instruction, 2G rows in flight per group.  The per-column state of the reference, eh[j] =
// {H(i,j-1), E(i+1,j)}, is not an array: it is a STREAM of 32-bit words (h | valid<<15 | e<<16 | query base<<29) that
// enters lane 0 from shared memory, moves from the low to the high half inside a lane, from lane to lane by one
// warp shuffle per step, and leaves the last lane for shared memory, where the next block of 2G rows picks it up.
// F and the row maximum never leave the lane.
//
// Exactness.  The reference trims the band row by row (beg/end from the zero cells of the previous row,
// bwa/ksw.c:502-505) and keeps stale cells outside the band; row i+1 here runs one column behind row i, so it
// discovers its band while streaming:
//   * left edge: cells left of the first non-zero column of the previous row evaluate to zero, computing them is
//     the same as skipping them (the block starts at the first non-zero column of the last finished row);
//   * right edge: end(i+1) = last non-zero column of row i, plus 2.  A cell is certainly inside when the incoming
//     word of its column or of the previous column is non-zero; beyond the previous row's extent the decision is
//     exact too (valid bit).  After two zero columns a row whose F is zero keeps computing (such cells are zero, i.e.
//     what the reference leaves there); a row whose F is still positive stops, which is what the reference does
//     unless row i turns non-zero again further right -- that case ("gap event") is detected and reported, and the
//     caller re-runs the read with the row-synchronous kernel (ksw_reg.cuh).  Nothing is approximated.
//   * per-row results (max, its last column, H at the query end) are committed in row order after each block, so
//     z-drop / zero-row breaks discard the rows in flight behind them exactly like the reference's loop exit.
// Restrictions (caller): no base > 3 in query or target, match/mismatch score matrix (bwa_fill_scmat), scores < 2^13,
// qlen <= WAVE_MAXQ.  wave_lane_* are plain HD code: tests/hostsim/wave_emul.cpp runs them lane by lane on the CPU.
#pragma once
#include "common.cuh"
#include "ksw.cuh"

namespace b200 {

#define WAVE_MAXQ 1023        // columns: stream words of one extension live in shared memory
#define WAVE_PAD(G) (2 * (G) + 2)                 // guard words in front of column 0 (and, plus 8, behind column qlen)
#define WAVE_WORDS(maxq, G) ((maxq) + 2 + 2 * WAVE_PAD(G) + 8)

namespace wv {

HD u32 perm(u32 a, u32 b, u32 sel)        // prmt.b32, default mode (bit 3 of a selector nibble: replicate the byte's sign)
{
#if defined(__CUDA_ARCH__)
    u32 d; asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel)); return d;
#else
    const u64 v = (u64)b << 32 | a; u32 d = 0;
    for (int k = 0; k < 4; ++k) {
        const u32 n = (sel >> (4 * k)) & 15u;
        u32 byte = (u32)(v >> (8 * (n & 7u))) & 0xffu;
        if (n & 8u) byte = (byte & 0x80u) ? 0xffu : 0u;
        d |= byte << (8 * k);
    }
    return d;
#endif
}
HD u32 signmask(u32 x) { return perm(x, 0u, 0xbb99u); }     // 0xffff in every half whose bit 15 is set
HD u32 add2(u32 a, u32 b)
{
#if defined(__CUDA_ARCH__)
    return __vadd2(a, b);
#else
    return ((a & 0xffffu) + (b & 0xffffu)) & 0xffffu | ((a >> 16) + (b >> 16)) << 16;
#endif
}
#if !defined(__CUDA_ARCH__)
inline int lo16(u32 a) { return (int)(i16)(a & 0xffffu); }
inline int hi16(u32 a) { return (int)(i16)(a >> 16); }
inline u32 pack16(int lo, int hi) { return (u32)(u16)(i16)lo | (u32)(u16)(i16)hi << 16; }
inline int mx(int a, int b) { return a > b ? a : b; }
#endif
HD u32 max3s(u32 a, u32 b, u32 c)
{
#if defined(__CUDA_ARCH__)
    return __vimax3_s16x2(a, b, c);
#else
    return pack16(mx(mx(lo16(a), lo16(b)), lo16(c)), mx(mx(hi16(a), hi16(b)), hi16(c)));
#endif
}
HD u32 addmax_relu(u32 a, u32 b, u32 c)      // max(a + b, c, 0)
{
#if defined(__CUDA_ARCH__)
    return __viaddmax_s16x2_relu(a, b, c);
#else
    return pack16(mx(mx((int)(i16)(lo16(a) + lo16(b)), lo16(c)), 0), mx(mx((int)(i16)(hi16(a) + hi16(b)), hi16(c)), 0));
#endif
}
// max(a, b) per half; pl / ph = (a >= b) in the low / high half
HD u32 bmax(u32 a, u32 b, bool &ph, bool &pl)
{
#if defined(__CUDA_ARCH__)
    return __vibmax_s16x2(a, b, &ph, &pl);
#else
    pl = lo16(a) >= lo16(b); ph = hi16(a) >= hi16(b);
    return pack16(pl ? lo16(a) : lo16(b), ph ? hi16(a) : hi16(b));
#endif
}
HD u32 min2s(u32 a, u32 b)
{
#if defined(__CUDA_ARCH__)
    return __vmins2(a, b);
#else
    return pack16(lo16(a) < lo16(b) ? lo16(a) : lo16(b), hi16(a) < hi16(b) ? hi16(a) : hi16(b));
#endif
}
HD u32 sel(u32 m, u32 a, u32 b) { return (a & m) | (b & ~m); }     // one LOP3

} // namespace wv

struct WaveConst {           // per extension, the same in all lanes
    u32 A2, CSC;             // a in both halves; 65536 - (a + b): score = A2 + mismatch * CSC per half, one 32-bit IMAD (no carry crosses
                             // the halves because b >= 1)
    u32 NOED, NED, NOEI, NEI;    // -(o_del + e_del), -e_del, -(o_ins + e_ins), -e_ins in both halves
};

HD u32 wave_rep(int v) { return (u32)(u16)(i16)v * 0x00010001u; }

HD WaveConst wave_const(int a, int b, int o_del, int e_del, int o_ins, int e_ins)
{
    WaveConst K;
    K.A2 = wave_rep(a); K.CSC = (u32)(65536 - (a + b));
    K.NOED = wave_rep(-(o_del + e_del)); K.NED = wave_rep(-e_del);
    K.NOEI = wave_rep(-(o_ins + e_ins)); K.NEI = wave_rep(-e_ins);
    return K;
}

// stream word: h (bits 0-13) | valid (bit 15) | e (bits 16-28) | query base (bits 29-31)
HD u32 wave_word(int h, int e, int q, bool valid) { return (u32)h | (valid ? 0x8000u : 0u) | (u32)e << 16 | (u32)q << 29; }

// One lane = two rows (low half: row 2g, high half: row 2g + 1) of the current block.
struct WaveLane {
    u32 H1, F, MX, T;                    // H(i, j-1), F(i, j), row maximum, target base -- per half
    u32 CL, CR, J;                       // j - lo, hi - j, j -- per half (16-bit counters)
    u32 DONE, NZP, WATCH, FORCE, GH;     // per-half masks
    u32 XC;                              // last valid column written
    u32 plo;                             // the low row's output word of the previous step = the high row's input
    int mj_lo, mj_hi, jl;
    u32 gap;

    // rows rl (low) and rl + 1 (high); columns start at cb; h1l / h1h = H(row, -1) where the row starts at column 0
    HD void setup(int rl, int tlen, int tl, int th, int cb, int gl, int w, int qlen, int h1l, int h1h)
    {
        const int rh = rl + 1;
        const int lol = cb > rl - w ? cb : rl - w, loh = cb > rh - w ? cb : rh - w;
        const int hil = qlen - 1 < rl + w ? qlen - 1 : rl + w, hih = qlen - 1 < rh + w ? qlen - 1 : rh + w;
        jl = cb - 2 * gl;                                 // the low row's column at step 0; the high row is one behind
        J = (u32)(u16)jl | (u32)(u16)(jl - 1) << 16;
        CL = (u32)(u16)(jl - lol) | (u32)(u16)(jl - 1 - loh) << 16;
        CR = (u32)(u16)(hil - jl) | (u32)(u16)(hih - (jl - 1)) << 16;
        T = (u32)tl | (u32)th << 16;
        H1 = (u32)(lol == 0 ? h1l : 0) | (u32)(loh == 0 ? h1h : 0) << 16;
        F = 0; MX = 0; NZP = WATCH = GH = 0; XC = 0; plo = 0; gap = 0;
        mj_lo = mj_hi = -1;
        DONE = (rl >= tlen || lol > qlen ? 0x0000ffffu : 0u) | (rh >= tlen || loh > qlen ? 0xffff0000u : 0u);
        FORCE = rl == 0 ? 0x0000ffffu : 0u;               // row 0 covers [0, min(qlen, w + 1)) whatever the initial row holds
    }

    // One anti-diagonal step.  win = the stream word of column jl coming from the row above the low row;
    // returns the high row's output word (column jl - 1).
    HD u32 step(const WaveConst &K, u32 win)
    {
        using namespace wv;
        const u32 A = perm(win, plo, 0x5410u), B = perm(win, plo, 0x7632u);
        const u32 HIN = A & 0x3fff3fffu, EIN = B & 0x1fff1fffu, Q = (B >> 13) & 0x00070007u;
        const u32 INV = signmask(A);
        const u32 BEF = signmask(CL), AFT = signmask(CR);
        const u32 NZ = ~signmask(add2(HIN | EIN, 0xffffffffu)) & INV;
        const u32 NFZ = ~signmask(add2(F, 0xffffffffu));
        const u32 CAN = ~(DONE | BEF | AFT);
        const u32 SURE = FORCE | NZ | NZP;
        const u32 LEGIT = CAN & (SURE | (INV & ~NFZ));
        const u32 TERM = ~(DONE | BEF | LEGIT);
        const u32 TERMW = TERM & ~(GH & ~NZ);              // a terminal reached while computing zeros writes nothing
        gap |= DONE & WATCH & NZ;
        WATCH |= TERM & ~AFT & INV & NFZ;
        GH = LEGIT & ~SURE;
        // the cell (bwa/ksw.c:455-482)
        // score = a or -b per half; M = H(i-1,j-1) + score where H(i-1,j-1) != 0, else min(score, 0) -- a non-positive M is
        // as good as the reference's 0 everywhere it is used (max with E, F >= 0; the relu of the E / F updates)
        const u32 SC = min2s(Q ^ T, 0x00010001u) * K.CSC + K.A2;
        const u32 M = min2s(add2(HIN, SC), min2s(HIN, 0x00010001u) * 0x7fffu);
        const u32 H = max3s(M, EIN, F);
        const u32 En = addmax_relu(M, K.NOED, add2(EIN, K.NED));
        const u32 Fn = addmax_relu(M, K.NOEI, add2(F, K.NEI));
        bool ph, pl;
        MX = bmax(H & LEGIT, MX, ph, pl);
        mj_lo = pl ? jl : mj_lo;
        mj_hi = ph ? jl - 1 : mj_hi;
        // outputs: computed cell -> {h1, E'}; terminal -> {h1, 0}; otherwise the incoming word (valid bit dropped once done)
        const u32 WR = LEGIT | TERMW;
        // (the valid bit of a word is exactly "this row wrote it": columns left of a row's start are never examined by the rows below)
        const u32 OA = (sel(WR, H1, A) & 0x7fff7fffu) | (WR & 0x80008000u);
        const u32 OB = sel(LEGIT, En, EIN & ~TERMW) | (B & 0xe000e000u);
        XC = sel(WR, J, XC);
        H1 = sel(LEGIT, H, H1);
        F = sel(LEGIT, Fn, F);
        DONE |= TERM;
        NZP = NZ;
        CL = add2(CL, 0x00010001u); CR = add2(CR, 0xffffffffu); J = add2(J, 0x00010001u); ++jl;
        plo = perm(OA, OB, 0x5410u);
        return perm(OA, OB, 0x7632u);
    }
};

// running state of ksw_extend2's row loop (bwa/ksw.c:446-500)
struct WaveAcc {
    int max, max_i, max_j, max_ie, gscore, max_off;
    bool broke;
    HD void init(int h0) { max = h0; max_i = max_j = max_ie = -1; gscore = -1; max_off = 0; broke = false; }
    // row i finished with maximum m at column mj; x = its last valid column (== qlen: the row reached the query end), h1 = H(i, x-1)
    // lo = the first column of the row (max of the block's start column and i - w)
    HD void commit(int i, int lo, int m, int mj, int x, int h1, int qlen, int zdrop, int e_del, int e_ins)
    {
        if (lo >= qlen) {        // band entirely right of the query: the reference's column loop does not run (bwa/ksw.c:455), j == beg
            if (lo == qlen) { max_ie = gscore > 0 ? max_ie : i; gscore = gscore > 0 ? gscore : 0; }
            broke = true;
            return;
        }
        if (x == qlen) { max_ie = gscore > h1 ? max_ie : i; gscore = gscore > h1 ? gscore : h1; }
        if (m == 0) { broke = true; return; }
        if (m > max) {
            max = m; max_i = i; max_j = mj;
            int k = mj - i; k = k < 0 ? -k : k;
            max_off = max_off > k ? max_off : k;
        } else if (zdrop > 0) {
            if (i - max_i > mj - max_j) { if (max - m - ((i - max_i) - (mj - max_j)) * e_del > zdrop) broke = true; }
            else { if (max - m - ((mj - max_j) - (i - max_i)) * e_ins > zdrop) broke = true; }
        }
    }
};

// the band adjustment of bwa/ksw.c:437-442.  (int)((double)x / e + 1.) == x / e + 1 in integers wherever the result exceeds 1
// (they differ only for -1 < x / e < 0, where both are clamped to 1), so no double-precision division is needed.
HD int wave_band(int qlen, int maxsc, int end_bonus, int o_del, int e_del, int o_ins, int e_ins, int w)
{
    int max_ins = (qlen * maxsc + end_bonus - o_ins) / e_ins + 1;
    max_ins = max_ins > 1 ? max_ins : 1;
    w = w < max_ins ? w : max_ins;
    int max_del = (qlen * maxsc + end_bonus - o_del) / e_del + 1;
    max_del = max_del > 1 ? max_del : 1;
    return w < max_del ? w : max_del;
}

HD int wave_h1_init(int h0, int o_del, int e_del, int row) { int v = h0 - (o_del + e_del * (row + 1)); return v > 0 ? v : 0; }

// can this extension run on the wavefront kernel?  (a, b: match / mismatch of the matrix)
HD bool wave_eligible(int qlen, int tlen, int h0, int a, int end_bonus)
{
    return qlen >= 1 && qlen <= WAVE_MAXQ && tlen < 30000 && a > 0 && h0 + qlen * a + end_bonus < 8000;      // (b >= 1: checked once per batch)
}

#if defined(__CUDACC__)
// G lanes (GroupCtx<G>) run one extension; every lane returns the same result.  ehs: qlen + 2 words of shared memory owned
// by the group.  Returns false on a gap event (result invalid, re-run with extend2_reg).
template <int G, class GCtx, class QSeq, class TSeq, class Ctr>
__device__ __noinline__ bool extend2_wave(const GCtx &g, int qlen, const QSeq &query, int tlen, const TSeq &target, int a, int b,
                             int o_del, int e_del, int o_ins, int e_ins, int w, int end_bonus, int zdrop, int h0,
                             u32 *ehs, ExtResult &R, Ctr &ctr)
{
    const int gl = g.gl;
    const WaveConst K = wave_const(a, b, o_del, e_del, o_ins, e_ins);
    w = wave_band(qlen, a, end_bonus, o_del, e_del, o_ins, e_ins, w);
    u32 gapped = 0;
    u32 *const E = ehs + WAVE_PAD(G);             // column j at E[j]; guard words on both sides (pipeline fill / drain touches them)
    for (int j = gl - WAVE_PAD(G); j <= qlen + WAVE_PAD(G) + 8; j += G) {        // eh[] after the reference's initialisation (bwa/ksw.c:428-432)
        int v = h0 - (o_ins + e_ins) - (j - 1) * e_ins;
        v = j == 0 ? h0 : (v > 0 ? v : 0);
        const int q = j >= 0 && j < qlen ? (int)query[j] : 0;
        gapped |= q > 3 ? 1u : 0u;                // ambiguous bases score -1 against everything: not a match / mismatch matrix
        E[j] = j >= 0 && j <= qlen ? wave_word(v, 0, q & 3, true) : 0u;
    }
    WaveAcc acc; acc.init(h0);
    int cb = 0, xprev = qlen;
    unsigned long long cells = 0;
    g.sync();
    // (Tried: warp-wide votes that keep the groups of a warp in the same block / step loops, a finished group stepping idle.  The
    // set-up and commit code is then issued once per warp instead of once per group, but every group waits for the slowest:
    // 74.3 vs 70.7 ms per 4 M reads.  The groups run their blocks independently.)
    for (int r0 = 0; r0 < tlen && !acc.broke; r0 += 2 * G) {
        const int rl = r0 + 2 * gl;
        WaveLane L;
        const int tb0 = rl < tlen ? (int)target[rl] : 0, tb1 = rl + 1 < tlen ? (int)target[rl + 1] : 0;
        gapped |= (tb0 | tb1) > 3 ? 1u : 0u;
        L.setup(rl, tlen, tb0 & 3, tb1 & 3, cb, gl, w, qlen,
                cb == 0 ? wave_h1_init(h0, o_del, e_del, rl) : 0, cb == 0 ? wave_h1_init(h0, o_del, e_del, rl + 1) : 0);
        u32 oh = 0;
        const bool first = gl == 0, last = gl == G - 1;
        u32 *pj = E + (cb - 2 * gl);                     // the word of this lane's low-row column
        for (int s = 0, smax = qlen + 2 - cb + 2 * G; s <= smax; s += 4) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                u32 win = (u32)g.up((int)oh, 1);
                const u32 w0 = *pj;                      // lane 0 takes its word from shared memory (valid bits are kept right there)
                win = first ? w0 : win;
                oh = L.step(K, win);
                if (last) pj[-1] = oh;
                ++pj;
            }
            if (__all_sync(g.mask, L.DONE == 0xffffffffu)) break;
        }
        g.sync();
        xprev = (int)(((u32)g.bcast((int)L.XC, G - 1) >> 16) & 0xffffu);
        // words beyond the last row's extent that this block did not stream keep an old valid bit: drop it
        for (int j = xprev + 1 + gl; j <= qlen; j += G) E[j] &= ~0x8000u;
        g.sync();
        // the next block starts at the first non-zero column of the stream the last row left behind
        int cbn = 1 << 20;
        for (int j = cb + gl; j <= qlen; j += G) if ((E[j] & 0x1fff3fffu) != 0) { cbn = j; break; }
        cbn = g.rmin(cbn);
        // commit the block's rows in order
        const u32 P2 = (u32)(u16)L.mj_lo | (u32)(u16)L.mj_hi << 16;
        gapped |= L.gap;
#pragma unroll
        for (int k = 0; k < 2 * G; ++k) {
            const int src = k >> 1, sh = (k & 1) * 16;
            const int m = (int)(((u32)g.bcast((int)L.MX, src) >> sh) & 0xffffu);
            const int mj = (int)(i16)(((u32)g.bcast((int)P2, src) >> sh) & 0xffffu);
            const int x = (int)(((u32)g.bcast((int)L.XC, src) >> sh) & 0xffffu);
            const int h1 = (int)(((u32)g.bcast((int)L.H1, src) >> sh) & 0xffffu);
            const int row = r0 + k;
            if (row < tlen && !acc.broke) {
                const int lo = cb > row - w ? cb : row - w;
                cells += x > lo ? (unsigned long long)(x - lo) : 0ull;
                acc.commit(row, lo, m, mj, x, h1, qlen, zdrop, e_del, e_ins);
            }
        }
        if (cbn < (1 << 20) && cbn > cb) cb = cbn;
        g.sync();
    }
    gapped = (u32)g.rmax((int)(gapped != 0));
    if (gl == 0) { ctr.sw_cells += cells; ctr.n_ext++; }
    R.score = acc.max; R.qle = acc.max_j + 1; R.tle = acc.max_i + 1; R.gtle = acc.max_ie + 1; R.gscore = acc.gscore; R.max_off = acc.max_off;
    return gapped == 0;
}
#endif

} // namespace b200
