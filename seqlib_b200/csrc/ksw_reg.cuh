// ksw_reg.cuh -- ksw_extend2 (bwa/ksw.c:416-515) with the DP state in registers: G lanes share one extension and
// each lane OWNS a fixed run of C = ceil((qlen+1)/G) columns for the whole call.
//
// ksw_group.cuh re-cuts the current band into G chunks every row and keeps H/E in shared memory: two barriers, a few
// dozen shared-memory round trips and a score-matrix lookup per cell on the critical path of every row.  For the
// reads this engine is built for (<= 151 columns) the band nearly always spans the whole query, so nothing is
// lost by fixing the ownership -- and then eh[] (bwa/ksw.c:412-414) is a pair of register arrays, the per-column
// scores of the four target bases are one packed register per column, and a row needs shuffles only:
//     phase A   fold the lane's active cells into the max-plus summary (n, L) of F          (reg_phase_a)
//     scan      inclusive max-plus scan over the lanes -> carry-in F of every lane
//     phase B   the true recurrence over the lane's cells, stale cells outside the band kept  (reg_phase_b)
//     phase C   H of the lane's last cell moves into the first column of the next lane        (reg_phase_c)
//     reduce    row maximum with its column, first/last non-zero cell (band trimming), H of the last cell
// Row-exact band trimming, z-drop and the stale-cell semantics of the reference are preserved: a
// column keeps its old (h, e) whenever the row does not touch it.  int32 like the reference.
// The three phase functions are plain HD code: tests/hostsim/reg_emul.cpp runs them lane by lane in lock step on
// the CPU and proves them equal to the scalar recurrence (ksw.cuh extend2) on millions of random tuples.
#pragma once
#include "common.cuh"
#include "ksw.cuh"

namespace b200 {

struct RegConst { int oe_del, e_del, oe_ins, e_ins; };

template <int CMAX>
struct RegLane {
    int hd[CMAX];      // eh[j].h of the lane's columns: H(i-1, j-1)
    int e[CMAX];       // eh[j].e
    u32 sc[CMAX];      // scores of column j against target base 0..3, one signed byte each
    int j0, C, qlen;
};

HD int reg_score(u32 packed, int t) { return (int)(i8)(packed >> (8 * t)); }

// columns [j0, j0 + C) of eh[] after the reference's initialisation (bwa/ksw.c:428-432)
template <int CMAX, class QSeq>
HD void reg_init(RegLane<CMAX> &S, int gl, int C, int qlen, const QSeq &query, const i8 *mat, int h0, int oe_ins, int e_ins)
{
    S.j0 = gl * C; S.C = C; S.qlen = qlen;
#pragma unroll
    for (int jj = 0; jj < CMAX; ++jj) {
        if (jj >= C) break;
        int j = S.j0 + jj;
        int v = h0 - oe_ins - (j - 1) * e_ins;
        S.hd[jj] = j == 0 ? h0 : (v > 0 ? v : 0);
        S.e[jj] = 0;
        u32 p = 0;
        if (j < qlen) {
            int q = query[j];
            p = (u32)(u8)mat[q] | (u32)(u8)mat[5 + q] << 8 | (u32)(u8)mat[10 + q] << 16 | (u32)(u8)mat[15 + q] << 24;
        }
        S.sc[jj] = p;
    }
}

HD int iaddmax(int a, int b, int c)      // max(a + b, c): one VIADDMNMX on sm_100a
{
#if defined(__CUDA_ARCH__)
    return __viaddmax_s32(a, b, c);
#else
    int v = a + b; return v > c ? v : c;
#endif
}
HD int imax3(int a, int b, int c)
{
#if defined(__CUDA_ARCH__)
    return __vimax3_s32(a, b, c);
#else
    int v = a > b ? a : b; return v > c ? v : c;
#endif
}

// Both phase functions are straight-line per column (selects, no data-dependent branches): the 32 lanes of a warp
// work on four different extensions with different bands, and any branch on band membership would split them.
// They require end > beg.

// summary of F over the lane's active cells: out(x) = max(x - n * e_ins, L)
template <int CMAX>
HD void reg_phase_a(const RegLane<CMAX> &S, int beg, int end, int t, const RegConst &K, int &n, int &L)
{
    const int NEG = -(1 << 29);
    const unsigned W = (unsigned)(end - beg);
    const int sh = 24 - 8 * t;
    n = 0; L = NEG;
#pragma unroll
    for (int jj = 0; jj < CMAX; ++jj) {
        if (jj >= S.C) break;
        const bool act = (unsigned)(S.j0 + jj - beg) < W;
        const int hp = S.hd[jj];
        const int sc = (int)(S.sc[jj] << sh) >> 24;
        const int M = hp ? hp + sc : 0;
        const int v = iaddmax(M, -K.oe_ins, 0);
        const int L2 = iaddmax(L, -K.e_ins, v);
        L = act ? L2 : L;
        n += act ? 1 : 0;
    }
}

struct RegRowOut {
    int key;            // (row maximum over the lane's cells) << 10 | (its last column + 1); -1: no cell
    int hlast;          // h of the lane's last owned column if that cell was computed, else -1
    int hend;           // h of column end-1 if the lane owns it
};

// the recurrence of bwa/ksw.c:455-482 over the lane's cells; f = carry-in of F
template <int CMAX>
HD void reg_phase_b(RegLane<CMAX> &S, int beg, int end, int t, int f, int h1_init, const RegConst &K, RegRowOut &o)
{
    const unsigned W = (unsigned)(end - beg);
    const int sh = 24 - 8 * t;
    o.key = -1; o.hlast = -1; o.hend = 0;
    int carry = 0; bool prev_act = false;
#pragma unroll
    for (int jj = 0; jj < CMAX; ++jj) {
        if (jj >= S.C) break;
        const int j = S.j0 + jj;
        const bool act = (unsigned)(j - beg) < W;
        const int old = S.hd[jj], e = S.e[jj];
        // eh[j].h after the row: h1 of the first cell, H(i, j-1) where the cell to the left was computed (this also
        // stores column `end`), untouched otherwise.  Column 0 of the lane gets its left neighbour in phase C.
        S.hd[jj] = j == beg ? h1_init : (prev_act ? carry : old);
        const int sc = (int)(S.sc[jj] << sh) >> 24;
        const int M = old ? old + sc : 0;
        const int h = imax3(M, e, f);
        const int e2 = iaddmax(e, -K.e_del, iaddmax(M, -K.oe_del, 0));
        const int f2 = iaddmax(f, -K.e_ins, iaddmax(M, -K.oe_ins, 0));
        const int k2 = h << 10 | (j + 1);
        o.key = act && k2 > o.key ? k2 : o.key;
        S.e[jj] = act ? e2 : (j == end ? 0 : e);          // eh[end].e = 0
        f = act ? f2 : f;
        carry = h;
        o.hend = j == end - 1 ? h : o.hend;
        prev_act = act;
    }
    if (prev_act) o.hlast = carry;
}

// hin = hlast of the previous lane (-1: its last cell was not computed)
template <int CMAX>
HD void reg_phase_c(RegLane<CMAX> &S, int gl, int beg, int hin)
{
    if (gl > 0 && hin >= 0 && S.j0 != beg) S.hd[0] = hin;
}

// bit jj set: column j0 + jj lies in [beg, end] and holds a non-zero h or e (what the band trimming of bwa/ksw.c:502-505 looks at)
template <int CMAX>
HD unsigned reg_nonzero_mask(const RegLane<CMAX> &S, int beg, int end)
{
    const unsigned W1 = (unsigned)(end - beg) + 1u;
    unsigned m = 0;
#pragma unroll
    for (int jj = 0; jj < CMAX; ++jj) {
        if (jj >= S.C) break;
        const bool in = (unsigned)(S.j0 + jj - beg) < W1;
        m |= (in && (S.hd[jj] | S.e[jj]) != 0) ? 1u << jj : 0u;
    }
    return m;
}

// first non-zero column in [beg, end) and last non-zero column in [beg, end] of the lane (1<<30 / -1: none)
HD void reg_mask_range(unsigned m, int j0, int end, int &mn, int &mx)
{
    mx = -1; mn = 1 << 30;
    if (m) {
#if defined(__CUDA_ARCH__)
        mx = j0 + 31 - __clz(m);
        int lo = j0 + __ffs(m) - 1;
#else
        mx = j0 + 31 - __builtin_clz(m);
        int lo = j0 + __builtin_ctz(m);
#endif
        if (lo < end) mn = lo;
    }
}

#if defined(__CUDACC__)
// G lanes (a GroupCtx<G> of ksw_group.cuh) run one extension; every lane returns the same result.
template <int G, int CMAX, class GCtx, class QSeq, class TSeq, class Ctr>
__device__ ExtResult extend2_reg(const GCtx &g, int qlen, const QSeq &query, int tlen, const TSeq &target, const i8 *smat,
                                 int o_del, int e_del, int o_ins, int e_ins, int w, int end_bonus, int zdrop, int h0, Ctr &ctr)
{
    const int gl = g.gl;
    RegConst K; K.oe_del = o_del + e_del; K.e_del = e_del; K.oe_ins = o_ins + e_ins; K.e_ins = e_ins;
    const int C = (qlen + 1 + G - 1) / G;
    RegLane<CMAX> S;
    reg_init(S, gl, C, qlen, query, smat, h0, K.oe_ins, e_ins);
    int maxsc = 0;
    for (int i = 0; i < 25; ++i) maxsc = maxsc > smat[i] ? maxsc : smat[i];
    int max_ins = (int)((double)(qlen * maxsc + end_bonus - o_ins) / e_ins + 1.);
    max_ins = max_ins > 1 ? max_ins : 1;
    w = w < max_ins ? w : max_ins;
    int max_del = (int)((double)(qlen * maxsc + end_bonus - o_del) / e_del + 1.);
    max_del = max_del > 1 ? max_del : 1;
    w = w < max_del ? w : max_del;
    int beg = 0, end = qlen, max = h0, max_i = -1, max_j = -1, max_ie = -1, gscore = -1, max_off = 0;
    unsigned long long cells = 0;
    for (int i = 0; i < tlen; ++i) {
        const int t = target[i];
        if (beg < i - w) beg = i - w;
        if (end > i + w + 1) end = i + w + 1;
        if (end > qlen) end = qlen;
        int h1_init = 0;
        if (beg == 0) { h1_init = h0 - (o_del + e_del * (i + 1)); if (h1_init < 0) h1_init = 0; }
        const int W = end - beg;
        cells += W > 0 ? W : 0;
        if (W <= 0) {                                 // no cell: m == 0, the reference leaves the loop (after its gscore bookkeeping)
            if (beg == qlen) { max_ie = gscore > h1_init ? max_ie : i; gscore = gscore > h1_init ? gscore : h1_init; }
            break;
        }
        int n, L;
        reg_phase_a(S, beg, end, t, K, n, L);
#pragma unroll
        for (int d = 1; d < G; d <<= 1) {            // inclusive max-plus scan
            int L2 = g.up(L, d), n2 = g.up(n, d);
            if (gl >= d) { int c = L2 - n * e_ins; L = c > L ? c : L; n += n2; }
        }
        int fin = g.up(L, 1);
        fin = gl == 0 ? 0 : (fin > 0 ? fin : 0);
        RegRowOut o;
        reg_phase_b(S, beg, end, t, fin, h1_init, K, o);
        reg_phase_c(S, gl, beg, g.up(o.hlast, 1));
        const int key = g.rmax(o.key);                  // max h, ties to the larger column: what `mj = m > h ? mj : j` keeps
        const int rm = key >> 10, rmj = (key & 1023) - 1;
        const int h1 = g.bcast(o.hend, (end - 1) / C);
        if (end == qlen) { max_ie = gscore > h1 ? max_ie : i; gscore = gscore > h1 ? gscore : h1; }
        if (rm == 0) break;
        if (rm > max) {
            max = rm; max_i = i; max_j = rmj;
            int k = rmj - i; k = k < 0 ? -k : k;
            max_off = max_off > k ? max_off : k;
        } else if (zdrop > 0) {
            if (i - max_i > rmj - max_j) { if (max - rm - ((i - max_i) - (rmj - max_j)) * e_del > zdrop) break; }
            else { if (max - rm - ((rmj - max_j) - (i - max_i)) * e_ins > zdrop) break; }
        }
        // first / last non-zero column in one reduction: halves (mx + 1, 0xFFFF - mn)
        int mn, mx;
        reg_mask_range(reg_nonzero_mask(S, beg, end), S.j0, end, mn, mx);
        unsigned pk = (unsigned)(mx + 1) << 16 | (unsigned)(0xFFFF - (mn > 0xFFFF ? 0xFFFF : mn));
        for (int s = G >> 1; s > 0; s >>= 1) pk = __vmaxu2(pk, __shfl_xor_sync(g.mask, pk, s, G));
        int gmx = (int)(pk >> 16) - 1, gmn = 0xFFFF - (int)(pk & 0xFFFF);
        int nbeg = gmn < end ? gmn : end;
        int last = gmx >= nbeg ? gmx : nbeg - 1;
        beg = nbeg;
        end = last + 2 < qlen ? last + 2 : qlen;
    }
    if (gl == 0) { ctr.sw_cells += cells; ctr.n_ext++; }
    ExtResult R; R.score = max; R.qle = max_j + 1; R.tle = max_i + 1; R.gtle = max_ie + 1; R.gscore = gscore; R.max_off = max_off;
    return R;
}
#endif

} // namespace b200
