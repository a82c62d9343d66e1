// ksw_group.cuh -- lane-cooperative ksw_extend2 (bwa/ksw.c:416-515): G lanes of a warp share one extension.
//
// Row i of the reference is a left-to-right sweep whose only loop-carried value is F (and the running max).
// F(i,j+1) = max(F(i,j) - e_ins, max(M(i,j) - oe_ins, 0)) does not depend on H(i,.), so it is a max-plus prefix
// over the row: every lane folds its contiguous column chunk into (n, L) with out(x) = max(x - n*e_ins, L), a
// log2(G)-step shuffle scan gives each lane its carry-in, and a second sweep produces H, E, the row maximum and
// the band-trimming information exactly as the scalar loop does.  The per-column state (H shifted by one column,
// E) lives in shared memory and is updated in place, so cells outside the current band keep their stale values
// like the reference's eh[] array.  int32 throughout, like the reference.
// The same algorithm, lane loops made explicit, is proven equal to the scalar recurrence on the CPU in
// tests/hostsim/group_emul.cpp.
#pragma once
#include "common.cuh"
#include "ksw.cuh"

namespace b200 {

template <int G>
struct GroupCtx {
    unsigned mask;    // lanes of this group inside the warp
    int gl;           // lane index inside the group
    __device__ __forceinline__ void sync() const { __syncwarp(mask); }
    __device__ __forceinline__ int up(int v, int d) const { return __shfl_up_sync(mask, v, d, G); }
    __device__ __forceinline__ int bcast(int v, int src) const { return __shfl_sync(mask, v, src, G); }
    __device__ __forceinline__ int rmax(int v) const { for (int o = G >> 1; o > 0; o >>= 1) { int x = __shfl_xor_sync(mask, v, o, G); v = v > x ? v : x; } return v; }
    __device__ __forceinline__ int rmin(int v) const { for (int o = G >> 1; o > 0; o >>= 1) { int x = __shfl_xor_sync(mask, v, o, G); v = v < x ? v : x; } return v; }
};

// H, E: qlen + 2 ints each (shared or global), smat: the 25-entry score matrix in shared memory.
template <int G, class QSeq, class TSeq, class Ctr>
__device__ ExtResult extend2_group(const GroupCtx<G> &g, int qlen, const QSeq &query, int tlen, const TSeq &target, const i8 *smat,
                                   int o_del, int e_del, int o_ins, int e_ins, int w, int end_bonus, int zdrop, int h0,
                                   int *H, int *E, Ctr &ctr)
{
    const int NEG = -(1 << 29);
    const int gl = g.gl;
    int oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
    for (int j = gl; j <= qlen; j += G) {
        int v = h0 - oe_ins - (j - 1) * e_ins;
        H[j] = j == 0 ? h0 : (v > 0 ? v : 0);
        E[j] = 0;
    }
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
    g.sync();
    for (int i = 0; i < tlen; ++i) {
        const i8 *qrow = smat + target[i] * 5;
        if (beg < i - w) beg = i - w;
        if (end > i + w + 1) end = i + w + 1;
        if (end > qlen) end = qlen;
        int h1_init = 0;
        if (beg == 0) { h1_init = h0 - (o_del + e_del * (i + 1)); if (h1_init < 0) h1_init = 0; }
        const int W = end - beg;
        const int C = W > 0 ? (W + G - 1) / G : 0;
        cells += W > 0 ? W : 0;
        int j0 = beg + gl * C, j1 = j0 + C < end ? j0 + C : end;
        if (j0 > end) j0 = j1 = end;
        // ---- phase A: fold the chunk into (n, L); remember the first old H
        int L = NEG, n = j1 > j0 ? j1 - j0 : 0;
        int saved = j0 < j1 ? H[j0] : 0;
        for (int j = j0; j < j1; ++j) {
            int hp = H[j];
            int M = hp ? hp + qrow[query[j]] : 0;
            int t = M - oe_ins; t = t > 0 ? t : 0;
            L = L - e_ins > t ? L - e_ins : t;
        }
#pragma unroll
        for (int d = 1; d < G; d <<= 1) {            // inclusive max-plus scan
            int L2 = g.up(L, d), n2 = g.up(n, d);
            if (gl >= d) { int c = L2 - n * e_ins; L = c > L ? c : L; n += n2; }
        }
        int fin = g.up(L, 1);
        fin = gl == 0 ? 0 : (fin > 0 ? fin : 0);
        g.sync();                                     // every lane has read its `saved` before anyone writes H
        // ---- phase B: the true recurrence over the chunk, in place
        int f = fin, hp = saved, m = -1, mj = -1, mn = 1 << 30, mx = -1, hl = 0;
        if (gl == 0) { H[beg] = h1_init; if (h1_init != 0) { if (beg < end) mn = beg; mx = beg; } }
        for (int j = j0; j < j1; ++j) {
            int hp_next = j + 1 < j1 ? H[j + 1] : 0;
            int e = E[j];
            int M = hp ? hp + qrow[query[j]] : 0;
            int h = M > e ? M : e; h = h > f ? h : f;
            if (h >= m) { m = h; mj = j; }
            int t = M - oe_del; t = t > 0 ? t : 0;
            e -= e_del; e = e > t ? e : t;
            E[j] = e;
            t = M - oe_ins; t = t > 0 ? t : 0;
            f -= e_ins; f = f > t ? f : t;
            H[j + 1] = h;
            if (h != 0) { if (j + 1 < end && j + 1 < mn) mn = j + 1; mx = j + 1; }
            if (e != 0) { if (j < mn) mn = j; if (j > mx) mx = j; }
            hl = h; hp = hp_next;
        }
        // ---- row reductions
        int rm = 0, rmj = -1, h1 = h1_init;
        if (W > 0) {
            rm = g.rmax(m);
            rmj = g.rmax(m == rm && j1 > j0 ? mj : -1);
            h1 = g.bcast(hl, (W - 1) / C);
        }
        if (gl == 0) { E[end] = 0; if (W <= 0) H[end] = h1; }
        int jj = W > 0 ? end : beg;
        if (jj == qlen) { max_ie = gscore > h1 ? max_ie : i; gscore = gscore > h1 ? gscore : h1; }
        if (rm == 0) break;
        if (rm > max) {
            max = rm; max_i = i; max_j = rmj;
            int k = rmj - i; k = k < 0 ? -k : k;
            max_off = max_off > k ? max_off : k;
        } else if (zdrop > 0) {
            if (i - max_i > rmj - max_j) { if (max - rm - ((i - max_i) - (rmj - max_j)) * e_del > zdrop) break; }
            else { if (max - rm - ((rmj - max_j) - (i - max_i)) * e_ins > zdrop) break; }
        }
        int gmn = g.rmin(mn), gmx = g.rmax(mx);
        int nbeg = gmn < end ? gmn : end;
        int last = gmx >= nbeg ? gmx : nbeg - 1;
        beg = nbeg;
        end = last + 2 < qlen ? last + 2 : qlen;
        g.sync();                                     // row i's H/E visible before row i+1 re-chunks the band
    }
    g.sync();
    if (gl == 0) { ctr.sw_cells += cells; ctr.n_ext++; }
    ExtResult R; R.score = max; R.qle = max_j + 1; R.tle = max_i + 1; R.gtle = max_ie + 1; R.gscore = gscore; R.max_off = max_off;
    return R;
}

// Lane-cooperative ksw_global2 (bwa/ksw.c:540-642): same chunked rows, F by an exact max-plus scan (no zero clamp here:
// F(i,j+1) = max(F(i,j) - e_ins, M(i,j) - oe_ins) with F(i,beg) = MINUS_INF), direction bytes z[i*n_col + j-beg] written by
// the lane that owns the column.  z == NULL: score only.  Proven equal to the scalar loop (score and every direction
// byte) by tests/hostsim/group_emul.cpp.  Returns eh[qlen].h.
template <int G, class QSeq, class TSeq>
__device__ int global2_group(const GroupCtx<G> &g, int qlen, const QSeq &query, int tlen, const TSeq &target, const i8 *smat,
                             int o_del, int e_del, int o_ins, int e_ins, int w, int *H, int *E, u8 *z, unsigned long long *cells_out)
{
    const int MINF = KSW_MINUS_INF, SENT = -2147483000;
    const int gl = g.gl;
    int oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
    int n_col = qlen < 2 * w + 1 ? qlen : 2 * w + 1;
    for (int j = gl; j <= qlen; j += G) {
        if (j == 0) { H[0] = 0; E[0] = MINF; }
        else if (j <= w) { H[j] = -(o_ins + e_ins * j); E[j] = MINF; }
        else { H[j] = MINF; E[j] = MINF; }
    }
    unsigned long long cells = 0;
    g.sync();
    for (int i = 0; i < tlen; ++i) {
        const i8 *qrow = smat + target[i] * 5;
        int beg = i > w ? i - w : 0, end = i + w + 1 < qlen ? i + w + 1 : qlen;
        int h1_init = beg == 0 ? -(o_del + e_del * (i + 1)) : MINF;
        const int W = end - beg;
        const int C = W > 0 ? (W + G - 1) / G : 0;
        cells += W > 0 ? W : 0;
        int j0 = beg + gl * C, j1 = j0 + C < end ? j0 + C : end;
        if (j0 > end) j0 = j1 = end;
        int L = SENT, n = 0;
        int saved = j0 < j1 ? H[j0] : 0;
        for (int j = j0; j < j1; ++j) {
            int m = H[j] + qrow[query[j]];
            int t = m - oe_ins;
            L = n == 0 ? t : (L - e_ins > t ? L - e_ins : t);
            ++n;
        }
#pragma unroll
        for (int d = 1; d < G; d <<= 1) {
            int L2 = g.up(L, d), n2 = g.up(n, d);
            if (gl >= d) {
                if (n == 0) { L = L2; n = n2; }
                else if (n2 != 0) { int c = L2 - n * e_ins; L = c > L ? c : L; n += n2; }
            }
        }
        int Lp = g.up(L, 1), np = g.up(n, 1);
        int fin;
        if (gl == 0 || np == 0) fin = MINF;
        else { int c = MINF - np * e_ins; fin = c > Lp ? c : Lp; }
        g.sync();
        int f = fin, hp = saved, hl = 0;
        if (gl == 0) H[beg] = h1_init;
        u8 *zi = z ? z + (i64)i * n_col - beg : (u8 *)0;
        for (int j = j0; j < j1; ++j) {
            int hp_next = j + 1 < j1 ? H[j + 1] : 0;
            int m = hp + qrow[query[j]], e = E[j];
            int d = m >= e ? 0 : 1;
            int h = m >= e ? m : e;
            d = h >= f ? d : 2;
            h = h >= f ? h : f;
            int t = m - oe_del;
            e -= e_del;
            d |= e > t ? 1 << 2 : 0;
            e = e > t ? e : t;
            E[j] = e;
            t = m - oe_ins;
            f -= e_ins;
            d |= f > t ? 2 << 4 : 0;
            f = f > t ? f : t;
            if (zi) zi[j] = (u8)d;
            H[j + 1] = h;
            hl = h; hp = hp_next;
        }
        (void)hl;
        if (gl == 0) { if (W <= 0) H[end] = h1_init; E[end] = MINF; }
        g.sync();
    }
    int score = H[qlen];
    g.sync();
    if (cells_out && gl == 0) *cells_out += cells;
    return score;
}

} // namespace b200
