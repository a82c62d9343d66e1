// group_emul.cpp -- TEST HARNESS ONLY.  Lock-step CPU emulation of the lane-cooperative ksw_extend2
// (seqlib_b200/csrc/ksw_group.cuh): G lanes share one extension, each row is split into contiguous column chunks,
// F is resolved with a max-plus scan over the lanes, band trimming with min/max reductions.  Every "phase" below is
// one straight-line region between two group barriers of the kernel; lanes are looped inside a phase.  Used to prove
// the algorithm equal to the scalar recurrence (ksw.cuh extend2) on millions of random tuples without a GPU.
#include <vector>
#include <cstdint>
#include <cstring>
#include <algorithm>
#include "../../seqlib_b200/csrc/ksw.cuh"

using namespace b200;

struct DummyCtr { unsigned long long sw_cells = 0, n_ext = 0; };

static ExtResult extend2_group_emul(int G, int qlen, const u8 *query, int tlen, const u8 *target, const i8 *mat, int o_del, int e_del,
                                    int o_ins, int e_ins, int w, int end_bonus, int zdrop, int h0, unsigned long long *cells_out)
{
    const int NEG = -(1 << 29);
    std::vector<int> H(qlen + 2), E(qlen + 2);
    int oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
    for (int j = 0; j <= qlen; ++j) { H[j] = j == 0 ? h0 : std::max(h0 - oe_ins - (j - 1) * e_ins, 0); E[j] = 0; }
    int maxsc = 0;
    for (int i = 0; i < 25; ++i) maxsc = std::max<int>(maxsc, mat[i]);
    int max_ins = (int)((double)(qlen * maxsc + end_bonus - o_ins) / e_ins + 1.); max_ins = std::max(max_ins, 1); w = std::min(w, max_ins);
    int max_del = (int)((double)(qlen * maxsc + end_bonus - o_del) / e_del + 1.); max_del = std::max(max_del, 1); w = std::min(w, max_del);
    int beg = 0, end = qlen, max = h0, max_i = -1, max_j = -1, max_ie = -1, gscore = -1, max_off = 0;
    unsigned long long cells = 0;
    std::vector<int> n(G), L(G), fin(G), saved(G), lm(G), lmj(G), lmin(G), lmax(G), hlast(G), j0v(G), j1v(G);
    for (int i = 0; i < tlen; ++i) {
        const i8 *qrow = mat + target[i] * 5;
        if (beg < i - w) beg = i - w;
        if (end > i + w + 1) end = i + w + 1;
        if (end > qlen) end = qlen;
        int h1_init = 0;
        if (beg == 0) { h1_init = h0 - (o_del + e_del * (i + 1)); if (h1_init < 0) h1_init = 0; }
        int W = end - beg;
        int C = W > 0 ? (W + G - 1) / G : 0;
        cells += W > 0 ? W : 0;
        // phase A: local F summaries, save first old H
        for (int l = 0; l < G; ++l) {
            int j0 = beg + l * C, j1 = std::min(j0 + C, end);
            if (C == 0) { j0 = j1 = beg; }
            j0v[l] = j0; j1v[l] = j1;
            int Ll = NEG, nl = 0;
            saved[l] = j0 < j1 ? H[j0] : 0;
            for (int j = j0; j < j1; ++j) {
                int hp = H[j];
                int M = hp ? hp + qrow[query[j]] : 0;
                int t = M - oe_ins; t = t > 0 ? t : 0;
                Ll = std::max(Ll - e_ins, t); ++nl;
            }
            n[l] = nl; L[l] = Ll;
        }
        // inclusive max-plus scan (Hillis-Steele, like shfl_up)
        for (int d = 1; d < G; d <<= 1) {
            std::vector<int> n2(n), L2(L);
            for (int l = d; l < G; ++l) { L[l] = std::max(L2[l - d] - n[l] * e_ins, L[l]); n[l] = n[l] + n2[l - d]; }
            // note: uses the pre-step n[l] in the L update, n updated after -- mirror in the kernel
        }
        for (int l = 0; l < G; ++l) fin[l] = l == 0 ? 0 : std::max(L[l - 1], 0);
        // ---- barrier ----
        // phase B: true recurrence per chunk, in place (old H[j+1] read before it is overwritten)
        for (int l = 0; l < G; ++l) {
            int j0 = j0v[l], j1 = j1v[l];
            int f = fin[l], hp = saved[l], m = -1, mj = -1, mn = 1 << 30, mx = -1, hl = 0;
            if (l == 0) { H[beg] = h1_init; if (h1_init != 0) { if (beg < end) mn = std::min(mn, beg); mx = std::max(mx, beg); } }
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
                if (h != 0) { if (j + 1 < end) mn = std::min(mn, j + 1); mx = std::max(mx, j + 1); }
                if (e != 0) { mn = std::min(mn, j); mx = std::max(mx, j); }
                hl = h; hp = hp_next;
            }
            lm[l] = m; lmj[l] = mj; lmin[l] = mn; lmax[l] = mx; hlast[l] = hl;
        }
        // reductions
        int m = 0, mj = -1, h1 = h1_init;
        if (W > 0) {
            m = -1;
            for (int l = 0; l < G; ++l) m = std::max(m, lm[l]);
            for (int l = 0; l < G; ++l) if (lm[l] == m && j0v[l] < j1v[l]) mj = std::max(mj, lmj[l]);
            h1 = hlast[(W - 1) / C];
        }
        E[end] = 0;   // (H[end] was written by the owner of column end-1, or is h1_init when W == 0 and beg == end)
        if (W <= 0) H[end] = h1;
        int jj = W > 0 ? end : beg;
        if (jj == qlen) { max_ie = gscore > h1 ? max_ie : i; gscore = gscore > h1 ? gscore : h1; }
        if (m == 0) break;
        if (m > max) {
            max = m; max_i = i; max_j = mj;
            int k = mj - i; k = k < 0 ? -k : k;
            max_off = max_off > k ? max_off : k;
        } else if (zdrop > 0) {
            if (i - max_i > mj - max_j) { if (max - m - ((i - max_i) - (mj - max_j)) * e_del > zdrop) break; }
            else { if (max - m - ((mj - max_j) - (i - max_i)) * e_ins > zdrop) break; }
        }
        int mn = 1 << 30, mx = -1;
        for (int l = 0; l < G; ++l) { mn = std::min(mn, lmin[l]); mx = std::max(mx, lmax[l]); }
        int nbeg = mn < end ? mn : end;
        int last = mx >= nbeg ? mx : nbeg - 1;
        beg = nbeg;
        end = last + 2 < qlen ? last + 2 : qlen;
    }
    if (cells_out) *cells_out = cells;
    ExtResult R; R.score = max; R.qle = max_j + 1; R.tle = max_i + 1; R.gtle = max_ie + 1; R.gscore = gscore; R.max_off = max_off;
    return R;
}

extern "C" int group_emul_check(int G, long n, const int *qlens, const int *tlens, const long *qoff, const long *toff, const u8 *qp, const u8 *tp,
                                const int *ws, const int *h0s, const i8 *mat, int o_del, int e_del, int o_ins, int e_ins, int end_bonus, int zdrop, long *first_bad)
{
    int bad = 0;
    for (long i = 0; i < n; ++i) {
        std::vector<EH> eh(qlens[i] + 2);
        DummyCtr c;
        BytesSeq q; q.p = qp + qoff[i]; q.step = 1;
        BytesSeq t; t.p = tp + toff[i]; t.step = 1;
        ExtResult a = extend2(qlens[i], q, tlens[i], t, mat, o_del, e_del, o_ins, e_ins, ws[i], end_bonus, zdrop, h0s[i], eh.data(), c);
        unsigned long long cells = 0;
        ExtResult b = extend2_group_emul(G, qlens[i], qp + qoff[i], tlens[i], tp + toff[i], mat, o_del, e_del, o_ins, e_ins, ws[i], end_bonus, zdrop, h0s[i], &cells);
        if (a.score != b.score || a.qle != b.qle || a.tle != b.tle || a.gtle != b.gtle || a.gscore != b.gscore || a.max_off != b.max_off || cells != c.sw_cells) {
            if (!bad && first_bad) *first_bad = i;
            ++bad;
        }
    }
    return bad;
}

// ---------------------------------------------------------------------------------------------------------
// lock-step emulation of global2_group (ksw_group.cuh): banded global alignment with traceback directions
static int global2_group_emul(int G, int qlen, const u8 *query, int tlen, const u8 *target, const i8 *mat, int o_del, int e_del, int o_ins, int e_ins,
                              int w, std::vector<u8> &z)
{
    const int MINF = KSW_MINUS_INF, SENT = -2147483000;
    std::vector<int> H(qlen + 2), E(qlen + 2);
    int oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
    int n_col = qlen < 2 * w + 1 ? qlen : 2 * w + 1;
    z.assign((size_t)n_col * tlen + 1, 0);
    H[0] = 0; E[0] = MINF;
    for (int j = 1; j <= qlen; ++j) { if (j <= w) { H[j] = -(o_ins + e_ins * j); E[j] = MINF; } else H[j] = E[j] = MINF; }
    std::vector<int> n(G), L(G), fin(G), saved(G), j0v(G), j1v(G);
    for (int i = 0; i < tlen; ++i) {
        const i8 *qrow = mat + target[i] * 5;
        int beg = i > w ? i - w : 0, end = i + w + 1 < qlen ? i + w + 1 : qlen;
        int h1_init = beg == 0 ? -(o_del + e_del * (i + 1)) : MINF;
        int W = end - beg, C = W > 0 ? (W + G - 1) / G : 0;
        for (int l = 0; l < G; ++l) {
            int j0 = beg + l * C, j1 = std::min(j0 + C, end);
            if (j0 > end) j0 = j1 = end;
            j0v[l] = j0; j1v[l] = j1;
            int Ll = SENT, nl = 0;
            saved[l] = j0 < j1 ? H[j0] : 0;
            for (int j = j0; j < j1; ++j) { int m = H[j] + qrow[query[j]]; int t = m - oe_ins; Ll = nl == 0 ? t : std::max(Ll - e_ins, t); ++nl; }
            n[l] = nl; L[l] = Ll;
        }
        for (int d = 1; d < G; d <<= 1) {
            std::vector<int> n2(n), L2(L);
            for (int l = d; l < G; ++l) {
                if (n[l] == 0) { L[l] = L2[l - d]; n[l] = n2[l - d]; }
                else if (n2[l - d] == 0) { /* keep */ }
                else { L[l] = std::max(L2[l - d] - n[l] * e_ins, L[l]); n[l] += n2[l - d]; }
            }
        }
        for (int l = 0; l < G; ++l) {
            if (l == 0 || n[l - 1] == 0) fin[l] = MINF - 0;
            else fin[l] = std::max(MINF - n[l - 1] * e_ins, L[l - 1]);
        }
        int hlast = h1_init;
        for (int l = 0; l < G; ++l) {
            int j0 = j0v[l], j1 = j1v[l];
            int f = fin[l], hp = saved[l];
            if (l == 0) H[beg] = h1_init;
            for (int j = j0; j < j1; ++j) {
                int hp_next = j + 1 < j1 ? H[j + 1] : 0;
                int m = hp + qrow[query[j]], e = E[j];
                u8 d = m >= e ? 0 : 1;
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
                z[(size_t)i * n_col + (j - beg)] = d;
                H[j + 1] = h;
                if (j == end - 1) hlast = h;
                hp = hp_next;
            }
        }
        if (W <= 0) H[end] = hlast;
        E[end] = MINF;
    }
    return H[qlen];
}

extern "C" int global_emul_check(int G, long n, const int *qlens, const int *tlens, const long *qoff, const long *toff, const u8 *qp, const u8 *tp,
                                 const int *ws, const i8 *mat, int o_del, int e_del, int o_ins, int e_ins, long *first_bad)
{
    int bad = 0;
    for (long i = 0; i < n; ++i) {
        int ql = qlens[i], tl = tlens[i], w = ws[i];
        std::vector<EH> eh(ql + 2);
        int n_col = ql < 2 * w + 1 ? ql : 2 * w + 1;
        std::vector<u8> z((size_t)n_col * tl + 1), z2;
        std::vector<u32> cg(ql + tl + 8);
        struct C2 { unsigned long long sw_cells = 0, n_global = 0; } c;
        BytesSeq q; q.p = qp + qoff[i]; q.step = 1;
        BytesSeq t; t.p = tp + toff[i]; t.step = 1;
        int nc = 0;
        int sa = global2(ql, q, tl, t, mat, o_del, e_del, o_ins, e_ins, w, eh.data(), z.data(), cg.data(), (int)cg.size(), &nc, c);
        int sb = global2_group_emul(G, ql, qp + qoff[i], tl, tp + toff[i], mat, o_del, e_del, o_ins, e_ins, w, z2);
        bool ok = sa == sb && memcmp(z.data(), z2.data(), (size_t)n_col * tl) == 0;
        if (!ok) { if (!bad && first_bad) *first_bad = i; ++bad; }
    }
    return bad;
}
