// reg_emul.cpp -- TEST HARNESS ONLY.  Lock-step CPU emulation of the register-resident lane-cooperative ksw_extend2
// (seqlib_b200/csrc/ksw_reg.cuh): the phase functions are the kernel's own HD code, called lane by lane; what the
// kernel does with warp shuffles between them (max-plus scan, reductions, the hand-over of H across lanes) is
// done here with plain loops over per-lane arrays.  Proves the scheme equal to the scalar recurrence (ksw.cuh
// extend2: score, qle, tle, gtle, gscore, max_off and the cell count) without a GPU.
#include <vector>
#include <cstdint>
#include <cstring>
#include <algorithm>
#include "../../seqlib_b200/csrc/ksw_reg.cuh"

using namespace b200;

struct DummyCtr { unsigned long long sw_cells = 0, n_ext = 0; };
static const int CMAX = 19;

static ExtResult extend2_reg_emul(int G, int qlen, const u8 *qp, int tlen, const u8 *target, const i8 *mat, int o_del, int e_del,
                                  int o_ins, int e_ins, int w, int end_bonus, int zdrop, int h0, unsigned long long *cells_out)
{
    RegConst K; K.oe_del = o_del + e_del; K.e_del = e_del; K.oe_ins = o_ins + e_ins; K.e_ins = e_ins;
    const int C = (qlen + 1 + G - 1) / G;
    std::vector<RegLane<CMAX> > S(G);
    BytesSeq query; query.p = qp; query.step = 1;
    for (int l = 0; l < G; ++l) reg_init(S[l], l, C, qlen, query, mat, h0, K.oe_ins, e_ins);
    int maxsc = 0;
    for (int i = 0; i < 25; ++i) maxsc = std::max<int>(maxsc, mat[i]);
    int max_ins = (int)((double)(qlen * maxsc + end_bonus - o_ins) / e_ins + 1.); max_ins = std::max(max_ins, 1); w = std::min(w, max_ins);
    int max_del = (int)((double)(qlen * maxsc + end_bonus - o_del) / e_del + 1.); max_del = std::max(max_del, 1); w = std::min(w, max_del);
    int beg = 0, end = qlen, max = h0, max_i = -1, max_j = -1, max_ie = -1, gscore = -1, max_off = 0;
    unsigned long long cells = 0;
    std::vector<int> n(G), L(G), fin(G);
    std::vector<RegRowOut> o(G);
    for (int i = 0; i < tlen; ++i) {
        const int t = target[i];
        if (beg < i - w) beg = i - w;
        if (end > i + w + 1) end = i + w + 1;
        if (end > qlen) end = qlen;
        int h1_init = 0;
        if (beg == 0) { h1_init = h0 - (o_del + e_del * (i + 1)); if (h1_init < 0) h1_init = 0; }
        const int W = end - beg;
        cells += W > 0 ? W : 0;
        if (W <= 0) {
            if (beg == qlen) { max_ie = gscore > h1_init ? max_ie : i; gscore = gscore > h1_init ? gscore : h1_init; }
            break;
        }
        for (int l = 0; l < G; ++l) reg_phase_a(S[l], beg, end, t, K, n[l], L[l]);
        for (int d = 1; d < G; d <<= 1) {            // Hillis-Steele, like shfl_up
            std::vector<int> n2(n), L2(L);
            for (int l = d; l < G; ++l) { int c = L2[l - d] - n[l] * e_ins; L[l] = std::max(c, L[l]); n[l] += n2[l - d]; }
        }
        for (int l = 0; l < G; ++l) fin[l] = l == 0 ? 0 : std::max(L[l - 1], 0);
        for (int l = 0; l < G; ++l) reg_phase_b(S[l], beg, end, t, fin[l], h1_init, K, o[l]);
        for (int l = G - 1; l >= 1; --l) reg_phase_c(S[l], l, beg, o[l - 1].hlast);
        int key = -1;
        for (int l = 0; l < G; ++l) key = std::max(key, o[l].key);
        const int rm = key >> 10, rmj = (key & 1023) - 1;
        const int h1 = o[(end - 1) / C].hend;
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
        unsigned hi = 0, lo = 0;
        for (int l = 0; l < G; ++l) {
            int mn, mx;
            reg_mask_range(reg_nonzero_mask(S[l], beg, end), S[l].j0, end, mn, mx);
            hi = std::max(hi, (unsigned)(mx + 1));
            lo = std::max(lo, (unsigned)(0xFFFF - (mn > 0xFFFF ? 0xFFFF : mn)));
        }
        int gmx = (int)hi - 1, gmn = 0xFFFF - (int)lo;
        int nbeg = gmn < end ? gmn : end;
        int last = gmx >= nbeg ? gmx : nbeg - 1;
        beg = nbeg;
        end = last + 2 < qlen ? last + 2 : qlen;
    }
    if (cells_out) *cells_out = cells;
    ExtResult R; R.score = max; R.qle = max_j + 1; R.tle = max_i + 1; R.gtle = max_ie + 1; R.gscore = gscore; R.max_off = max_off;
    return R;
}

extern "C" int reg_emul_check(int G, long n, const int *qlens, const int *tlens, const long *qoff, const long *toff, const u8 *qp, const u8 *tp,
                              const int *ws, const int *h0s, const i8 *mat, int o_del, int e_del, int o_ins, int e_ins, int end_bonus, int zdrop, long *first_bad)
{
    int bad = 0;
    for (long i = 0; i < n; ++i) {
        if (qlens[i] + 1 > G * CMAX) continue;
        std::vector<EH> eh(qlens[i] + 2);
        DummyCtr c;
        BytesSeq q; q.p = qp + qoff[i]; q.step = 1;
        BytesSeq t; t.p = tp + toff[i]; t.step = 1;
        ExtResult a = extend2(qlens[i], q, tlens[i], t, mat, o_del, e_del, o_ins, e_ins, ws[i], end_bonus, zdrop, h0s[i], eh.data(), c);
        unsigned long long cells = 0;
        ExtResult b = extend2_reg_emul(G, qlens[i], qp + qoff[i], tlens[i], tp + toff[i], mat, o_del, e_del, o_ins, e_ins, ws[i], end_bonus, zdrop, h0s[i], &cells);
        if (a.score != b.score || a.qle != b.qle || a.tle != b.tle || a.gtle != b.gtle || a.gscore != b.gscore || a.max_off != b.max_off || cells != c.sw_cells) {
            if (!bad && first_bad) *first_bad = i;
            ++bad;
        }
    }
    return bad;
}
