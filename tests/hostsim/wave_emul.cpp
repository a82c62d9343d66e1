// wave_emul.cpp -- TEST HARNESS ONLY.  Lock-step CPU emulation of the packed-16-bit anti-diagonal wavefront ksw_extend2
// (seqlib_b200/csrc/ksw_wave.cuh): WaveLane::setup / step and WaveAcc::commit are the kernel's own HD code, called lane
// by lane; the warp shuffle that hands a stream word to the next lane, the shared-memory column stream and the group
// votes of extend2_wave are plain loops here.  Proves the scheme equal to the scalar recurrence (ksw.cuh extend2) on
// every job it does not report as a gap event, and measures how often it does.
#include <vector>
#include <cstdint>
#include <cstring>
#include <algorithm>
#include "../../seqlib_b200/csrc/ksw_wave.cuh"

using namespace b200;

struct DummyCtrW { unsigned long long sw_cells = 0, n_ext = 0; };

// returns false on a gap event
static bool extend2_wave_emul(int G, int qlen, const u8 *query, int tlen, const u8 *target, int a, int b, int o_del, int e_del,
                              int o_ins, int e_ins, int w, int end_bonus, int zdrop, int h0, ExtResult &R, unsigned long long *cells_out)
{
    const WaveConst K = wave_const(a, b, o_del, e_del, o_ins, e_ins);
    w = wave_band(qlen, a, end_bonus, o_del, e_del, o_ins, e_ins, w);
    const int PAD = WAVE_PAD(G);
    std::vector<u32> ehs_store(WAVE_WORDS(qlen, G), 0u);
    u32 *const ehs = ehs_store.data() + PAD;                 // column j at ehs[j], guard words on both sides
    for (int j = 0; j <= qlen; ++j) {
        int v = h0 - (o_ins + e_ins) - (j - 1) * e_ins;
        v = j == 0 ? h0 : (v > 0 ? v : 0);
        ehs[j] = wave_word(v, 0, j < qlen ? (int)query[j] : 0, true);
    }
    WaveAcc acc; acc.init(h0);
    int cb = 0, xprev = qlen;
    unsigned long long cells = 0;
    u32 gapped = 0;
    std::vector<WaveLane> L(G);
    std::vector<u32> oh(G), win(G);
    for (int r0 = 0; r0 < tlen && !acc.broke; r0 += 2 * G) {
        for (int gl = 0; gl < G; ++gl) {
            const int rl = r0 + 2 * gl;
            L[gl].setup(rl, tlen, rl < tlen ? (int)target[rl] : 0, rl + 1 < tlen ? (int)target[rl + 1] : 0, cb, gl, w, qlen,
                        cb == 0 ? wave_h1_init(h0, o_del, e_del, rl) : 0, cb == 0 ? wave_h1_init(h0, o_del, e_del, rl + 1) : 0);
            oh[gl] = 0;
        }
        int cbn = 1 << 20;
        for (int s = 0, smax = qlen + 2 - cb + 2 * G; s <= smax; ++s) {
            for (int gl = G - 1; gl >= 1; --gl) win[gl] = oh[gl - 1];          // shfl_up by one lane
            win[0] = ehs[L[0].jl];
            bool all_done = true;
            for (int gl = 0; gl < G; ++gl) {
                const int jh = L[gl].jl - 1;
                oh[gl] = L[gl].step(K, win[gl]);
                if (gl == G - 1) ehs[jh] = oh[gl];
                all_done = all_done && L[gl].DONE == 0xffffffffu;
            }
            if (all_done) break;
        }
        const int xnew = (int)((L[G - 1].XC >> 16) & 0xffffu);
        for (int j = xnew + 1; j <= qlen; ++j) ehs[j] &= ~0x8000u;
        for (int j = cb; j <= qlen; ++j) if ((ehs[j] & 0x1fff3fffu) != 0) { cbn = j; break; }
        for (int gl = 0; gl < G; ++gl) gapped |= L[gl].gap;
        for (int k = 0; k < 2 * G; ++k) {
            const WaveLane &S = L[k >> 1];
            const int sh = (k & 1) * 16;
            const int m = (int)((S.MX >> sh) & 0xffffu);
            const int mj = (k & 1) ? S.mj_hi : S.mj_lo;
            const int x = (int)((S.XC >> sh) & 0xffffu);
            const int h1 = (int)((S.H1 >> sh) & 0xffffu);
            const int row = r0 + k;
            if (row < tlen && !acc.broke) {
                const int lo = cb > row - w ? cb : row - w;
                cells += x > lo ? (unsigned long long)(x - lo) : 0ull;
                acc.commit(row, lo, m, mj, x, h1, qlen, zdrop, e_del, e_ins);
            }
        }
        xprev = (int)((L[G - 1].XC >> 16) & 0xffffu);
        if (cbn < (1 << 20) && cbn > cb) cb = cbn;
    }
    if (cells_out) *cells_out = cells;
    R.score = acc.max; R.qle = acc.max_j + 1; R.tle = acc.max_i + 1; R.gtle = acc.max_ie + 1; R.gscore = acc.gscore; R.max_off = acc.max_off;
    return gapped == 0;
}

// returns the number of wrong results among the jobs the wavefront accepted; *n_gap = jobs it handed back, *n_run = jobs it ran
extern "C" int wave_emul_check(int G, long n, const int *qlens, const int *tlens, const long *qoff, const long *toff, const u8 *qp, const u8 *tp,
                               const int *ws, const int *h0s, const i8 *mat, int o_del, int e_del, int o_ins, int e_ins, int end_bonus, int zdrop,
                               long *first_bad, long *n_gap, long *n_run, double *cell_ratio)
{
    int bad = 0;
    long gaps = 0, run = 0;
    unsigned long long c_ref = 0, c_wave = 0;
    const int a = mat[0], b = -mat[1];
    for (long i = 0; i < n; ++i) {
        if (!wave_eligible(qlens[i], tlens[i], h0s[i], a, end_bonus)) continue;
        bool has_n = false;
        for (int k = 0; k < qlens[i]; ++k) has_n = has_n || qp[qoff[i] + k] > 3;
        for (int k = 0; k < tlens[i]; ++k) has_n = has_n || tp[toff[i] + k] > 3;
        if (has_n) continue;
        std::vector<EH> eh(qlens[i] + 2);
        DummyCtrW c;
        BytesSeq q; q.p = qp + qoff[i]; q.step = 1;
        BytesSeq t; t.p = tp + toff[i]; t.step = 1;
        ExtResult ex = extend2(qlens[i], q, tlens[i], t, mat, o_del, e_del, o_ins, e_ins, ws[i], end_bonus, zdrop, h0s[i], eh.data(), c);
        ExtResult got; unsigned long long cells = 0;
        ++run;
        if (!extend2_wave_emul(G, qlens[i], qp + qoff[i], tlens[i], tp + toff[i], a, b, o_del, e_del, o_ins, e_ins, ws[i], end_bonus, zdrop, h0s[i], got, &cells)) { ++gaps; continue; }
        c_ref += c.sw_cells; c_wave += cells;
        if (ex.score != got.score || ex.qle != got.qle || ex.tle != got.tle || ex.gtle != got.gtle || ex.gscore != got.gscore || ex.max_off != got.max_off) {
            if (!bad && first_bad) *first_bad = i;
            ++bad;
        }
    }
    if (n_gap) *n_gap = gaps;
    if (n_run) *n_run = run;
    if (cell_ratio) *cell_ratio = c_ref ? (double)c_wave / (double)c_ref : 0.0;
    return bad;
}
