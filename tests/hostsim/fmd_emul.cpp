// fmd_emul.cpp -- TEST HARNESS ONLY.  The FMD-index rank code (fmd.cuh), the per-string overlap records (unitig.cuh), the
// seed-order walk (utg_walk.h) and the graph cleaning (mag_host.h) compiled for the host, over a BWT handed in by the test
// (the reference's own, or the model of tests/fmdmodel.py).  Lets the CPU suite check the device logic without a GPU.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/seqlib_b200.h"
#include "../../seqlib_b200/csrc/fmd.cuh"
#include "../../seqlib_b200/csrc/unitig.cuh"
#include "../../seqlib_b200/csrc/utg_walk.h"
#include "../../seqlib_b200/csrc/mag_host.h"

using namespace b200;

namespace {

struct HostFmd {
    std::vector<FmdBlock> blk;
    FmdIndex idx;
    void build(const uint8_t *bwt, uint64_t n)
    {
        uint64_t n_blk = (n + 127) / 128 + 1;
        blk.assign(n_blk, FmdBlock{});
        uint32_t run[4] = {0, 0, 0, 0};
        uint64_t tot[6] = {0, 0, 0, 0, 0, 0};
        for (uint64_t b = 0; b < n_blk; ++b) {
            FmdBlock &B = blk[b];
            for (int i = 0; i < 4; ++i) B.cnt[i] = run[i];
            for (int j = 0; j < 128; ++j) {
                uint64_t x = (b << 7) + j;
                if (x >= n) break;
                int s = bwt[x], h = j >> 6, t = j & 63;
                ++tot[s];
                if (s == 0) B.p2[h] |= 1ull << t;
                else { ++run[s - 1]; B.p0[h] |= (uint64_t)((s - 1) & 1) << t; B.p1[h] |= (uint64_t)((s - 1) >> 1) << t; }
            }
        }
        idx.blk = blk.data(); idx.n = n; idx.n_str = tot[0];
        idx.cnt[0] = 0;
        for (int c = 0; c < 6; ++c) idx.cnt[c + 1] = idx.cnt[c] + tot[c];
    }
};

} // namespace

extern "C" {

// rld_rank1a for a batch of positions over the given BWT
int fmd_emul_rank(const uint8_t *bwt, uint64_t n, int64_t n_q, const uint64_t *q, uint64_t *ranks, int32_t *sym)
{
    HostFmd F; F.build(bwt, n);
    for (int64_t i = 0; i < n_q; ++i) sym[i] = fmd_rank1a(F.idx, q[i], ranks + 6 * i);
    return 0;
}

// fml_fmi2mag (+ fml_mag_clean when stage >= 1) over the given BWT; returns a malloc'd mag_g_print text
char *fmd_emul_mag_text(const uint8_t *bwt, uint64_t n, const b200_fml_opt_t *opt, int stage, int64_t *text_len, float *rdist,
                        int64_t *stat /* [0] nodes with overlap record, [1] max neighbours, [2] max marks, [3] max intervals */)
{
    HostFmd F; F.build(bwt, n);
    const int cap = 4096, s_cap = 1 << 16, mark_cap = 4096;
    const bool two_phase = getenv("FMD_EMUL_TWO_PHASE") != nullptr;      // the loops the device's group kernel runs
    std::vector<u8> scratch(utg_scratch_bytes(cap, s_cap, mark_cap, two_phase) + 16);
    u8 *sp = (u8 *)(((uintptr_t)scratch.data() + 15) & ~(uintptr_t)15);
    UtgScratch S;
    utg_scratch_bind(S, sp, cap, s_cap, mark_cap, two_phase);
    std::vector<UtgNode> node(F.idx.n_str);
    std::vector<u8> seq; std::vector<UtgNei> nei; std::vector<UtgMark> mark;
    int64_t st[4] = {0, 0, 0, 0};
    for (u64 x = 0; x < F.idx.n_str; ++x) {
        UtgNode &N = node[x];
        if (two_phase) utg_node(F.idx, opt->min_asm_ovlp, x, S, N, UtgScalarCoop()); else utg_node(F.idx, opt->min_asm_ovlp, x, S, N, UtgNoCoop());
        N.seq_off = seq.size(); N.nei_off = nei.size(); N.mark_off = mark.size();
        seq.insert(seq.end(), S.s, S.s + N.len + N.ext_len);
        for (int i = 0; i < N.n_nei; ++i) nei.push_back(UtgNei{S.nei[i].x[0], S.nei[i].x[1], S.nei[i].x[2], (i64)S.nei[i].info});
        for (int i = 0; i < N.n_mark_r + N.n_mark_c; ++i) mark.push_back(S.mark[i]);
        if (!(N.flags & (UTG_CONTAINED | UTG_SHORT))) ++st[0];
        if (getenv("FMD_EMUL_CHECK_K") && N.len > 0 && N.ret_k != N.x0 + (x - N.x1)) ++st[3];
        if (N.n_nei > st[1]) st[1] = N.n_nei;
        if (N.n_mark_r + N.n_mark_c > st[2]) st[2] = N.n_mark_r + N.n_mark_c;
    }
    if (stat) memcpy(stat, st, sizeof(st));
    UtgPools P; P.node = node.data(); P.n_str = F.idx.n_str; P.seq = seq.data(); P.nei = nei.data(); P.mark = mark.data();
    Mag g;
    utg_walk_all(P, opt->min_asm_ovlp, opt->min_merge_len, g);
    if (rdist) *rdist = g.rdist;
    if (stage >= 100) g.fml_clean_steps(*opt, stage - 100);
    else if (stage >= 1) g.fml_clean(*opt);
    std::string t = g.text();
    char *out = (char *)malloc(t.size() + 1);
    memcpy(out, t.data(), t.size()); out[t.size()] = 0;
    *text_len = (int64_t)t.size();
    return out;
}

void fmd_emul_free(void *p) { free(p); }

} // extern "C"
