// hostindex.h -- TEST HARNESS ONLY: CPU construction of the device index image
// (same layout as DevIndex) from bwa-layout arrays, used by hostsim.
#pragma once
#include <vector>
#include <cstring>
#include "../../seqlib_b200/csrc/hostutil.h"
#include "../../seqlib_b200/csrc/seed2.cuh"

namespace b200 {

struct HostIndex {
    DevIndex dev;
    std::vector<OccBlock> occ;
    std::vector<u64> sa, text;
    std::vector<i64> coff;
    std::vector<i32> calt;
    std::vector<PIntv> tab_entries;      // prefix-interval tables (seed2.cuh), built like k_seedtab_level does
    SeedTab tab;

    struct NoCtr { unsigned long long occ_blocks = 0; };
    void build_tab(int K)
    {
        while (K > 0 && (1ull << (2 * K)) > dev.seq_len) --K;
        tab_entries.assign(seedtab_entries(K) + 2, PIntv());
        // the device table is 32-byte aligned; keep the same sector arithmetic valid on the host
        PIntv *base = tab_entries.data();
        if ((uintptr_t)base & 31) ++base;
        for (int j = 1; j <= K; ++j)
            for (u64 key = 0; key < (1ull << (2 * j)); ++key) {
                PIntv *out = base + seedtab_level_off(j) + key;
                if (j == 1) { Intv t; set_intv(dev, (int)key, t); *out = pintv_pack(t.x0, t.x1, t.x2, 0); continue; }
                u64 parent = key & ((1ull << (2 * (j - 1))) - 1);
                int c = (int)(key >> (2 * (j - 1)));
                u64 x0, x1, x2, na, no, ns; u32 e;
                pintv_unpack(base[seedtab_level_off(j - 1) + parent], x0, x1, x2, e);
                NoCtr ctr;
                extend_lean(dev, x1, x0, x2, 3 - c, na, no, ns, ctr);
                *out = pintv_pack(no, na, ns, 0);
            }
        tab.base = base; tab.K = K;
    }

    HostIndex(const b200_index_view_t &v, int sa_shift)
    {
        memset(&dev, 0, sizeof(dev));
        dev.primary = v.primary; for (int i = 0; i < 5; ++i) dev.L2[i] = v.L2[i];
        dev.seq_len = v.seq_len; dev.l_pac = v.l_pac;
        u64 N = v.seq_len, nblk = (N + 63) / 64 + 1;
        occ.assign(nblk, OccBlock());
        u32 run[4] = {0, 0, 0, 0};
        for (u64 b = 0; b < nblk; ++b) {
            OccBlock &o = occ[b];
            for (int c = 0; c < 4; ++c) o.cnt[c] = run[c];
            o.sym[0] = o.sym[1] = 0;
            for (int j = 0; j < 64; ++j) {
                u64 x = b * 64 + j;
                if (x >= N) break;
                // bwt_B0 (bwa/bwt.h:74-80)
                u32 wv = v.bwt[((x >> 7) << 4) + 8 + ((x & 0x7f) >> 4)];
                int s = (wv >> ((~x & 0xf) << 1)) & 3;
                occ_set_sym(o, j, (u64)s);
                ++run[s];
            }
        }
        (void)sa_shift;
        sa.assign(v.sa, v.sa + v.n_sa);
        int sh = 0; while ((1 << sh) < v.sa_intv) ++sh;
        dev.sa_shift = sh;
        text.assign((N + 31) / 32 + 1, 0);
        for (u64 p = 0; p < N; ++p) {
            i64 f = p < (u64)v.l_pac ? (i64)p : (i64)(N - 1 - p);
            int c = (v.pac[f >> 2] >> ((~f & 3) << 1)) & 3;
            if (p >= (u64)v.l_pac) c = 3 - c;
            text[p >> 5] |= (u64)c << (2 * (p & 31));
        }
        coff.resize(v.n_seqs + 1); calt.resize(v.n_seqs);
        for (int i = 0; i < v.n_seqs; ++i) { coff[i] = v.contigs[i].offset; calt[i] = v.contigs[i].is_alt; }
        coff[v.n_seqs] = v.l_pac;
        dev.occ = occ.data(); dev.n_occ = nblk; dev.sa = sa.data(); dev.n_sa = sa.size(); dev.text = text.data();
        dev.n_seqs = v.n_seqs; dev.contig_off = coff.data(); dev.contig_alt = calt.data();
        tab.base = nullptr; tab.K = 0;
    }
};

} // namespace b200
