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
    std::vector<ChainEnt> tab_entries;   // prefix-chain table (seed2.cuh), built like k_seedtab_level / k_chain_build do
    SeedTab tab;

    struct NoCtr { unsigned long long occ_blocks = 0; };
    void build_tab(int K)
    {
        while (K > 0 && (1ull << (2 * K)) > dev.seq_len) --K;
        if (K < 4) { tab.base = nullptr; tab.K = 0; return; }
        std::vector<PIntv> lev(seedtab_entries(K - 1) + 1);
        for (int j = 1; j < K; ++j)
            for (u64 key = 0; key < (1ull << (2 * j)); ++key) {
                PIntv *out = lev.data() + seedtab_level_off(j) + key;
                if (j == 1) { Intv t; set_intv(dev, (int)key, t); *out = pintv_pack(t.x0, t.x1, t.x2, 0); continue; }
                u64 parent = key & ((1ull << (2 * (j - 1))) - 1);
                int c = (int)(key >> (2 * (j - 1)));
                u64 x0, x1, x2, na, no, ns; u32 e;
                pintv_unpack(lev[seedtab_level_off(j - 1) + parent], x0, x1, x2, e);
                NoCtr ctr;
                extend_lean(dev, x1, x0, x2, 3 - c, na, no, ns, ctr);
                *out = pintv_pack(no, na, ns, 0);
            }
        tab_entries.assign((1ull << (2 * K)) + 1, ChainEnt());
        // the device table is 32-byte aligned; keep the same sector arithmetic valid on the host
        ChainEnt *base = tab_entries.data();
        for (u64 key = 0; key < (1ull << (2 * K)); ++key) {
            u64 sz[18], x0 = 0, x1 = 0, x2 = 0, na, no, ns; u32 e;
            for (int m = 1; m < K; ++m) { pintv_unpack(lev[seedtab_level_off(m) + (key & ((1ull << (2 * m)) - 1))], x0, x1, x2, e); sz[m] = x2; }
            NoCtr ctr;
            extend_lean(dev, x1, x0, x2, 3 - (int)(key >> (2 * (K - 1))), na, no, ns, ctr);
            sz[K] = ns;
            base[key] = chain_make(K, no, na, ns, sz);
        }
        tab.base = base; tab.K = K;
    }
    // full suffix array (what the device index holds) from the sampled one
    void densify_sa()
    {
        if (dev.sa_shift == 0) return;
        std::vector<u64> d(dev.seq_len + 1);
        struct C0 { unsigned long long occ_blocks = 0, sa_reads = 0; } ctr;
        for (u64 k = 1; k <= dev.seq_len; ++k) d[k] = sa_lookup(dev, k, ctr);
        d[0] = ~0ull;
        sa.swap(d); dev.sa = sa.data(); dev.n_sa = sa.size(); dev.sa_shift = 0;
    }
    // the engine's k_verify_text on the host
    bool verify_text() const
    {
        if (dev.sa_shift != 0) return false;
        for (u64 k = 0; k <= dev.seq_len; ++k) {
            if (k == dev.primary) continue;
            const u64 p = k ? sa[k] : dev.seq_len;
            if (p == 0 || p > dev.seq_len) return false;
            const u64 x = k - (k > dev.primary);
            if (occ_sym(occ[x >> 6], (int)(x & 63)) != text_base(dev, (i64)(p - 1))) return false;
        }
        return true;
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
        tab.base = nullptr; tab.K = 0; tab.text = 0;
    }
};

} // namespace b200
