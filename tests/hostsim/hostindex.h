// hostindex.h -- TEST HARNESS ONLY: CPU construction of the device index image
// (same layout as DevIndex) from bwa-layout arrays, used by hostsim.
#pragma once
#include <vector>
#include <cstring>
#include "../../seqlib_b200/csrc/hostutil.h"

namespace b200 {

struct HostIndex {
    DevIndex dev;
    std::vector<OccBlock> occ;
    std::vector<u64> sa, text;
    std::vector<i64> coff;
    std::vector<i32> calt;

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
    }
};

} // namespace b200
