// utg_walk.h -- the order-dependent half of fermi-lite's unitig construction, host C++.
//
//   utg_walk_all  <- fml_fmi2mag_core / worker / unitig1 / unitig_unidir (fermi-lite/unitig.c:274-447)
//
// The reference visits the seeds in the order (prime * i) % n_strings and threads three bitmaps through the walks
// (`used`: strings already absorbed, `bend`: strings known to sit at a bifurcation, `visited`: unitig ends already
// emitted).  Every FM-index question a walk asks was answered per string by the device (unitig.cuh: UtgNode); here the
// walks are replayed in the reference's order by chasing those records -- a few memory reads per absorbed read.
#pragma once
#include <vector>
#include <string>
#include "unitig.cuh"
#include "mag_host.h"

namespace b200 {

struct UtgPools {            // host copies of what the node kernel produced
    const UtgNode *node; u64 n_str;
    const u8 *seq; const UtgNei *nei; const UtgMark *mark;
};

struct UtgWalker {
    const UtgPools &P;
    int min_match, min_merge_len;
    std::vector<u64> used, bend, visited;
    // state of the walk in progress (aux_t + the kstrings of thrdat_t)
    std::string s, cov;
    std::vector<MagEdge> nei;           // copy_nei's view of a->nei: (x[0], info)
    std::vector<int> dcov;              // pending coverage increments of the current direction, as a difference array

    UtgWalker(const UtgPools &p, int mm, int mml) : P(p), min_match(mm), min_merge_len(mml)
    {
        size_t w = (size_t)((p.n_str + 63) / 64);
        used.assign(w, 0); bend.assign(w, 0); visited.assign(w, 0);
    }
    static bool get(const std::vector<u64> &b, u64 x) { return b[x >> 6] >> (x & 63) & 1; }
    static void set(std::vector<u64> &b, u64 x) { b[x >> 6] |= 1ull << (x & 63); }
    void set_bits(u64 x0, u64 x1, u64 x2) { for (u64 k = 0; k < x2; ++k) { set(used, x0 + k); set(used, x1 + k); } }
    void apply_marks(const UtgNode &N, bool check_left_part)
    {
        const UtgMark *m = P.mark + N.mark_off + (check_left_part ? N.n_mark_r : 0);
        int n = check_left_part ? N.n_mark_c : N.n_mark_r;
        for (int i = 0; i < n; ++i) set_bits(m[i].x0, m[i].x1, m[i].x2);
    }

    // unitig_unidir.  Sentinel row x holds the string whose interval has x[1] = x (rows are ordered by the reverse
    // complement), so the record of "the string with x[0] = k" is node[its x[1]]: `cur` is always a ROW.
    int unidir(u64 cur, int beg0, u64 k0, u64 &end, int &is_loop)
    {
        int beg = beg0, ori_l = (int)s.size(), n_reads = 0;
        is_loop = 0;
        for (;;) {
            const UtgNode &N = P.node[cur];
            nei.clear();
            if (N.flags & (UTG_CONTAINED | UTG_SHORT | UTG_OVERFLOW)) throw std::logic_error("unitig walk reached a string without overlap record");
            apply_marks(N, false);                         // try_right's set_bits happen whatever it returns
            if (N.rbeg < 0) break;
            int rbeg = beg + N.rbeg;
            if (N.n_nei > 1) {
                const UtgNei *nn = P.nei + N.nei_off;
                for (int i = 0; i < N.n_nei; ++i) nei.push_back(MagEdge{nn[i].x0, (u64)nn[i].ovlp});
                set(bend, end); break;
            }
            const UtgNei &n0 = N.nei0;                     // the single neighbour, inline in the record
            __builtin_prefetch(&P.node[n0.x1]);            // the record the walk needs next, if this step is accepted
            __builtin_prefetch((const char *)&P.node[n0.x1] + 64);
            nei.push_back(MagEdge{n0.x0, (u64)n0.ovlp});
            u64 k = n0.x0;
            if (k == end) break;
            if (get(bend, k)) { set(bend, k); break; }
            apply_marks(N, true);
            if (N.cl < 0) { set(bend, k); break; }
            if (k == k0) { is_loop = 1; break; }
            if (n0.x1 == end) { nei.clear(); break; }
            if ((int)n0.ovlp < min_merge_len) break;
            end = n0.x1;
            set_bits(n0.x0, n0.x1, n0.x2);
            ++n_reads;
            if (N.ext_len <= 8) s.append((const char *)N.ext8, (size_t)N.ext_len);
            else s.append((const char *)(P.seq + N.seq_off + N.len), (size_t)N.ext_len);
            // the reference bumps cov[rbeg, ori_l) by one, saturating at '~': saturating adds commute, so the bumps are
            // collected as a difference array and applied once per direction (flush_cov)
            dcov.resize(s.size() + 1, 0);
            ++dcov[rbeg]; --dcov[ori_l];
            cov.append((size_t)N.ext_len, '"');
            beg = rbeg; ori_l = (int)s.size();
            cur = n0.x1;
        }
        flush_cov();
        return n_reads;
    }
    void flush_cov()
    {
        int run = 0;
        for (size_t i = 0; i < dcov.size() && i < cov.size(); ++i) {
            run += dcov[i];
            if (run) { int c = (unsigned char)cov[i] + run; cov[i] = (char)(c > '~' ? '~' : c); }
        }
        dcov.clear();
    }

    // unitig1 + the bookkeeping of worker(); returns true when a vertex was appended to g
    bool seed(u64 x, Mag &g)
    {
        if (get(used, x)) return false;
        const UtgNode &N = P.node[x];
        if (N.flags & UTG_OVERFLOW) throw std::logic_error("unitig node overflowed its scratch");
        if (N.flags & UTG_DUP) return false;
        set_bits(N.x0, N.x1, N.x2);
        if (N.flags & UTG_CONTAINED) return false;
        if (N.flags & UTG_SHORT) return false;
        int n_reads = 1, is_loop = 0, seed_len = N.len;
        s.assign((const char *)(P.seq + N.seq_off), (size_t)N.len);
        cov.assign((size_t)N.len, '"');
        u64 end[2] = {N.x1, N.x0};
        std::vector<MagEdge> nei0, nei1;
        bool done = false;
        // (the reference skips this call when the seed has no overlap candidate at all; the call is then a no-op)
        n_reads += unidir(x, 0, N.x0, end[0], is_loop);
        nei0 = nei;
        if (is_loop) { nei1.push_back(MagEdge{end[0], nei[0].y}); done = true; }
        if (!done) {
            Mag::revcomp6(s); Mag::reverse(cov);
            // the string at the other end is the seed's reverse complement: x[0] = N.x1, found in row N.x0
            n_reads += unidir(N.x0, (int)s.size() - seed_len, N.x1, end[1], is_loop);
            nei1 = nei;
        }
        // worker(): keep the unitig unless one of its ends was emitted before
        if (get(visited, end[0])) return false;
        set(visited, end[0]);
        if (get(visited, end[1])) return false;
        set(visited, end[1]);
        g.v.emplace_back();
        MagV &z = g.v.back();
        z.len = (int)s.size(); z.nsr = n_reads; z.k[0] = end[0]; z.k[1] = end[1];
        z.nei[0].swap(nei0); z.nei[1].swap(nei1);
        z.seq = s; z.cov = cov;
        return true;
    }
};

// fml_fmi2mag_core with n_threads = 1
inline void utg_walk_all(const UtgPools &P, int min_match, int min_merge_len, Mag &g)
{
    static const u64 primes[] = {123457, 234571, 345679, 456791, 567899, 0};
    g.v.clear();
    if (P.n_str == 0) return;
    u64 prime = 0;
    for (int j = 0; primes[j] > 0; ++j) if (P.n_str % primes[j] != 0) { prime = primes[j]; break; }
    if (!prime) throw std::logic_error("no usable prime for the seed order");
    UtgWalker W(P, min_match, min_merge_len);
    for (u64 i = 0; i < P.n_str; ++i) W.seed((prime * i) % P.n_str, g);
    g.build_hash();
    g.amend();
    g.rdist = (float)g.cal_rdist();
}

} // namespace b200
