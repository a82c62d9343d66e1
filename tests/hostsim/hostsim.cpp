// hostsim.cpp -- TEST HARNESS ONLY: runs the per-read device functions of
// seqlib_b200/csrc/*.cuh on the CPU (compiled by g++ with HD = inline), one read
// after the other, so the stage logic can be checked against the reference in a
// container without a GPU.  It is not part of libseqlib_b200.so and the product
// never falls back to it.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <string>
#include "../../seqlib_b200/csrc/pipeline.cuh"
#include "hostindex.h"

using namespace b200;

struct SimResults {
    std::vector<int64_t> hit_off;
    std::vector<b200_hit_t> hits;
    std::vector<uint32_t> cigar;
    std::vector<char> md;
    std::vector<uint32_t> ovf;
    std::vector<int64_t> intv_off; std::vector<Intv> intv;
    std::vector<int64_t> chn_off; std::vector<int64_t> chn; std::vector<int64_t> seed_off; std::vector<int64_t> seeds;
    std::vector<int64_t> reg_off; std::vector<b200_hit_t> regs;
};

extern "C" {

void *hostsim_index_from_view(const b200_index_view_t *v, int sa_shift)
{
    return new HostIndex(*v, sa_shift);
}
void hostsim_index_destroy(void *h) { delete (HostIndex *)h; }

static unsigned char nt4(unsigned char c)
{
    switch (c) { case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2; case 'T': case 't': return 3; default: return c < 4 ? c : 4; }
}

// caps_small != 0: run with deliberately small capacities first and re-run overflowing reads with big ones
void *hostsim_align(void *hidx, const b200_mem_opt_t *o, int64_t n, const char *seqs, const int64_t *off, const int64_t *ids, int small)
{
    HostIndex *hi = (HostIndex *)hidx;
    const DevIndex &ix = hi->dev;
    Opt opt = opt_from_abi(*o);
    SimResults *R = new SimResults;
    int maxlen = 1;
    for (int64_t i = 0; i < n; ++i) maxlen = std::max<int>(maxlen, (int)(off[i + 1] - off[i]));
    std::vector<u8> seq(off[n] + 1);
    for (int64_t i = 0; i < off[n]; ++i) seq[i] = nt4((unsigned char)seqs[i]);
    std::vector<double> logtab = make_log_table(maxlen, opt);
    Caps caps[2];
    caps[0] = default_caps(maxlen, small != 0);
    caps[1] = big_caps(maxlen, opt);
    R->hit_off.assign(n + 1, 0);
    R->ovf.assign(n, 0);
    R->intv_off.assign(n + 1, 0); R->chn_off.assign(n + 1, 0); R->reg_off.assign(n + 1, 0);
    R->seed_off.push_back(0);
    for (int64_t r = 0; r < n; ++r) {
        for (int pass = 0; pass < 2; ++pass) {
            const Caps &c = caps[pass];
            std::vector<Intv> intv(c.intv); std::vector<Chain> chains(c.chains); std::vector<Seed> sd(c.seeds); std::vector<Reg> regs(c.regs);
            std::vector<b200_hit_t> hits(c.hits); std::vector<u32> cg((size_t)c.hits * c.cigar); std::vector<char> md((size_t)c.hits * c.md);
            i32 n_intv = 0, n_chains = 0, n_regs = 0, n_hits = 0; float frac = 0; u32 ovf = 0;
            Batch B; memset(&B, 0, sizeof(B));
            int64_t one_off[2] = {0, off[r + 1] - off[r]};
            B.n_reads = 1; B.seq = seq.data() + off[r]; B.seq_off = one_off; B.hash_id = &ids[r]; B.ovf = &ovf;
            B.intv = intv.data(); B.n_intv = &n_intv; B.chains = chains.data(); B.seeds = sd.data(); B.n_chains = &n_chains; B.frac_rep = &frac;
            B.regs = regs.data(); B.n_regs = &n_regs; B.hits = hits.data(); B.cigar = cg.data(); B.md = md.data(); B.n_hits = &n_hits;
            CtrLocal ctr;
            std::vector<u8> s1(seed_scratch_bytes(c) + 64), s2(chain_scratch_bytes(c) + 64), s3(extend_scratch_bytes(c) + 64), s4(finalize_scratch_bytes(c) + 64);
            bool dbg = getenv("HOSTSIM_DEBUG") != 0;
            if (dbg) fprintf(stderr, "read %ld pass %d seed\n", (long)r, pass);
            stage_seed(ix, opt, c, B, 0, 0, s1.data(), ctr);
            if (dbg) fprintf(stderr, " n_intv %d ovf %u; chain\n", n_intv, ovf);
            stage_chain(ix, opt, c, B, 0, 0, s2.data(), ctr);
            if (dbg) fprintf(stderr, " n_chains %d ovf %u; extend\n", n_chains, ovf);
            std::vector<Reg> raw;
            stage_extend(ix, opt, c, B, 0, 0, s3.data(), ctr);
            raw.assign(regs.begin(), regs.begin() + n_regs);
            if (dbg) fprintf(stderr, " n_regs %d ovf %u; finalize\n", n_regs, ovf);
            stage_finalize(ix, opt, c, B, 0, 0, s4.data(), logtab.data(), (int)logtab.size(), ctr);
            if (ovf && pass == 0) { R->ovf[r] = ovf; continue; }
            if (ovf) { R->ovf[r] |= 0x80000000u | ovf; }
            for (int i = 0; i < n_intv; ++i) R->intv.push_back(intv[i]);
            for (int i = 0; i < n_chains; ++i) {
                const Chain &ch = chains[i];
                int64_t row[6] = {ch.pos, ch.rid, ch.w, ch.kept, ch.n, ch.first};
                R->chn.insert(R->chn.end(), row, row + 6);
                for (int s = 0; s < ch.n; ++s) {
                    const Seed &q = sd[ch.head + s];
                    int64_t srow[4] = {q.rbeg, q.qbeg, q.len, q.score};
                    R->seeds.insert(R->seeds.end(), srow, srow + 4);
                }
                R->seed_off.push_back((int64_t)R->seeds.size() / 4);
            }
            for (size_t i = 0; i < raw.size(); ++i) {
                b200_hit_t h; memset(&h, 0, sizeof(h));
                h.rb = raw[i].rb; h.re = raw[i].re; h.qb = raw[i].qb; h.qe = raw[i].qe; h.rid = raw[i].rid; h.score = raw[i].score;
                h.truesc = raw[i].truesc; h.w = raw[i].w; h.seedcov = raw[i].seedcov; h.seedlen0 = raw[i].seedlen0; h.frac_rep = raw[i].frac_rep;
                R->regs.push_back(h);
            }
            for (int i = 0; i < n_hits; ++i) {
                b200_hit_t h = hits[i];
                h.cigar_off = (int64_t)R->cigar.size(); h.md_off = (int64_t)R->md.size();
                for (int k = 0; k < h.n_cigar; ++k) R->cigar.push_back(cg[(size_t)i * c.cigar + k]);
                for (int k = 0; k <= h.md_len; ++k) R->md.push_back(md[(size_t)i * c.md + k]);
                R->hits.push_back(h);
            }
            break;
        }
        R->hit_off[r + 1] = (int64_t)R->hits.size();
        R->intv_off[r + 1] = (int64_t)R->intv.size();
        R->chn_off[r + 1] = (int64_t)R->chn.size() / 6;
        R->reg_off[r + 1] = (int64_t)R->regs.size();
    }
    return R;
}

void hostsim_results_view(void *h, b200_results_view_t *v)
{
    SimResults *R = (SimResults *)h;
    v->n_reads = (int64_t)R->hit_off.size() - 1; v->hit_off = R->hit_off.data(); v->hits = R->hits.data();
    v->cigar = R->cigar.data(); v->md = R->md.data();
    v->n_hits = (int64_t)R->hits.size(); v->n_cigar = (int64_t)R->cigar.size(); v->n_md = (int64_t)R->md.size();
}
const uint32_t *hostsim_ovf(void *h) { return ((SimResults *)h)->ovf.data(); }
void hostsim_stage_views(void *h, const int64_t **intv_off, const void **intv, const int64_t **chn_off, const int64_t **chn,
                         const int64_t **seed_off, const int64_t **seeds, const int64_t **reg_off, const void **regs)
{
    SimResults *R = (SimResults *)h;
    *intv_off = R->intv_off.data(); *intv = R->intv.data(); *chn_off = R->chn_off.data(); *chn = R->chn.data();
    *seed_off = R->seed_off.data(); *seeds = R->seeds.data(); *reg_off = R->reg_off.data(); *regs = R->regs.data();
}
void hostsim_results_free(void *h) { delete (SimResults *)h; }

} // extern "C"
