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
    // one set of pools for the whole call, like one chunk on the device
    size_t pc[N_POOLS] = {(size_t)n * 64 + 4096, (size_t)n * 16 + 4096, (size_t)n * 64 + 4096, (size_t)n * 16 + 4096,
                          (size_t)n * 8 + 4096, (size_t)n * 64 + 4096, (size_t)n * 256 + 4096};
    std::vector<Intv> p_intv(pc[POOL_INTV] * 8); std::vector<Chain> p_chain(pc[POOL_CHAIN] * 8); std::vector<Seed> p_seed(pc[POOL_SEED] * 8);
    std::vector<Reg> p_reg(pc[POOL_REG] * 8); std::vector<b200_hit_t> p_hit(pc[POOL_HIT] * 8); std::vector<u32> p_cig(pc[POOL_CIGAR] * 8);
    std::vector<char> p_md(pc[POOL_MD] * 8);
    unsigned long long used[N_POOLS] = {0};
    std::vector<ReadRec> rec(n + 1); std::vector<u32> ovf(n + 1, 0);
    Batch B; memset(&B, 0, sizeof(B));
    B.n_reads = n; B.seq = seq.data(); B.seq_off = off; B.hash_id = ids; B.ovf = ovf.data(); B.rec = rec.data();
    B.pool.intv = p_intv.data(); B.pool.chains = p_chain.data(); B.pool.seeds = p_seed.data(); B.pool.regs = p_reg.data();
    B.pool.hits = p_hit.data(); B.pool.cigar = p_cig.data(); B.pool.md = p_md.data(); B.pool.used = used;
    B.pool.cap[POOL_INTV] = p_intv.size(); B.pool.cap[POOL_CHAIN] = p_chain.size(); B.pool.cap[POOL_SEED] = p_seed.size();
    B.pool.cap[POOL_REG] = p_reg.size(); B.pool.cap[POOL_HIT] = p_hit.size(); B.pool.cap[POOL_CIGAR] = p_cig.size(); B.pool.cap[POOL_MD] = p_md.size();
    bool dbg = getenv("HOSTSIM_DEBUG") != 0;
    int seed_v2 = getenv("HOSTSIM_SEED_V2") ? atoi(getenv("HOSTSIM_SEED_V2")) : 0;   // list capacity of the seed2.cuh machine, 0 = seed_fsm
    int tab_K = getenv("HOSTSIM_SEED_TAB_K") ? atoi(getenv("HOSTSIM_SEED_TAB_K")) : 0;   // prefix-interval tables for the seed2 machine
    if (seed_v2 && tab_K > 0) { if (hi->tab.K != tab_K) hi->build_tab(tab_K); }
    else { hi->tab.base = nullptr; hi->tab.K = 0; }
    const int seed_text = getenv("HOSTSIM_SEED_TEXT") ? atoi(getenv("HOSTSIM_SEED_TEXT")) : 0;      // the machine's text path (needs the full suffix array)
    if (seed_v2 && seed_text) { hi->densify_sa(); hi->tab.text = hi->verify_text() ? 1 : 0; if (!hi->tab.text) fprintf(stderr, "hostsim: text != text of the BWT\n"); }
    else hi->tab.text = 0;
    const SeedTab *tabp = seed_v2 && (tab_K > 0 || seed_text) ? &hi->tab : nullptr;
    std::vector<std::vector<Reg> > raws(n);
    unsigned long long tot_occ = 0, tot_tab = 0;
    for (int64_t r = 0; r < n; ++r) {
        for (int pass = 0; pass < 2; ++pass) {
            const Caps &c = caps[pass];
            CtrLocal ctr;
            std::vector<u8> s1(seed_scratch_bytes(c) + 64), s2(chain_scratch_bytes(c) + 64), s3(extend_scratch_bytes(c) + 64), s4(finalize_scratch_bytes(c) + 64);
            ovf[r] = 0;
            if (dbg) fprintf(stderr, "read %ld pass %d\n", (long)r, pass);
            if (seed_v2) stage_seed_v2(ix, opt, c, B, r, s1.data(), ctr, seed_v2, tabp);
            else stage_seed(ix, opt, c, B, r, s1.data(), ctr);
            tot_occ += ctr.occ_blocks; tot_tab += ctr.tab_hi;
            stage_chain(ix, opt, c, B, r, s2.data(), ctr, logtab.data(), (int)logtab.size());
            stage_extend(ix, opt, c, B, r, s3.data(), ctr);
            raws[r].assign(B.pool.regs + rec[r].reg_off, B.pool.regs + rec[r].reg_off + rec[r].n_regs);
            stage_finalize(ix, opt, c, B, r, s4.data(), logtab.data(), (int)logtab.size(), ctr);
            if (ovf[r] && pass == 0) { R->ovf[r] = ovf[r]; continue; }
            if (ovf[r]) R->ovf[r] |= 0x80000000u | ovf[r];
            break;
        }
        const ReadRec &rr = rec[r];
        for (int i = 0; i < rr.n_intv; ++i) R->intv.push_back(B.pool.intv[rr.intv_off + i]);
        for (int i = 0; i < rr.n_chains; ++i) {
            const Chain &ch = B.pool.chains[rr.chain_off + i];
            int64_t row[6] = {ch.pos, ch.rid, ch.w, ch.kept, ch.n, ch.first};
            R->chn.insert(R->chn.end(), row, row + 6);
            for (int s = 0; s < ch.n; ++s) {
                const Seed &q = B.pool.seeds[rr.seed_off + ch.head + s];
                int64_t srow[4] = {q.rbeg, q.qbeg, q.len, q.score};
                R->seeds.insert(R->seeds.end(), srow, srow + 4);
            }
            R->seed_off.push_back((int64_t)R->seeds.size() / 4);
        }
        for (size_t i = 0; i < raws[r].size(); ++i) {
            const Reg &q = raws[r][i];
            b200_hit_t h; memset(&h, 0, sizeof(h));
            h.rb = q.rb; h.re = q.re; h.qb = q.qb; h.qe = q.qe; h.rid = q.rid; h.score = q.score;
            h.truesc = q.truesc; h.w = q.w; h.seedcov = q.seedcov; h.seedlen0 = q.seedlen0; h.frac_rep = q.frac_rep;
            R->regs.push_back(h);
        }
        for (int i = 0; i < rr.n_hits; ++i) {
            b200_hit_t h = B.pool.hits[rr.hit_off + i];
            int64_t co = h.cigar_off, mo = h.md_off;
            h.cigar_off = (int64_t)R->cigar.size(); h.md_off = (int64_t)R->md.size();
            for (int k = 0; k < h.n_cigar; ++k) R->cigar.push_back(B.pool.cigar[co + k]);
            for (int k = 0; k <= h.md_len; ++k) R->md.push_back(B.pool.md[mo + k]);
            R->hits.push_back(h);
        }
        R->hit_off[r + 1] = (int64_t)R->hits.size();
        R->intv_off[r + 1] = (int64_t)R->intv.size();
        R->chn_off[r + 1] = (int64_t)R->chn.size() / 6;
        R->reg_off[r + 1] = (int64_t)R->regs.size();
    }
    if (getenv("HOSTSIM_STATS")) fprintf(stderr, "hostsim: seeding of %ld reads: %.1f Occ blocks, %.1f chain entries per read\n", (long)n, (double)tot_occ / (n ? n : 1), (double)tot_tab / (n ? n : 1));
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
