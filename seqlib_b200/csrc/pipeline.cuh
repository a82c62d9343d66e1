// pipeline.cuh -- the four per-read stages and the HBM record layout between
// them.  Each stage function is what one GPU thread executes for one read; the
// __global__ wrappers in engine.cu only pick the read and the scratch slot.
//
//   stage_seed     reads  -> intervals          (mem_collect_intv)
//   stage_chain    intervals -> filtered chains (mem_chain + mem_chain_flt)
//   stage_extend   chains -> raw regions        (mem_chain2aln per chain)
//   stage_finalize regions -> hits              (mem_sort_dedup_patch, mem_mark_primary_se, mem_reg2aln)
//
// A stage works inside its thread's scratch slot (capacities = Caps), then
// bump-allocates exactly what it produced from the chunk-wide pools and copies
// it there; ReadRec keeps (offset, count) per read.  A read that does not fit
// the scratch slot sets OVF bits and is re-run by the host driver in a spill
// pass with few threads and large slots; a full pool fails the chunk, which is
// retried with larger pools.
#pragma once
#include <cmath>
#include <cstdlib>
#include "common.cuh"
#include "seed.cuh"
#include "seed2.cuh"
#include "chain.cuh"
#include "seedsw.cuh"
#include "extend.cuh"
#include "finalize.cuh"
#include "../../include/seqlib_b200.h"

namespace b200 {

struct Caps {           // per-slot scratch capacities
    int intv;           // intervals per read
    int wchains;        // chains while chaining
    int wseeds;         // seeds while chaining
    int seeds;          // seeds of the kept chains (bounds the sort scratch of chain2aln)
    int regs;           // regions per read (= hits per read)
    int cigar;          // cigar words per hit
    int md;             // md bytes per hit (incl. NUL)
    int maxlen;         // longest read in the batch
    i64 z;              // direction bytes
};

enum { POOL_INTV = 0, POOL_CHAIN, POOL_SEED, POOL_REG, POOL_HIT, POOL_CIGAR, POOL_MD, N_POOLS };
enum { OVF_POOL = 64 };

struct Pools {
    Intv *intv; Chain *chains; Seed *seeds; Reg *regs; b200_hit_t *hits; u32 *cigar; char *md;
    i64 cap[N_POOLS];
    unsigned long long *used;   // N_POOLS bump counters
};

struct ReadRec {
    i64 intv_off, chain_off, seed_off, reg_off, hit_off;
    i32 n_intv, n_chains, n_seeds, n_regs, n_hits, n_cigar, n_md;
    float frac_rep;
};

struct DpJob { i64 hit; i32 rid, w2; };   // a hit whose CIGAR needs a banded global alignment

struct Batch {
    DpJob *dp_jobs; unsigned long long *n_dp_jobs; i64 cap_dp_jobs;   // NULL: resolve every hit inline
    i64 n_reads;
    const u8 *seq;           // nt4 codes, concatenated
    const i64 *seq_off;      // n_reads + 1
    const i64 *hash_id;      // n_reads
    u32 *ovf;                // per read overflow bits
    u32 *work;               // optional: per read estimate of extension work (for load-balanced scheduling)
    i32 *retry_list; unsigned long long *n_retry;    // reads the wavefront extension kernel hands to the row-synchronous one
    ReadRec *rec;            // per read
    Pools pool;
};

struct CtrLocal {
    unsigned long long occ_blocks, sa_reads, ref_bytes, sw_cells, n_ext, n_global, tab_lo, tab_hi;
    HD CtrLocal() : occ_blocks(0), sa_reads(0), ref_bytes(0), sw_cells(0), n_ext(0), n_global(0), tab_lo(0), tab_hi(0) {}
};

HD i64 pool_alloc(const Pools &P, int which, i64 n)
{
    if (n == 0) return 0;
#if defined(__CUDA_ARCH__)
    i64 off = (i64)atomicAdd(P.used + which, (unsigned long long)n);
#else
    i64 off = (i64)P.used[which]; P.used[which] += (unsigned long long)n;
#endif
    return off + n <= P.cap[which] ? off : -1;
}

HD u8 *align8(u8 *p) { return (u8 *)(((uintptr_t)p + 7) & ~(uintptr_t)7); }

// ------------------------------------------------------------------ seed
HD size_t seed_scratch_bytes(const Caps &c) { return sizeof(Intv) * (2 * (size_t)(c.maxlen + 1) + (size_t)c.intv); }

template <bool LOOPS>
HD void stage_seed_t(const DevIndex &ix, const Opt &opt, const Caps &caps, const Batch &B, i64 rid, u8 *scratch, CtrLocal &ctr)
{
    int len = (int)(B.seq_off[rid + 1] - B.seq_off[rid]);
    const u8 *seq = B.seq + B.seq_off[rid];
    ReadRec &R = B.rec[rid];
    R.n_intv = 0; R.intv_off = 0;
    if (B.ovf[rid]) return;
    Intv *prev = (Intv *)scratch, *curr = prev + (caps.maxlen + 1);
    IntvSink out; out.a = curr + (caps.maxlen + 1); out.n = 0; out.cap = caps.intv; out.overflow = false;
    if (len >= opt.min_seed_len) {
        collect_intv(ix, opt, len, seq, out, prev, curr, ctr);        // the reference's loop nest (seed.cuh)
    }
    if (out.overflow) { B.ovf[rid] |= OVF_INTV; return; }
    i64 off = pool_alloc(B.pool, POOL_INTV, out.n);
    if (off < 0) { B.ovf[rid] |= OVF_POOL; return; }
    for (int i = 0; i < out.n; ++i) B.pool.intv[off + i] = out.a[i];
    R.n_intv = out.n; R.intv_off = off;
}

HD void stage_seed(const DevIndex &ix, const Opt &opt, const Caps &caps, const Batch &B, i64 rid, u8 *scratch, CtrLocal &ctr)
{
    stage_seed_t<true>(ix, opt, caps, B, rid, scratch, ctr);
}

// The single-extension-site machine of seed2.cuh with a `list_cap`-entry work list, falling back to the reference-shaped
// loops when the read is not eligible or the list overflows: what k_seed2 + the spill pass compute together.
HD void stage_seed_v2(const DevIndex &ix, const Opt &opt, const Caps &caps, const Batch &B, i64 rid, u8 *scratch, CtrLocal &ctr, int list_cap,
                      const SeedTab *tab = nullptr)
{
    int len = (int)(B.seq_off[rid + 1] - B.seq_off[rid]);
    const u8 *seq = B.seq + B.seq_off[rid];
    ReadRec &R = B.rec[rid];
    R.n_intv = 0; R.intv_off = 0;
    if (B.ovf[rid]) return;
    Intv *prev = (Intv *)scratch, *curr = prev + (caps.maxlen + 1);
    IntvSink out; out.a = curr + (caps.maxlen + 1); out.n = 0; out.cap = caps.intv; out.overflow = false;
    while (list_cap > 2 * (caps.maxlen + 1) || (list_cap & (list_cap - 1))) list_cap &= list_cap - 1;      // a power of two that fits `prev` (16-byte entries)
    bool ok = seed2_eligible(ix, len, seq) && collect_intv_v2(ix, opt, len, seq, out, (PIntv *)prev, list_cap, ctr, tab);
    if (!ok && !out.overflow) { out.n = 0; if (len >= opt.min_seed_len) collect_intv(ix, opt, len, seq, out, prev, curr, ctr); }
    if (out.overflow) { B.ovf[rid] |= OVF_INTV; return; }
    i64 off = pool_alloc(B.pool, POOL_INTV, out.n);
    if (off < 0) { B.ovf[rid] |= OVF_POOL; return; }
    for (int i = 0; i < out.n; ++i) B.pool.intv[off + i] = out.a[i];
    R.n_intv = out.n; R.intv_off = off;
}

// ------------------------------------------------------------------ chain
HD size_t chain_nodes(const Caps &c) { return (size_t)c.wchains / 3 + 8; }
HD size_t chain_scratch_bytes(const Caps &c)
{
    return sizeof(Chain) * c.wchains + sizeof(Seed) * c.wseeds + sizeof(BtNode) * chain_nodes(c) + sizeof(i32) * 2 * (size_t)c.wchains + 64;
}

// mem_flt_chained_seeds (bwa/bwamem.c:624-641) over the kept chains; a no-op unless 5.5 ln(len) <= 0.05 len
HD void flt_chained_seeds(const DevIndex &ix, const Opt &opt, int len, const u8 *query, double log_len, ChainWork &w, const i32 *kept, int n)
{
    double min_l = opt.min_chain_weight ? (double)(1.1f * (float)opt.min_chain_weight) : (double)5.5f * log_len;
    int min_HSP_score = (int)(opt.a * min_l + .499);
    if (min_l > (double)(0.05f * (float)len)) return;
#if !defined(__CUDA_ARCH__)
    if (getenv("HOSTSIM_NO_SEEDSW")) return;        // test harness: lets the CPU suite prove the filter matters for its inputs
#endif
    for (int i = 0; i < n; ++i) {
        Chain &c = w.chains[kept[i]];
        int head = -1, tail = -1, k = 0;
        for (int s = c.head; s >= 0;) {
            Seed &sd = w.seeds[s];
            int next = sd.next;
            sd.score = seed_sw(ix, opt, len, query, sd.rbeg, sd.qbeg, sd.len);
            if (sd.score < 0 || sd.score >= min_HSP_score) {
                sd.score = sd.score < 0 ? sd.len * opt.a : sd.score;
                sd.next = -1;
                if (tail >= 0) w.seeds[tail].next = s; else head = s;
                tail = s; ++k;
            }
            s = next;
        }
        c.head = head; c.tail = tail; c.n = k;
    }
}

HD void stage_chain(const DevIndex &ix, const Opt &opt, const Caps &caps, const Batch &B, i64 rid, u8 *scratch, CtrLocal &ctr,
                    const double *log_tab = nullptr, int n_log = 0)
{
    int len = (int)(B.seq_off[rid + 1] - B.seq_off[rid]);
    ReadRec &R = B.rec[rid];
    R.n_chains = 0; R.n_seeds = 0; R.chain_off = R.seed_off = 0; R.frac_rep = 0.f;
    if (B.work) B.work[rid] = 0;
    if (B.ovf[rid]) return;
    ChainWork w;
    u8 *p = scratch;
    w.chains = (Chain *)p; p += sizeof(Chain) * caps.wchains; w.cap_chains = caps.wchains;
    w.seeds = (Seed *)p; p += sizeof(Seed) * caps.wseeds; w.cap_seeds = caps.wseeds;
    w.nodes = (BtNode *)p; p += sizeof(BtNode) * chain_nodes(caps); w.cap_nodes = (int)chain_nodes(caps);
    i32 *order = (i32 *)p; p += sizeof(i32) * caps.wchains;
    i32 *kept = (i32 *)p;
    int n_order = 0;
    int l_rep = build_chains(ix, opt, len, B.pool.intv + R.intv_off, R.n_intv, w, order, &n_order, ctr);
    if (w.ovf) { B.ovf[rid] |= w.ovf; return; }
    int n = filter_chains(opt, w, order, n_order, kept);
    {
#if defined(__CUDA_ARCH__)
        double lg = log_tab[len < n_log ? len : n_log - 1];       // host libm values (hostutil.h)
#else
        double lg = log_tab && len < n_log ? log_tab[len] : std::log((double)len);
#endif
        flt_chained_seeds(ix, opt, len, B.seq + B.seq_off[rid], lg, w, order, n);
    }
    int ns = 0;
    for (int i = 0; i < n; ++i) ns += w.chains[order[i]].n;
    if (ns > caps.seeds) { B.ovf[rid] |= OVF_SEED; return; }
    i64 coff = pool_alloc(B.pool, POOL_CHAIN, n), soff = pool_alloc(B.pool, POOL_SEED, ns);
    if (coff < 0 || soff < 0) { B.ovf[rid] |= OVF_POOL; return; }
    Chain *oc = B.pool.chains + coff;
    Seed *os = B.pool.seeds + soff;
    ns = 0;
    for (int i = 0; i < n; ++i) {           // linearise the kept chains
        const Chain &c = w.chains[order[i]];
        oc[i] = c;
        oc[i].head = ns;
        for (int s = c.head; s >= 0; s = w.seeds[s].next) os[ns++] = w.seeds[s];
        oc[i].tail = ns - 1;
    }
    R.n_chains = n; R.n_seeds = ns; R.chain_off = coff; R.seed_off = soff;
    R.frac_rep = (float)l_rep / len;
    if (B.work) {
        // Scheduling hint only (24 bits, sorted descending): the extension kernel runs several reads per warp in lock step, and
        // its groups stay together only while their DP problems have the same shape.  The shape of an extension is its query
        // length (the target window follows from it), so reads are grouped by the EXACT query lengths of the first seed
        // chain2aln will extend (left side first, then right), then by the query bases left over all seeds.
        u32 est = 0;
        for (int i = 0; i < ns; ++i) est += (u32)(len - os[i].len);
        u32 first = 0, second = 0;
        if (n > 0 && oc[0].n > 0) {
            const Seed *cs = os + oc[0].head;
            int best = 0;
            for (int i = 1; i < oc[0].n; ++i) if (cs[i].score >= cs[best].score) best = i;
            u32 ql = (u32)cs[best].qbeg, qr = (u32)(len - cs[best].qbeg - cs[best].len);
            first = ql ? ql : qr; second = ql ? qr : 0;
            if (first > 255) first = 255;
            if (second > 255) second = 255;
        }
        est >>= 2;
        B.work[rid] = first << 16 | second << 8 | (est < 255u ? est : 255u);
    }
}

// ------------------------------------------------------------------ extend
HD size_t extend_scratch_bytes(const Caps &c)
{
    return sizeof(EH) * (size_t)(c.maxlen + 2) + sizeof(u64) * (size_t)c.seeds + sizeof(Reg) * (size_t)c.regs + 64;
}

HD void stage_extend(const DevIndex &ix, const Opt &opt, const Caps &caps, const Batch &B, i64 rid, u8 *scratch, CtrLocal &ctr)
{
    ReadRec &R = B.rec[rid];
    R.n_regs = 0; R.reg_off = 0;
    if (B.ovf[rid]) return;
    int len = (int)(B.seq_off[rid + 1] - B.seq_off[rid]);
    const u8 *seq = B.seq + B.seq_off[rid];
    u8 *p = scratch;
    EH *eh = (EH *)p; p += sizeof(EH) * (size_t)(caps.maxlen + 2);
    u64 *srt = (u64 *)p; p += sizeof(u64) * (size_t)caps.seeds;
    RegSink av; av.a = (Reg *)p; av.n = 0; av.cap = caps.regs; av.overflow = false;
    const Chain *oc = B.pool.chains + R.chain_off;
    const Seed *os = B.pool.seeds + R.seed_off;
    for (int i = 0; i < R.n_chains; ++i) {
        chain2aln(ix, opt, len, seq, os + oc[i].head, oc[i].n, oc[i].rid, R.frac_rep, av, srt, eh, ctr);
        if (av.overflow) { B.ovf[rid] |= OVF_REG; return; }
    }
    i64 off = pool_alloc(B.pool, POOL_REG, av.n);
    if (off < 0) { B.ovf[rid] |= OVF_POOL; return; }
    for (int i = 0; i < av.n; ++i) B.pool.regs[off + i] = av.a[i];
    R.n_regs = av.n; R.reg_off = off;
}

// ------------------------------------------------------------------ finalize
HD size_t finalize_scratch_bytes(const Caps &c)
{
    return sizeof(EH) * (size_t)(c.maxlen + 2) + (size_t)c.z + 8 + sizeof(i32) * 2 * (size_t)c.regs + sizeof(u32) * (size_t)c.cigar + (size_t)c.md + 64;
}

HD void stage_finalize(const DevIndex &ix, const Opt &opt, const Caps &caps, const Batch &B, i64 rid, u8 *scratch,
                       const double *log_tab, int n_log, CtrLocal &ctr)
{
    ReadRec &R = B.rec[rid];
    R.n_hits = R.n_cigar = R.n_md = 0; R.hit_off = 0;
    if (B.ovf[rid]) return;
    int len = (int)(B.seq_off[rid + 1] - B.seq_off[rid]);
    const u8 *seq = B.seq + B.seq_off[rid];
    FinScratch fs;
    u8 *p = scratch;
    fs.eh = (EH *)p; p += sizeof(EH) * (size_t)(caps.maxlen + 2);
    fs.z = p; fs.z_cap = caps.z; p = align8(p + caps.z);
    fs.zidx = (i32 *)p; p += sizeof(i32) * 2 * (size_t)caps.regs;
    u32 *cg = (u32 *)p; p += sizeof(u32) * (size_t)caps.cigar;
    char *md = (char *)p;
    fs.qbuf = 0;
    fs.log_tab = log_tab; fs.n_log = n_log;
    Reg *a = B.pool.regs + R.reg_off;
    int n = sort_dedup_patch(ix, opt, seq, R.n_regs, a, fs, ctr);
    for (int i = 0; i < n; ++i)                       // mem_align1_core tail (bwa/bwamem.c:1111-1115)
        if (a[i].rid >= 0 && ix.contig_alt[a[i].rid]) a[i].is_alt = 1;
    mark_primary_se(opt, n, a, B.hash_id[rid], fs.zidx);
    i64 hoff = pool_alloc(B.pool, POOL_HIT, n);
    if (hoff < 0) { B.ovf[rid] |= OVF_POOL; return; }
    b200_hit_t *H = B.pool.hits + hoff;
    int tot_c = 0, tot_m = 0;
    for (int i = 0; i < n; ++i) {
        b200_hit_t &h = H[i];
        const Reg &r = a[i];
        h.rb = r.rb; h.re = r.re; h.hash = r.hash; h.qb = r.qb; h.qe = r.qe; h.rid = r.rid;
        h.score = r.score; h.truesc = r.truesc; h.sub = r.sub; h.alt_sc = r.alt_sc; h.csub = r.csub; h.sub_n = r.sub_n;
        h.w = r.w; h.seedcov = r.seedcov; h.secondary = r.secondary; h.secondary_all = r.secondary_all;
        h.seedlen0 = r.seedlen0; h.n_comp = r.n_comp; h.is_alt = r.is_alt; h.frac_rep = r.frac_rep;
        if (B.dp_jobs && r.rb >= 0 && r.re >= 0) {
            AlnPlan pl = reg2aln_plan(opt, &r, fs);
            if (pl.need_host) { B.ovf[rid] |= OVF_OUT; return; }
            if (!reg2aln_is_ungapped(&r, pl.w2)) {             // queue for k_finalize_dp
#if defined(__CUDA_ARCH__)
                i64 slot = (i64)atomicAdd(B.n_dp_jobs, 1ull);
#else
                i64 slot = (i64)(*B.n_dp_jobs)++;
#endif
                if (slot >= B.cap_dp_jobs) { B.ovf[rid] |= OVF_POOL; return; }
                DpJob j; j.hit = hoff + i; j.rid = (i32)rid; j.w2 = pl.w2;
                B.dp_jobs[slot] = j;
                h.flag = pl.flag; h.mapq = pl.mapq; h.pos = -1; h.is_rev = 0; h.NM = 0; h.aln_sub = 0; h.n_cigar = 0; h.md_len = 0;
                h.cigar_off = 0; h.md_off = 0;
                continue;
            }
        }
        AlnOut o = reg2aln(ix, opt, len, seq, &a[i], fs, cg, caps.cigar, md, caps.md, ctr);
        if (o.overflow || o.need_host) { B.ovf[rid] |= OVF_OUT; return; }
        i64 co = pool_alloc(B.pool, POOL_CIGAR, o.n_cigar), mo = pool_alloc(B.pool, POOL_MD, o.md_len + 1);
        if (co < 0 || mo < 0) { B.ovf[rid] |= OVF_POOL; return; }
        for (int k = 0; k < o.n_cigar; ++k) B.pool.cigar[co + k] = cg[k];
        for (int k = 0; k < o.md_len; ++k) B.pool.md[mo + k] = md[k];
        B.pool.md[mo + o.md_len] = 0;
        tot_c += o.n_cigar; tot_m += o.md_len + 1;
        h.pos = o.pos;
        h.flag = o.flag; h.is_rev = o.is_rev; h.mapq = o.mapq; h.NM = o.NM; h.aln_sub = o.sub;
        h.n_cigar = o.n_cigar; h.md_len = o.md_len; h.cigar_off = co; h.md_off = mo;
        if (o.rid != r.rid) h.rid = -1000;   // the reference asserts equality (bwa/bwamem.c:1183)
    }
    R.n_hits = n; R.n_cigar = tot_c; R.n_md = tot_m; R.hit_off = hoff;
}

} // namespace b200
