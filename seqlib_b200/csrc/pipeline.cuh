// pipeline.cuh -- the four per-read stages and the HBM record layout between
// them.  Each stage function is what one GPU thread executes for one read; the
// __global__ wrappers in engine.cu only pick the read and the scratch slot.
//
//   stage_seed     reads  -> intervals          (mem_collect_intv)
//   stage_chain    intervals -> filtered chains (mem_chain + mem_chain_flt)
//   stage_extend   chains -> raw regions        (mem_chain2aln per chain)
//   stage_finalize regions -> hits              (mem_sort_dedup_patch, mem_mark_primary_se, mem_reg2aln)
//
// Per-read slots have fixed capacities (Caps).  A read that does not fit sets
// its overflow bits and is re-run by the host driver with the large Caps.
#pragma once
#include "common.cuh"
#include "seed.cuh"
#include "chain.cuh"
#include "extend.cuh"
#include "finalize.cuh"
#include "../../include/seqlib_b200.h"

namespace b200 {

struct Caps {
    int intv;        // intervals per read
    int wchains;     // chains while chaining (scratch)
    int wseeds;      // seeds while chaining (scratch)
    int chains;      // kept chains per read (output of stage_chain)
    int seeds;       // seeds of kept chains per read
    int regs;        // regions per read
    int hits;        // hits per read
    int cigar;       // cigar words per hit
    int md;          // md bytes per hit (incl. NUL)
    int maxlen;      // longest read in the batch
    i64 z;           // direction bytes per scratch slot
};

// Views of the per-batch HBM buffers.  Read i of the batch owns slot i of every array.
struct Batch {
    i64 n_reads;
    const u8 *seq;           // nt4 codes, concatenated
    const i64 *seq_off;      // n_reads + 1
    const i64 *hash_id;      // n_reads
    const i32 *order;        // optional indirection: slot i processes read order[i] (NULL = identity)
    u32 *ovf;                // per read overflow bits (indexed by read id)
    // stage outputs
    Intv *intv; i32 *n_intv;
    Chain *chains; Seed *seeds; i32 *n_chains; float *frac_rep;
    Reg *regs; i32 *n_regs;
    b200_hit_t *hits; u32 *cigar; char *md; i32 *n_hits;
};

struct CtrLocal {
    unsigned long long occ_blocks, sa_reads, ref_bytes, sw_cells, n_ext, n_global;
    HD CtrLocal() : occ_blocks(0), sa_reads(0), ref_bytes(0), sw_cells(0), n_ext(0), n_global(0) {}
};

// ------------------------------------------------------------------ seed
// scratch per slot: 2 * (maxlen + 1) Intv
HD size_t seed_scratch_bytes(const Caps &c) { return sizeof(Intv) * 2 * (size_t)(c.maxlen + 1); }

HD void stage_seed(const DevIndex &ix, const Opt &opt, const Caps &caps, const Batch &B, i64 slot, i64 rid, u8 *scratch, CtrLocal &ctr)
{
    int len = (int)(B.seq_off[rid + 1] - B.seq_off[rid]);
    const u8 *seq = B.seq + B.seq_off[rid];
    IntvSink out; out.a = B.intv + slot * caps.intv; out.n = 0; out.cap = caps.intv; out.overflow = false;
    Intv *prev = (Intv *)scratch, *curr = prev + (caps.maxlen + 1);
    if (len >= opt.min_seed_len) collect_intv(ix, opt, len, seq, out, prev, curr, ctr);
    B.n_intv[slot] = out.overflow ? 0 : out.n;
    if (out.overflow) B.ovf[rid] |= OVF_INTV;
}

// ------------------------------------------------------------------ chain
HD size_t chain_scratch_bytes(const Caps &c)
{
    size_t nodes = (size_t)c.wchains / 4 + 8;
    return sizeof(Chain) * c.wchains + sizeof(Seed) * c.wseeds + sizeof(BtNode) * nodes + sizeof(i32) * 2 * (size_t)c.wchains + 64;
}

HD void stage_chain(const DevIndex &ix, const Opt &opt, const Caps &caps, const Batch &B, i64 slot, i64 rid, u8 *scratch, CtrLocal &ctr)
{
    int len = (int)(B.seq_off[rid + 1] - B.seq_off[rid]);
    B.n_chains[slot] = 0; B.frac_rep[slot] = 0.f;
    if (B.ovf[rid]) return;
    ChainWork w;
    size_t nodes = (size_t)caps.wchains / 4 + 8;
    u8 *p = scratch;
    w.chains = (Chain *)p; p += sizeof(Chain) * caps.wchains; w.cap_chains = caps.wchains;
    w.seeds = (Seed *)p; p += sizeof(Seed) * caps.wseeds; w.cap_seeds = caps.wseeds;
    w.nodes = (BtNode *)p; p += sizeof(BtNode) * nodes; w.cap_nodes = (int)nodes;
    i32 *order = (i32 *)p; p += sizeof(i32) * caps.wchains;
    i32 *kept = (i32 *)p;
    int n_order = 0;
    int l_rep = build_chains(ix, opt, len, B.intv + slot * caps.intv, B.n_intv[slot], w, order, &n_order, ctr);
    if (w.ovf) { B.ovf[rid] |= w.ovf; return; }
    int n = filter_chains(opt, w, order, n_order, kept);
    // mem_flt_chained_seeds (bwa/bwamem.c:624-641) is a no-op below ~730 bp; longer reads are rejected by the host driver
    // linearise kept chains into the read's output slot
    Chain *oc = B.chains + slot * caps.chains;
    Seed *os = B.seeds + slot * caps.seeds;
    int ns = 0;
    if (n > caps.chains) { B.ovf[rid] |= OVF_CHAIN; return; }
    for (int i = 0; i < n; ++i) {
        const Chain &c = w.chains[order[i]];
        if (ns + c.n > caps.seeds) { B.ovf[rid] |= OVF_SEED; return; }
        oc[i] = c;
        oc[i].head = ns;
        for (int s = c.head; s >= 0; s = w.seeds[s].next) os[ns++] = w.seeds[s];
        oc[i].tail = ns - 1;
    }
    B.n_chains[slot] = n;
    B.frac_rep[slot] = (float)l_rep / len;
}

// ------------------------------------------------------------------ extend
// scratch per slot: (maxlen+1) EH + caps.seeds u64
HD size_t extend_scratch_bytes(const Caps &c) { return sizeof(EH) * (size_t)(c.maxlen + 2) + sizeof(u64) * (size_t)c.seeds; }

HD void stage_extend(const DevIndex &ix, const Opt &opt, const Caps &caps, const Batch &B, i64 slot, i64 rid, u8 *scratch, CtrLocal &ctr)
{
    B.n_regs[slot] = 0;
    if (B.ovf[rid]) return;
    int len = (int)(B.seq_off[rid + 1] - B.seq_off[rid]);
    const u8 *seq = B.seq + B.seq_off[rid];
    EH *eh = (EH *)scratch;
    u64 *srt = (u64 *)(scratch + sizeof(EH) * (size_t)(caps.maxlen + 2));
    RegSink av; av.a = B.regs + slot * caps.regs; av.n = 0; av.cap = caps.regs; av.overflow = false;
    const Chain *oc = B.chains + slot * caps.chains;
    const Seed *os = B.seeds + slot * caps.seeds;
    int n = B.n_chains[slot];
    float fr = B.frac_rep[slot];
    for (int i = 0; i < n; ++i) {
        chain2aln(ix, opt, len, seq, os + oc[i].head, oc[i].n, oc[i].rid, fr, av, srt, eh, ctr);
        if (av.overflow) { B.ovf[rid] |= OVF_REG; return; }
    }
    B.n_regs[slot] = av.n;
}

// ------------------------------------------------------------------ finalize
// scratch per slot: (maxlen + 2 + extra) EH + z bytes + 2*regs i32
HD size_t finalize_scratch_bytes(const Caps &c)
{
    return sizeof(EH) * (size_t)(c.maxlen + 2) + (size_t)c.z + sizeof(i32) * 2 * (size_t)c.regs + 64;
}

HD void stage_finalize(const DevIndex &ix, const Opt &opt, const Caps &caps, const Batch &B, i64 slot, i64 rid, u8 *scratch,
                       const double *log_tab, int n_log, CtrLocal &ctr)
{
    B.n_hits[slot] = 0;
    if (B.ovf[rid]) return;
    int len = (int)(B.seq_off[rid + 1] - B.seq_off[rid]);
    const u8 *seq = B.seq + B.seq_off[rid];
    FinScratch fs;
    u8 *p = scratch;
    fs.eh = (EH *)p; p += sizeof(EH) * (size_t)(caps.maxlen + 2);
    fs.z = p; fs.z_cap = caps.z; p += caps.z;
    p = (u8 *)(((uintptr_t)p + 7) & ~(uintptr_t)7);
    fs.zidx = (i32 *)p;
    fs.qbuf = 0;
    fs.log_tab = log_tab; fs.n_log = n_log;
    Reg *a = B.regs + slot * caps.regs;
    int n = B.n_regs[slot];
    n = sort_dedup_patch(ix, opt, seq, n, a, fs, ctr);
    for (int i = 0; i < n; ++i)                       // mem_align1_core tail (bwa/bwamem.c:1111-1115)
        if (a[i].rid >= 0 && ix.contig_alt[a[i].rid]) a[i].is_alt = 1;
    mark_primary_se(opt, n, a, B.hash_id[rid], fs.zidx);
    if (n > caps.hits) { B.ovf[rid] |= OVF_OUT; return; }
    b200_hit_t *H = B.hits + slot * caps.hits;
    for (int i = 0; i < n; ++i) {
        u32 *cg = B.cigar + (slot * caps.hits + i) * (i64)caps.cigar;
        char *md = B.md + (slot * caps.hits + i) * (i64)caps.md;
        AlnOut o = reg2aln(ix, opt, len, seq, &a[i], fs, cg, caps.cigar, md, caps.md, ctr);
        if (o.overflow || o.need_host) { B.ovf[rid] |= o.need_host ? OVF_SCRATCH : OVF_OUT; return; }
        b200_hit_t &h = H[i];
        const Reg &r = a[i];
        h.rb = r.rb; h.re = r.re; h.pos = o.pos; h.hash = r.hash; h.qb = r.qb; h.qe = r.qe; h.rid = r.rid;
        h.score = r.score; h.truesc = r.truesc; h.sub = r.sub; h.alt_sc = r.alt_sc; h.csub = r.csub; h.sub_n = r.sub_n;
        h.w = r.w; h.seedcov = r.seedcov; h.secondary = r.secondary; h.secondary_all = r.secondary_all;
        h.seedlen0 = r.seedlen0; h.n_comp = r.n_comp; h.is_alt = r.is_alt; h.frac_rep = r.frac_rep;
        h.flag = o.flag; h.is_rev = o.is_rev; h.mapq = o.mapq; h.NM = o.NM; h.aln_sub = o.sub;
        h.n_cigar = o.n_cigar; h.md_len = o.md_len; h.cigar_off = 0; h.md_off = 0;
        if (o.rid != r.rid) h.rid = -1000;   // the reference asserts equality (bwa/bwamem.c:1183)
    }
    B.n_hits[slot] = n;
}

} // namespace b200
