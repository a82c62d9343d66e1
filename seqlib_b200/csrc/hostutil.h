// hostutil.h -- small host-side helpers shared by the engine and the test harness.
#pragma once
#include <cmath>
#include <vector>
#include <algorithm>
#include "pipeline.cuh"

namespace b200 {

inline Opt opt_from_abi(const b200_mem_opt_t &o)
{
    Opt p;
    p.a = o.a; p.b = o.b; p.o_del = o.o_del; p.e_del = o.e_del; p.o_ins = o.o_ins; p.e_ins = o.e_ins;
    p.pen_clip5 = o.pen_clip5; p.pen_clip3 = o.pen_clip3; p.w = o.w; p.zdrop = o.zdrop;
    p.max_mem_intv = o.max_mem_intv; p.T = o.T; p.flag = o.flag; p.min_seed_len = o.min_seed_len;
    p.min_chain_weight = o.min_chain_weight; p.max_chain_extend = o.max_chain_extend; p.split_factor = o.split_factor;
    p.split_width = o.split_width; p.max_occ = o.max_occ; p.max_chain_gap = o.max_chain_gap;
    p.mask_level = o.mask_level; p.drop_ratio = o.drop_ratio; p.mask_level_redun = o.mask_level_redun;
    p.mapQ_coef_len = o.mapQ_coef_len; p.mapQ_coef_fac = o.mapQ_coef_fac;
    for (int i = 0; i < 25; ++i) p.mat[i] = o.mat[i];
    return p;
}

// log(i) from the host libm for every integer the MAPQ formula can see
// (mem_approx_mapq_se, bwa/bwamem.c:982-1006): alignment span, seed coverage, sub_n + 1.
inline std::vector<double> make_log_table(int maxlen, const Opt &opt)
{
    int n = 2 * maxlen + 8 * opt.w + 4096;
    std::vector<double> t(n);
    t[0] = 0.0;
    for (int i = 1; i < n; ++i) t[i] = std::log((double)i);
    return t;
}

// Scratch-slot capacities: the main pass (typical short reads) and the spill pass for reads that overflow it.
inline Caps default_caps(int maxlen, bool tiny = false)
{
    Caps c;
    c.maxlen = maxlen;
    if (tiny) { c.intv = 6; c.wchains = 3; c.wseeds = 6; c.seeds = 4; c.regs = 2; c.cigar = 3; c.md = 6; c.z = 1024; }
    // main-pass slots: sized so that reads inside interspersed repeats (hundreds of seeds, dozens of chains and regions) stay in the main
    // pass -- on a reference with a 10 % / 12 %-divergence repeat family 3.8 % of the reads overflowed slots a quarter this size and the
    // spill pass took 80 % of the time (scripts/repeat_probe.py); untouched slot space costs nothing but address range
    else { c.intv = 64; c.wchains = 256; c.wseeds = 640; c.seeds = 384; c.regs = 96; c.cigar = 24; c.md = 96; c.z = (i64)maxlen * 256; }
    return c;
}

// tier: 1 = the spill pass proper (up to 64 seeds per interval: every read of ordinary genomes), 2 = the last resort for
// repeat-saturated reads (every interval may contribute opt.max_occ seeds, bwa/bwamem.c:300-313), run with a handful of threads
inline Caps big_caps(int maxlen, const Opt &opt, int tier = 1)
{
    Caps c;
    c.maxlen = maxlen;
    c.intv = 4 * maxlen + 64;
    i64 per = tier >= 2 ? std::max(opt.max_occ, 1) : std::min(std::max(opt.max_occ, 1), 64);
    i64 ws = (i64)c.intv * per + 1024;
    c.wseeds = (int)std::min<i64>(ws, tier >= 2 ? (1 << 22) : (1 << 20));
    c.wchains = c.wseeds;
    c.seeds = c.wseeds;
    c.regs = std::min(c.seeds, tier >= 2 ? (1 << 20) : 8192);
    c.cigar = 2 * maxlen + 8;
    c.md = 4 * maxlen + 64;
    int w4 = opt.w << 2;
    i64 ncol = std::min<i64>(maxlen, 2 * (i64)w4 + 1);
    c.z = ncol * ((i64)maxlen + 2 * w4 + 64);
    return c;
}

} // namespace b200
