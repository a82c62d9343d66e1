// extend_group.cuh -- stage_extend with G lanes per read (device only).
// Control flow of mem_chain2aln (bwa/bwamem.c:658-812) is executed redundantly and uniformly by the G lanes of a
// group (it is a few dozen scalar decisions per seed); the two ksw_extend2 calls per seed are the cooperative
// extend2_group.  Memory written by one lane (region list, sort scratch) is published with a group barrier.
#pragma once
#include "pipeline.cuh"
#include "ksw_group.cuh"
#include "ksw_reg.cuh"
#include "ksw_wave.cuh"

namespace b200 {

// per-group shared memory: H[maxlen+2], E[maxlen+2] ints, then the read as bytes (maxlen, padded to 4)
__host__ __device__ inline size_t group_smem_bytes(int maxlen) { return (size_t)(maxlen + 2) * 8 + (size_t)((maxlen + 4) & ~3); }

// MODE 0: DP state in shared memory (ksw_group.cuh); 1: in registers, row-synchronous (ksw_reg.cuh; reads of at most
// G * EXT_REG_CMAX - 1 bases); 2: the packed 16-bit anti-diagonal wavefront (ksw_wave.cuh; H = the group's stream words;
// *retry is set when an extension must be re-run by the row-synchronous kernel, the read is then abandoned)
#define EXT_REG_CMAX 19
template <int G, int MODE, class Ctr>
__device__ void chain2aln_group(const GroupCtx<G> &g, const DevIndex &ix, const Opt &opt, const i8 *smat, int l_query, const u8 *query,
                                const Seed *cs, int cn, int c_rid, float c_frac_rep, RegSink &av, u64 *srt, int *H, int *E, Ctr &ctr,
                                bool *retry = nullptr)
{
    const bool REG = MODE != 0;
    int i, k, max_off[2], aw[2];
    i64 l_pac = ix.l_pac, rmax[2], tmp, max = 0;
    if (cn == 0) return;
    rmax[0] = l_pac << 1; rmax[1] = 0;
    for (i = 0; i < cn; ++i) {
        const Seed &t = cs[i];
        i64 b = t.rbeg - (t.qbeg + cal_max_gap(opt, t.qbeg));
        i64 e = t.rbeg + t.len + ((l_query - t.qbeg - t.len) + cal_max_gap(opt, l_query - t.qbeg - t.len));
        rmax[0] = rmax[0] < b ? rmax[0] : b;
        rmax[1] = rmax[1] > e ? rmax[1] : e;
        if (t.len > max) max = t.len;
    }
    rmax[0] = rmax[0] > 0 ? rmax[0] : 0;
    rmax[1] = rmax[1] < l_pac << 1 ? rmax[1] : l_pac << 1;
    if (rmax[0] < l_pac && l_pac < rmax[1]) {
        if (cs[0].rbeg < l_pac) rmax[1] = l_pac;
        else rmax[0] = l_pac;
    }
    {
        int is_rev;
        int rid = pos2rid(ix, depos(ix, cs[0].rbeg, &is_rev));
        i64 far_beg = ix.contig_off[rid], far_end = ix.contig_off[rid + 1];
        if (is_rev) { i64 t2 = far_beg; far_beg = (l_pac << 1) - far_end; far_end = (l_pac << 1) - t2; }
        rmax[0] = rmax[0] > far_beg ? rmax[0] : far_beg;
        rmax[1] = rmax[1] < far_end ? rmax[1] : far_end;
        if (g.gl == 0) ctr.ref_bytes += (unsigned long long)((rmax[1] - rmax[0] + 3) >> 2);
    }
    if (g.gl == 0) {
        for (i = 0; i < cn; ++i) srt[i] = (u64)cs[i].score << 32 | (u64)i;
        introsort((size_t)cn, srt, U64Less());
    }
    g.sync();
    for (k = cn - 1; k >= 0; --k) {
        const Seed s = cs[(u32)srt[k]];
        for (i = 0; i < av.n; ++i) {
            const Reg *p = &av.a[i];
            i64 rd; int qd, w, max_gap;
            if (s.rbeg < p->rb || s.rbeg + s.len > p->re || s.qbeg < p->qb || s.qbeg + s.len > p->qe) continue;
            if (s.len - p->seedlen0 > .1 * l_query) continue;
            qd = s.qbeg - p->qb; rd = s.rbeg - p->rb;
            max_gap = cal_max_gap(opt, qd < rd ? qd : (int)rd);
            w = max_gap < p->w ? max_gap : p->w;
            if (qd - rd < w && rd - qd < w) break;
            qd = p->qe - (s.qbeg + s.len); rd = p->re - (s.rbeg + s.len);
            max_gap = cal_max_gap(opt, qd < rd ? qd : (int)rd);
            w = max_gap < p->w ? max_gap : p->w;
            if (qd - rd < w && rd - qd < w) break;
        }
        if (i < av.n) {
            for (i = k + 1; i < cn; ++i) {
                if (srt[i] == 0) continue;
                const Seed *t = &cs[(u32)srt[i]];
                if (t->len < s.len * .95) continue;
                if (s.qbeg <= t->qbeg && s.qbeg + s.len - t->qbeg >= s.len >> 2 && t->qbeg - s.qbeg != t->rbeg - s.rbeg) break;
                if (t->qbeg <= s.qbeg && t->qbeg + t->len - s.qbeg >= s.len >> 2 && s.qbeg - t->qbeg != s.rbeg - t->rbeg) break;
            }
            if (i == cn) {
                g.sync();                      // everyone has finished reading srt[] for this seed
                if (g.gl == 0) srt[k] = 0;
                g.sync();
                continue;
            }
        }
        if (av.n >= av.cap) { av.overflow = true; return; }
        Reg a;
        a.rb = a.re = 0; a.qb = a.qe = a.rid = a.score = a.truesc = a.sub = a.alt_sc = a.csub = a.sub_n = 0;
        a.w = a.seedcov = a.secondary = a.secondary_all = a.seedlen0 = a.n_comp = a.is_alt = 0;
        a.frac_rep = 0; a.pad_ = 0; a.hash = 0;
        a.w = aw[0] = aw[1] = opt.w;
        a.score = a.truesc = -1;
        a.rid = c_rid;
        if (REG) {
            // Left and right extension share ONE call site: a group takes its pending sides in order (left first, the
            // right extension starts from the left one's score), so groups of a warp that are on different sides of
            // their seeds still run the DP rows together.
            const int qe = s.qbeg + s.len;
            const i64 re = s.rbeg + s.len - rmax[0];
            if (!s.qbeg) { a.score = a.truesc = s.len * opt.a; a.qb = 0; a.rb = s.rbeg; }
            if (qe == l_query) { a.qe = l_query; a.re = s.rbeg + s.len; }
            int pending = (s.qbeg ? 1 : 0) | (qe != l_query ? 2 : 0);
            while (pending) {
                const int side = (pending & 1) ? 0 : 1;
                pending &= ~(1 << side);
                const int sc0 = side ? a.score : s.len * opt.a;
                const int ql = side ? l_query - qe : s.qbeg;
                const int tl = side ? (int)(rmax[1] - rmax[0] - re) : (int)(s.rbeg - rmax[0]);
                const int pen = side ? opt.pen_clip3 : opt.pen_clip5;
                BytesSeq qs; qs.p = side ? query + qe : query + s.qbeg - 1; qs.step = side ? 1 : -1;
                int qle = 0, tle = 0, gtle = 0, gscore = 0;
#pragma unroll 1
                for (i = 0; i < B200_MAX_BAND_TRY; ++i) {
                    int prev = a.score;
                    aw[side] = opt.w << i;
                    TextSeqC rc(&ix, side ? rmax[0] + re : s.rbeg - 1, side ? 1 : -1);
                    ExtResult r;
                    if (MODE == 2) {
                        if (!wave_eligible(ql, tl, sc0, opt.a, pen) ||
                            !extend2_wave<G>(g, ql, qs, tl, rc, opt.a, opt.b, opt.o_del, opt.e_del, opt.o_ins, opt.e_ins, aw[side], pen, opt.zdrop, sc0, (u32 *)H, r, ctr)) {
                            *retry = true;
                            return;
                        }
                    } else r = extend2_reg<G, EXT_REG_CMAX>(g, ql, qs, tl, rc, smat, opt.o_del, opt.e_del, opt.o_ins, opt.e_ins, aw[side], pen, opt.zdrop, sc0, ctr);
                    a.score = r.score; qle = r.qle; tle = r.tle; gtle = r.gtle; gscore = r.gscore; max_off[side] = r.max_off;
                    if (a.score == prev || max_off[side] < (aw[side] >> 1) + (aw[side] >> 2)) break;
                }
                const bool local = gscore <= 0 || gscore <= a.score - pen;
                if (side == 0) {
                    if (local) { a.qb = s.qbeg - qle; a.rb = s.rbeg - tle; a.truesc = a.score; }
                    else { a.qb = 0; a.rb = s.rbeg - gtle; a.truesc = gscore; }
                } else {
                    if (local) { a.qe = qe + qle; a.re = rmax[0] + re + tle; a.truesc += a.score - sc0; }
                    else { a.qe = l_query; a.re = rmax[0] + re + gtle; a.truesc += gscore - sc0; }
                }
            }
        } else {
        if (s.qbeg) {
                int qle = 0, tle = 0, gtle = 0, gscore = 0;
                tmp = s.rbeg - rmax[0];
                BytesSeq qs; qs.p = query + s.qbeg - 1; qs.step = -1;
                TextSeq rs; rs.ix = &ix; rs.pos = s.rbeg - 1; rs.step = -1;
                for (i = 0; i < B200_MAX_BAND_TRY; ++i) {
                    int prev = a.score;
                    aw[0] = opt.w << i;
                    ExtResult r = extend2_group(g, s.qbeg, qs, (int)tmp, rs, smat, opt.o_del, opt.e_del, opt.o_ins, opt.e_ins,
                                                aw[0], opt.pen_clip5, opt.zdrop, s.len * opt.a, H, E, ctr);
                    a.score = r.score; qle = r.qle; tle = r.tle; gtle = r.gtle; gscore = r.gscore; max_off[0] = r.max_off;
                    if (a.score == prev || max_off[0] < (aw[0] >> 1) + (aw[0] >> 2)) break;
                }
                if (gscore <= 0 || gscore <= a.score - opt.pen_clip5) { a.qb = s.qbeg - qle; a.rb = s.rbeg - tle; a.truesc = a.score; }
                else { a.qb = 0; a.rb = s.rbeg - gtle; a.truesc = gscore; }
            } else { a.score = a.truesc = s.len * opt.a; a.qb = 0; a.rb = s.rbeg; }
            if (s.qbeg + s.len != l_query) {
                int qle = 0, tle = 0, qe, gtle = 0, gscore = 0, sc0 = a.score;
                i64 re;
                qe = s.qbeg + s.len;
                re = s.rbeg + s.len - rmax[0];
                BytesSeq qs; qs.p = query + qe; qs.step = 1;
                TextSeq rs; rs.ix = &ix; rs.pos = rmax[0] + re; rs.step = 1;
                for (i = 0; i < B200_MAX_BAND_TRY; ++i) {
                    int prev = a.score;
                    aw[1] = opt.w << i;
                    ExtResult r = extend2_group(g, l_query - qe, qs, (int)(rmax[1] - rmax[0] - re), rs, smat, opt.o_del, opt.e_del, opt.o_ins, opt.e_ins,
                                                aw[1], opt.pen_clip3, opt.zdrop, sc0, H, E, ctr);
                    a.score = r.score; qle = r.qle; tle = r.tle; gtle = r.gtle; gscore = r.gscore; max_off[1] = r.max_off;
                    if (a.score == prev || max_off[1] < (aw[1] >> 1) + (aw[1] >> 2)) break;
                }
                if (gscore <= 0 || gscore <= a.score - opt.pen_clip3) { a.qe = qe + qle; a.re = rmax[0] + re + tle; a.truesc += a.score - sc0; }
                else { a.qe = l_query; a.re = rmax[0] + re + gtle; a.truesc += gscore - sc0; }
            } else { a.qe = l_query; a.re = s.rbeg + s.len; }
        }
        for (i = 0, a.seedcov = 0; i < cn; ++i) {
            const Seed &t = cs[i];
            if (t.qbeg >= a.qb && t.qbeg + t.len <= a.qe && t.rbeg >= a.rb && t.rbeg + t.len <= a.re) a.seedcov += t.len;
        }
        a.w = aw[0] > aw[1] ? aw[0] : aw[1];
        a.seedlen0 = s.len;
        a.frac_rep = c_frac_rep;
        if (g.gl == 0) av.a[av.n] = a;
        ++av.n;
        g.sync();
    }
}

// One group handles read `rid`.  scratch: per-group slot in HBM (srt + region list); smem: per-group shared memory.
template <int G, int MODE>
__device__ void stage_extend_group(const GroupCtx<G> &g, const DevIndex &ix, const Opt &opt, const Caps &caps, const Batch &B, i64 rid,
                                   u8 *scratch, u8 *smem, const i8 *smat, CtrLocal &ctr)
{
    const bool REG = MODE != 0;
    ReadRec &R = B.rec[rid];
    if (B.ovf[rid]) { if (g.gl == 0) { R.n_regs = 0; R.reg_off = 0; } return; }
    int len = (int)(B.seq_off[rid + 1] - B.seq_off[rid]);
    const u8 *seq = B.seq + B.seq_off[rid];
    int *H = (int *)smem, *E = H + (caps.maxlen + 2);
    const u8 *q = seq;                           // REG: the read is touched once per extension, straight from HBM
    if (!REG) {
        u8 *qs = (u8 *)(E + (caps.maxlen + 2));
        for (int j = g.gl; j < len; j += G) qs[j] = seq[j];
        q = qs;
    }
    u8 *p = scratch;
    u64 *srt = (u64 *)p; p += sizeof(u64) * (size_t)caps.seeds;
    RegSink av; av.a = (Reg *)p; av.n = 0; av.cap = caps.regs; av.overflow = false;
    g.sync();
    const Chain *oc = B.pool.chains + R.chain_off;
    const Seed *os = B.pool.seeds + R.seed_off;
    int n_chains = R.n_chains;
    float frac = R.frac_rep;
    bool retry = false;
    for (int i = 0; i < n_chains; ++i) {
        chain2aln_group<G, MODE>(g, ix, opt, smat, len, q, os + oc[i].head, oc[i].n, oc[i].rid, frac, av, srt, H, E, ctr, &retry);
        if (MODE == 2 && retry) {       // hand the read to the row-synchronous kernel (k_extend_group over B.retry_list)
            if (g.gl == 0) { R.n_regs = 0; R.reg_off = 0; B.retry_list[atomicAdd(B.n_retry, 1ull)] = (i32)rid; }
            return;
        }
        if (av.overflow) { if (g.gl == 0) { B.ovf[rid] |= OVF_REG; R.n_regs = 0; R.reg_off = 0; } return; }
    }
    i64 off = 0;
    if (g.gl == 0) off = pool_alloc(B.pool, POOL_REG, av.n);
    off = (i64)__shfl_sync(g.mask, (unsigned long long)off, 0, G);
    if (off < 0) { if (g.gl == 0) { B.ovf[rid] |= OVF_POOL; R.n_regs = 0; R.reg_off = 0; } return; }
    // copy the regions to the pool, word-parallel
    const u32 *src = (const u32 *)av.a; u32 *dst = (u32 *)(B.pool.regs + off);
    int words = av.n * (int)(sizeof(Reg) / 4);
    for (int k = g.gl; k < words; k += G) dst[k] = src[k];
    if (g.gl == 0) { R.n_regs = av.n; R.reg_off = off; }
}

__host__ __device__ inline size_t extend_group_scratch_bytes(const Caps &c) { return sizeof(u64) * (size_t)c.seeds + sizeof(Reg) * (size_t)c.regs + 64; }

} // namespace b200
