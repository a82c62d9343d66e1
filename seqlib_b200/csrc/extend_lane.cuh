// extend_lane.cuh -- stage_extend with ONE lane per read and a flattened DP loop.
//   control  <- mem_chain2aln (bwa/bwamem.c:658-812), cal_max_gap (:647-654)
//   DP       <- ksw_extend2   (bwa/ksw.c:416-515)
//
// Why: the 8-lanes-per-read kernel (extend_group.cuh) spends ~140 lane-instructions per cell of the reference's adaptive
// band, because every lane walks its fixed columns whether they are inside the band or not and every row pays a scan.
// Here a lane owns a whole read and executes exactly the reference's recurrence, one band cell per loop iteration
// (~25 instructions), and the loop nest (reads > chains > seeds > sides > band retries > rows > columns) is turned inside
// out into ONE loop whose body is "advance my own cell": lanes of a warp sit in different rows, problems and reads
// at the same time and still issue the same cell instructions together.  Everything that is not a cell (row change,
// problem set-up, the control flow of mem_chain2aln, claiming the next read) is the COLD step; lanes waiting for it are
// batched (the kernel runs the cold code when four lanes want it or nobody has cells left), so its ~60 instructions are
// paid once per several row ends, not once per row end.
//
// Column state: ksw_extend2's eh[j] = (h, e) plus the query base of column j, in one 32-bit word
//     bits 0..2  query code (0..4)     bits 4..16  h (13 bits)     bits 17..29  e (13 bits)
// kept in shared memory, word j of lane l at [j * 32 + l] (bank = lane: conflict-free whatever j each lane is at).
// 13 bits bound the scores to 8191, i.e. reads of at most 8191 / max(a, max(mat)) bases: lane_extend_eligible();
// longer reads take the group kernel.  Stale columns outside the band keep their value exactly like the reference's
// array.  The score matrix row of the current target base sits in two registers and a PRMT picks the entry with the
// column word itself as selector.
//
// The machine is HD: tests/hostsim runs it with stride 1 on the CPU against the golden regions.
#pragma once
#include "pipeline.cuh"
#if !defined(__CUDACC__)
struct uint4 { unsigned int x, y, z, w; };
#endif

namespace b200 {

#define LANE_H_MAX 8191
#define LANE_MAXLEN 512

HD int lane_maxmat(const Opt &opt)
{
    int mx = 0;                                     // ksw_extend2 starts its maximum at 0 (bwa/ksw.c:436)
    for (int i = 0; i < 25; ++i) mx = mx > opt.mat[i] ? mx : opt.mat[i];
    return mx;
}

HD bool lane_extend_eligible(const Opt &opt, int maxlen)
{
    int mx = lane_maxmat(opt);
    if (mx < opt.a) mx = opt.a;
    if (mx < 1) mx = 1;
    return maxlen <= LANE_MAXLEN && (i64)maxlen * mx <= LANE_H_MAX && opt.e_del > 0 && opt.e_ins > 0;
}

// rows[t] = mat[t*5 + 0..3] as four bytes, rows[4 + t] = mat[t*5 + 4] in byte 0
HD void lane_fill_rows(const Opt &opt, u32 *rows)
{
    for (int t = 0; t < 4; ++t) {
        u32 lo = 0;
        for (int q = 0; q < 4; ++q) lo |= (u32)(u8)opt.mat[t * 5 + q] << (8 * q);
        rows[t] = lo; rows[4 + t] = (u32)(u8)opt.mat[t * 5 + 4];
    }
}

HD size_t extend_lane_scratch_bytes(const Caps &c) { return sizeof(u64) * (size_t)c.seeds + sizeof(Reg) * (size_t)c.regs + 64; }

enum { LS_CELL = 0, LS_INIT = 1, LS_ROW = 2, LS_CTL = 3, LS_DEAD = 4 };   // LS_ROW: row change pending, LS_CTL: control step pending
enum { CK_FIRSTROW = 1, CK_ROWEND = 2 };
#define LANE_U 4            // band cells per hot step
#define LANE_INIT_U 8       // columns per initialisation step
enum { LP_READ = 0, LP_CHAIN, LP_SEED, LP_SIDE, LP_RESULT };

struct DpIn { int ql, qpos, qstep, tl, tstep, w, pen, h0; i64 tpos; };    // one ksw_extend2 call (qpos: index of query base 0 in the read)

// the lane's read as 4-bit codes, 8 per word, in shared memory (word k of lane l at [k * 32 + l]): the column
// initialisation of every DP problem takes its query codes from here, not from HBM
template <int S>
struct LaneQ {
    u32 qbase;                  // device: shared-window address of this lane's word 0
    u32 *qbuf;                  // host
    HD u32 ld(int k) const
    {
#if defined(__CUDA_ARCH__)
        u32 v;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(qbase + (u32)k * (S * 4u)));
        return v;
#else
        return qbuf[k * S];
#endif
    }
    HD void st(int k, u32 v) const
    {
#if defined(__CUDA_ARCH__)
        asm volatile("st.shared.u32 [%0], %1;" :: "r"(qbase + (u32)k * (S * 4u)), "r"(v));
#else
        qbuf[k * S] = v;
#endif
    }
    HD u32 code(int pos) const { return (ld(pos >> 3) >> (4 * (pos & 7))) & 7u; }
};
struct DpOut { int max, max_i, max_j, max_ie, gscore, moff; };                     // what it returns

// ------------------------------------------------------------------ DP: ksw_extend2 turned into steps.
// Scalars only and no address taken, so the whole struct lives in registers on the device.
template <int S>
struct LaneDP {
    u32 cbase;                  // device: shared-window address of this lane's column 0
    u32 *cols;                  // host: the column array
    LaneQ<S> q;
    const u32 *rows; const u64 *text;
    int maxmat, oe_del, oe_ins, e_del, e_ins, o_del, o_ins, zdrop;
    int st, ck;
    int qlen, tlen, w, h0, qpos, qstep, tstep; i64 tpos;
    int i, j, beg, end, h1, f, m, mj, max, max_i, max_j, max_ie, gscore, moff, hinit;
    u32 wq[LANE_U], rlo, rhi; u64 tw, tw_next; i64 twi, tlast;
    unsigned long long cells;

    HD u64 text_word(i64 wi) const                  // clamped: the look-ahead may step one word outside the text
    {
        wi = wi < 0 ? 0 : (wi > tlast ? tlast : wi);
#if defined(__CUDA_ARCH__)
        return __ldg(text + wi);
#else
        return text[wi];
#endif
    }

    HD u32 ld(int c) const
    {
#if defined(__CUDA_ARCH__)
        u32 v;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(cbase + (u32)c * (S * 4u)));
        return v;
#else
        return cols[c * S];
#endif
    }
    HD void st_(int c, u32 v) const
    {
#if defined(__CUDA_ARCH__)
        asm volatile("st.shared.u32 [%0], %1;" :: "r"(cbase + (u32)c * (S * 4u)), "r"(v));
#else
        cols[c * S] = v;
#endif
    }

    HD void setup(const Opt &opt, const u64 *text_, i64 text_bases, const u32 *rows_)
    {
        rows = rows_; text = text_; tlast = (text_bases - 1) >> 5;
        maxmat = lane_maxmat(opt);
        oe_del = opt.o_del + opt.e_del; oe_ins = opt.o_ins + opt.e_ins; e_del = opt.e_del; e_ins = opt.e_ins; o_del = opt.o_del; o_ins = opt.o_ins;
        zdrop = opt.zdrop;
        st = LS_CTL; ck = CK_ROWEND;
        twi = -1; tw = tw_next = 0; cells = 0;
        max = 0; max_i = max_j = max_ie = -1; gscore = -1; moff = 0;
    }

    HD void start(const DpIn &in)
    {
        qlen = in.ql; qpos = in.qpos; qstep = in.qstep; tlen = in.tl; tpos = in.tpos; tstep = in.tstep; h0 = in.h0;
        int max_ins = (int)((double)(qlen * maxmat + in.pen - o_ins) / e_ins + 1.);
        max_ins = max_ins > 1 ? max_ins : 1;
        w = in.w < max_ins ? in.w : max_ins;
        int max_del = (int)((double)(qlen * maxmat + in.pen - o_del) / e_del + 1.);
        max_del = max_del > 1 ? max_del : 1;
        w = w < max_del ? w : max_del;
        max = h0; max_i = max_j = -1; max_ie = -1; gscore = -1; moff = 0;
        beg = 0; end = qlen; i = 0;
        j = 0; hinit = h0; st = LS_INIT;
        // the text word of row 0 and the one the walk enters next (loaded 32 rows before it is needed)
        twi = tpos >> 5;
        if (tlen > 0) { tw = text_word(twi); tw_next = text_word(twi + tstep); }
    }

    HD DpOut out() const { DpOut r; r.max = max; r.max_i = max_i; r.max_j = max_j; r.max_ie = max_ie; r.gscore = gscore; r.moff = moff; return r; }

    // LANE_INIT_U columns of the eh[] initialisation (bwa/ksw.c:431-434) together with their query codes
    HD void init_step()
    {
#pragma unroll
        for (int u = 0; u < LANE_INIT_U; ++u) {
            if (j <= qlen) {
                u32 qb = j < qlen ? q.code(qpos + j * qstep) : 0u;
                st_(j, qb | (u32)hinit << 4);
                if (j == 0) hinit = h0 > oe_ins ? h0 - oe_ins : 0;
                else hinit = hinit > e_ins ? hinit - e_ins : 0;
                ++j;
            }
        }
        if (j > qlen) { st = LS_ROW; ck = CK_FIRSTROW; }
    }

    HD void row_begin()
    {
        if (beg < i - w) beg = i - w;
        if (end > i + w + 1) end = i + w + 1;
        if (end > qlen) end = qlen;
        if (beg == 0) { h1 = h0 - (o_del + e_del * (i + 1)); if (h1 < 0) h1 = 0; }
        else h1 = 0;
        f = 0; m = 0; mj = -1;
        i64 p = tpos + (i64)i * tstep;
        if ((p >> 5) != twi) {                          // one word further along the walk
            twi = p >> 5;
            tw = tw_next;
            tw_next = text_word(twi + tstep);
        }
        int tb = (int)((tw >> (2 * (p & 31))) & 3);
        rlo = rows[tb]; rhi = rows[4 + tb];
        cells += (unsigned long long)(end > beg ? end - beg : 0);
        j = beg;
        if (beg < end) {
#pragma unroll
            for (int u = 0; u < LANE_U; ++u) wq[u] = ld(beg + u);       // up to column end + LANE_U - 1 <= qlen + LANE_U - 1: allocated
            st = LS_CELL;
        } else { st = LS_ROW; ck = CK_ROWEND; }
    }

    // up to LANE_U cells of the band (bwa/ksw.c:468-490).  Only f, h1 and (m, mj) are carried from cell to cell, so the
    // LANE_U recurrences overlap in the pipeline; the words of the next LANE_U columns are loaded before they are needed.
    HD void cell_step()
    {
        u32 wv[LANE_U];
#pragma unroll
        for (int u = 0; u < LANE_U; ++u) wv[u] = wq[u];
        const int j0 = j;
#pragma unroll
        for (int u = 0; u < LANE_U; ++u) wq[u] = ld(j0 + LANE_U + u);
#pragma unroll
        for (int u = 0; u < LANE_U; ++u) {
            if (j0 + u < end) {
                int M = (int)((wv[u] >> 4) & 0x1fffu), e = (int)(wv[u] >> 17);
#if defined(__CUDA_ARCH__)
                int sc = (int)(i8)__byte_perm(rlo, rhi, wv[u]);   // selector nibble 0 = query code (bit 3 of the word is always 0)
#else
                int sc = (int)(i8)(((u64)rhi << 32 | rlo) >> (8 * (wv[u] & 7u)));
#endif
                M = M ? M + sc : 0;
                int h = M > e ? M : e;
                h = h > f ? h : f;
                mj = m > h ? mj : j0 + u;
                m = m > h ? m : h;
                int t = M - oe_del; t = t > 0 ? t : 0;
                e -= e_del; e = e > t ? e : t;
                t = M - oe_ins; t = t > 0 ? t : 0;
                f -= e_ins; f = f > t ? f : t;
                st_(j0 + u, (wv[u] & 7u) | (u32)h1 << 4 | (u32)e << 17);
                h1 = h;
            }
        }
        j = j0 + LANE_U < end ? j0 + LANE_U : end;
        if (j >= end) { st = LS_ROW; ck = CK_ROWEND; }
    }

    // end of row i (bwa/ksw.c:491-513) and the start of the next one; true when the DP is over
    HD bool row_end()
    {
        st_(end, (ld(end) & 7u) | (u32)h1 << 4);                       // eh[end].h = h1, eh[end].e = 0
        if (j == qlen) {
            max_ie = gscore > h1 ? max_ie : i;
            gscore = gscore > h1 ? gscore : h1;
        }
        if (m == 0) return true;
        if (m > max) {
            max = m; max_i = i; max_j = mj;
            int d = mj - i; d = d < 0 ? -d : d;
            moff = moff > d ? moff : d;
        } else if (zdrop > 0) {
            if (i - max_i > mj - max_j) {
                if (max - m - ((i - max_i) - (mj - max_j)) * e_del > zdrop) return true;
            } else {
                if (max - m - ((mj - max_j) - (i - max_i)) * e_ins > zdrop) return true;
            }
        }
        int jj;
        for (jj = beg; jj < end && (ld(jj) >> 4) == 0; ++jj) {}
        beg = jj;
        for (jj = end; jj >= beg && (ld(jj) >> 4) == 0; --jj) {}
        end = jj + 2 < qlen ? jj + 2 : qlen;
        if (++i >= tlen) return true;
        row_begin();
        return false;
    }

    // the row step (state LS_ROW): end of a row and/or start of the next one.  Leaves the lane in LS_CELL, LS_ROW (an empty
    // row) or LS_CTL (the DP is over).
    HD void row_step()
    {
        bool done;
        if (ck == CK_ROWEND) done = row_end();
        else if (tlen > 0) { row_begin(); done = false; }
        else done = true;
        if (done) st = LS_CTL;
    }
};

// ------------------------------------------------------------------ control: mem_chain2aln as a resumable machine
// what every lane shares (kernel parameters)
struct LaneEnv {
    const DevIndex *ix; const Opt *opt; const Caps *caps; const Batch *B;
    unsigned long long *work_ctr; i64 n_work; const i32 *order;
    const u32 *packed4; int qw4;    // the reads as 4-bit codes, qw4 words (a multiple of 4) per read; NULL: pack from the bytes
};
#define LANE_CLAIM 4        // reads a lane claims per atomic

// per-lane control state.  On the device it lives in shared memory (lane l at l * LANE_CTL_WORDS words, an odd stride):
// with the whole carve-out given to shared memory the L1 is too small for local memory, and this state is what the
// latency-bound control steps read and write.
struct LaneCtl {
    u8 *scratch;
    unsigned long long ref_bytes, n_ext;
    i64 wk, wk_end;                 // claimed work items [wk, wk_end)
    // read
    i64 rid; int len; const u8 *query; u64 *srt; RegSink av; const Chain *oc; const Seed *os; int n_chains, ci; float frac;
    // chain
    const Seed *cs; int cn, c_rid, k; i64 rmax0, rmax1;
    // seed / region under construction
    Seed s; Reg *a; int pending, side, tries, aw0, aw1, sc0, prev, phase;

    HD void setup(const LaneEnv &E, u8 *scratch_)
    {
        scratch = scratch_;
        ref_bytes = n_ext = 0;
        wk = wk_end = 0;
        srt = (u64 *)scratch;
        av.a = (Reg *)(scratch + sizeof(u64) * (size_t)E.caps->seeds); av.n = 0; av.cap = E.caps->regs; av.overflow = false;
        phase = LP_READ;
    }

    HD void start_side(const Opt *opt, DpIn &in) const
    {
        const int qe = s.qbeg + s.len;
        const i64 re = s.rbeg + s.len - rmax0;
        if (side == 0) {
            in.ql = s.qbeg; in.qpos = s.qbeg - 1; in.qstep = -1; in.tl = (int)(s.rbeg - rmax0); in.tpos = s.rbeg - 1; in.tstep = -1;
            in.w = aw0; in.pen = opt->pen_clip5; in.h0 = s.len * opt->a;
        } else {
            in.ql = len - qe; in.qpos = qe; in.qstep = 1; in.tl = (int)(rmax1 - rmax0 - re); in.tpos = rmax0 + re; in.tstep = 1;
            in.w = aw1; in.pen = opt->pen_clip3; in.h0 = sc0;
        }
    }

    template <class Q>
    HD bool claim_read(const LaneEnv &E, const Q &q)
    {
        const Batch *B = E.B;
        for (;;) {
            if (wk >= wk_end) {
#if defined(__CUDA_ARCH__)
                wk = (i64)atomicAdd(E.work_ctr, (unsigned long long)LANE_CLAIM);
#else
                wk = (i64)*E.work_ctr; *E.work_ctr += LANE_CLAIM;
#endif
                wk_end = wk + LANE_CLAIM < E.n_work ? wk + LANE_CLAIM : E.n_work;
                if (wk >= E.n_work) return false;
            }
            rid = E.order ? (i64)E.order[wk] : wk;
            ++wk;
            ReadRec &R = B->rec[rid];
            R.n_regs = 0; R.reg_off = 0;
            if (B->ovf[rid]) continue;
            len = (int)(B->seq_off[rid + 1] - B->seq_off[rid]);
            query = B->seq + B->seq_off[rid];
            // the read as 4-bit codes, into the lane's shared memory
            if (E.packed4) {
                const uint4 *src = reinterpret_cast<const uint4 *>(E.packed4 + (size_t)rid * E.qw4);
                for (int x = 0; x * 8 < len; x += 4) {
                    uint4 v = src[x >> 2];
                    q.st(x, v.x); q.st(x + 1, v.y); q.st(x + 2, v.z); q.st(x + 3, v.w);
                }
            } else {
                for (int x = 0; x < len; x += 8) {
                    u32 v = 0;
                    for (int y = 0; y < 8; ++y) if (x + y < len) v |= (u32)query[x + y] << (4 * y);
                    q.st(x >> 3, v);
                }
            }
            av.n = 0; av.overflow = false;
            oc = B->pool.chains + R.chain_off;
            os = B->pool.seeds + R.seed_off;
            n_chains = R.n_chains; frac = R.frac_rep;
            ci = 0;
            return true;
        }
    }

    HD void finish_read(const LaneEnv &E)
    {
        const Batch *B = E.B;
        ReadRec &R = B->rec[rid];
        i64 off = pool_alloc(B->pool, POOL_REG, av.n);
        if (off < 0) { B->ovf[rid] |= OVF_POOL; return; }
        const u32 *src = (const u32 *)av.a; u32 *dst = (u32 *)(B->pool.regs + off);
        int words = av.n * (int)(sizeof(Reg) / 4);
        for (int x = 0; x < words; ++x) dst[x] = src[x];
        R.n_regs = av.n; R.reg_off = off;
    }

    // r: the result of the DP that just ended (ignored unless phase == LP_RESULT).
    // false: no reads left; true: `in` holds the next DP problem
    template <class Q>
    HD bool advance(const LaneEnv &E, const Q &q, const DpOut &r, DpIn &in)
    {
        const Opt &o = *E.opt;
        const Opt *opt = E.opt;
        const DevIndex *ix = E.ix;
        const Batch *B = E.B;
        const i64 l_pac = ix->l_pac;
        for (;;) {
            if (phase == LP_RESULT) {
                const int aw = side ? aw1 : aw0;
                const int pen = side ? o.pen_clip3 : o.pen_clip5;
                n_ext++;
                a->score = r.max;
                if (!(a->score == prev || r.moff < (aw >> 1) + (aw >> 2)) && tries + 1 < B200_MAX_BAND_TRY) {
                    ++tries; prev = a->score;
                    if (side) aw1 = o.w << tries; else aw0 = o.w << tries;
                    start_side(opt, in);
                    return true;
                }
                const int qle = r.max_j + 1, tle = r.max_i + 1, gtle = r.max_ie + 1;
                const bool local = r.gscore <= 0 || r.gscore <= a->score - pen;
                if (side == 0) {
                    if (local) { a->qb = s.qbeg - qle; a->rb = s.rbeg - tle; a->truesc = a->score; }
                    else { a->qb = 0; a->rb = s.rbeg - gtle; a->truesc = r.gscore; }
                } else {
                    const int qe = s.qbeg + s.len;
                    const i64 re = s.rbeg + s.len - rmax0;
                    if (local) { a->qe = qe + qle; a->re = rmax0 + re + tle; a->truesc += a->score - sc0; }
                    else { a->qe = len; a->re = rmax0 + re + gtle; a->truesc += r.gscore - sc0; }
                }
                phase = LP_SIDE;
            } else if (phase == LP_SIDE) {
                if (pending) {
                    side = (pending & 1) ? 0 : 1;
                    pending &= ~(1 << side);
                    tries = 0; prev = a->score; sc0 = a->score;
                    if (side) aw1 = o.w; else aw0 = o.w;
                    start_side(opt, in);
                    phase = LP_RESULT;
                    return true;
                }
                int cov = 0;
                for (int x = 0; x < cn; ++x) {
                    const Seed &t = cs[x];
                    if (t.qbeg >= a->qb && t.qbeg + t.len <= a->qe && t.rbeg >= a->rb && t.rbeg + t.len <= a->re) cov += t.len;
                }
                a->seedcov = cov;
                a->w = aw0 > aw1 ? aw0 : aw1;
                a->seedlen0 = s.len;
                a->frac_rep = frac;
                ++av.n;
                --k;
                phase = LP_SEED;
            } else if (phase == LP_SEED) {
                if (k < 0) { ++ci; phase = LP_CHAIN; continue; }
                s = cs[(u32)srt[k]];
                int x;
                for (x = 0; x < av.n; ++x) {
                    const Reg *p = &av.a[x];
                    i64 rd; int qd, w2, max_gap;
                    if (s.rbeg < p->rb || s.rbeg + s.len > p->re || s.qbeg < p->qb || s.qbeg + s.len > p->qe) continue;
                    if (s.len - p->seedlen0 > .1 * len) continue;
                    qd = s.qbeg - p->qb; rd = s.rbeg - p->rb;
                    max_gap = cal_max_gap(o, qd < rd ? qd : (int)rd);
                    w2 = max_gap < p->w ? max_gap : p->w;
                    if (qd - rd < w2 && rd - qd < w2) break;
                    qd = p->qe - (s.qbeg + s.len); rd = p->re - (s.rbeg + s.len);
                    max_gap = cal_max_gap(o, qd < rd ? qd : (int)rd);
                    w2 = max_gap < p->w ? max_gap : p->w;
                    if (qd - rd < w2 && rd - qd < w2) break;
                }
                if (x < av.n) {
                    for (x = k + 1; x < cn; ++x) {
                        if (srt[x] == 0) continue;
                        const Seed *t = &cs[(u32)srt[x]];
                        if (t->len < s.len * .95) continue;
                        if (s.qbeg <= t->qbeg && s.qbeg + s.len - t->qbeg >= s.len >> 2 && t->qbeg - s.qbeg != t->rbeg - s.rbeg) break;
                        if (t->qbeg <= s.qbeg && t->qbeg + t->len - s.qbeg >= s.len >> 2 && s.qbeg - t->qbeg != s.rbeg - t->rbeg) break;
                    }
                    if (x == cn) { srt[k] = 0; --k; continue; }
                }
                if (av.n >= av.cap) {
                    B->ovf[rid] |= OVF_REG;
                    phase = LP_READ;
                    continue;
                }
                a = &av.a[av.n];
                a->rb = a->re = 0; a->qb = a->qe = a->rid = a->score = a->truesc = a->sub = a->alt_sc = a->csub = a->sub_n = 0;
                a->w = a->seedcov = a->secondary = a->secondary_all = a->seedlen0 = a->n_comp = a->is_alt = 0;
                a->frac_rep = 0; a->pad_ = 0; a->hash = 0;
                a->w = aw0 = aw1 = o.w;
                a->score = a->truesc = -1;
                a->rid = c_rid;
                const int qe = s.qbeg + s.len;
                if (!s.qbeg) { a->score = a->truesc = s.len * o.a; a->qb = 0; a->rb = s.rbeg; }
                if (qe == len) { a->qe = len; a->re = s.rbeg + s.len; }
                pending = (s.qbeg ? 1 : 0) | (qe != len ? 2 : 0);
                phase = LP_SIDE;
            } else if (phase == LP_CHAIN) {
                if (ci >= n_chains) { finish_read(E); phase = LP_READ; continue; }
                cs = os + oc[ci].head; cn = oc[ci].n; c_rid = oc[ci].rid;
                if (cn == 0) { ++ci; continue; }
                i64 r0 = l_pac << 1, r1 = 0;
                for (int x = 0; x < cn; ++x) {
                    const Seed &t = cs[x];
                    i64 b = t.rbeg - (t.qbeg + cal_max_gap(o, t.qbeg));
                    i64 e = t.rbeg + t.len + ((len - t.qbeg - t.len) + cal_max_gap(o, len - t.qbeg - t.len));
                    r0 = r0 < b ? r0 : b;
                    r1 = r1 > e ? r1 : e;
                }
                r0 = r0 > 0 ? r0 : 0;
                r1 = r1 < l_pac << 1 ? r1 : l_pac << 1;
                if (r0 < l_pac && l_pac < r1) {
                    if (cs[0].rbeg < l_pac) r1 = l_pac;
                    else r0 = l_pac;
                }
                {   // bns_fetch_seq (bwa/bntseq.c:426-451): clip the window to the contig of the first seed
                    int is_rev;
                    int rr = pos2rid(*ix, depos(*ix, cs[0].rbeg, &is_rev));
                    i64 far_beg = ix->contig_off[rr], far_end = ix->contig_off[rr + 1];
                    if (is_rev) { i64 t2 = far_beg; far_beg = (l_pac << 1) - far_end; far_end = (l_pac << 1) - t2; }
                    r0 = r0 > far_beg ? r0 : far_beg;
                    r1 = r1 < far_end ? r1 : far_end;
                    ref_bytes += (unsigned long long)((r1 - r0 + 3) >> 2);
                }
                rmax0 = r0; rmax1 = r1;
                for (int x = 0; x < cn; ++x) srt[x] = (u64)cs[x].score << 32 | (u64)x;
                introsort((size_t)cn, srt, U64Less());
                k = cn - 1;
                phase = LP_SEED;
            } else {                                        // LP_READ
                if (!claim_read(E, q)) return false;
                phase = LP_CHAIN;
            }
        }
    }
};

// host driver (tests/hostsim): runs the two machines over one read
#if !defined(__CUDA_ARCH__)
inline void stage_extend_lane_host(const DevIndex &ix, const Opt &opt, const Caps &caps, const Batch &B, i64 rid, u8 *scratch, u32 *cols, u32 *qbuf, CtrLocal &ctr)
{
    u32 rows[8];
    lane_fill_rows(opt, rows);
    unsigned long long next = 0;
    i32 ord = (i32)rid;
    LaneEnv E; E.ix = &ix; E.opt = &opt; E.caps = &caps; E.B = &B; E.work_ctr = &next; E.n_work = 1; E.order = &ord; E.packed4 = nullptr; E.qw4 = 0;
    LaneCtl c;
    c.setup(E, scratch);
    LaneDP<1> d;
    d.cols = cols; d.cbase = 0;
    d.q.qbuf = qbuf; d.q.qbase = 0;
    d.setup(opt, ix.text, (i64)ix.seq_len, rows);
    for (;;) {
        if (d.st == LS_CELL) d.cell_step();
        else if (d.st == LS_INIT) d.init_step();
        else if (d.st == LS_ROW) d.row_step();
        else {
            DpIn in;
            if (!c.advance(E, d.q, d.out(), in)) break;
            d.start(in);
        }
    }
    ctr.sw_cells += d.cells; ctr.n_ext += c.n_ext; ctr.ref_bytes += c.ref_bytes;
}
#endif

} // namespace b200
