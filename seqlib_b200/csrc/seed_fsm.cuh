// seed_fsm.cuh -- mem_collect_intv (bwa/bwamem.c:140-188) as a resumable state machine.
//
// seed.cuh states the algorithm as the reference does: nested loops with a bwt_extend call in three places.  On a
// GPU that shape is expensive: lanes of a warp sit in different loops, so the one costly operation -- two dependent
// 32-byte Occ gathers plus the popcounts -- is issued once per call site with a fraction of the lanes.  Here the same
// loops are written as a coroutine (switch-resumed, all live variables in SeedFsm) that *yields* every extension
// request; the kernel owns a single extend_c call that all lanes of the warp reach together, and a lane whose read is
// finished immediately claims the next read, so gathers stay 32 wide and in flight together.
// Results are identical to collect_intv() (checked on the CPU by tests/hostsim, which can run either form).
#pragma once
#include "seed.cuh"

namespace b200 {

struct SeedFsm {
    int state;
    int pass, x, k, old_n, from, split_len;                     // mem_collect_intv
    int sx, min_intv, i, j, c, ncurr, nprev, base, ret;         // bwt_smem1a
    Intv ik;
    Intv *prev, *curr;
    Intv req, res;                                              // pending extension: res = extend_c(req, req_c, req_back)
    int req_c, req_back;
};

enum { S_START = 0, S_TASK, S_FWD, S_FWD_RES, S_FWD_END, S_BWD_ROW, S_BWD_ITEM, S_BWD_RES, S_BWD_END, S_P3_NEXT, S_P3_FWD, S_P3_RES, S_SORT, S_DONE };

HD void seed_fsm_init(SeedFsm &f, Intv *prev, Intv *curr) { f.state = S_START; f.prev = prev; f.curr = curr; }

#define FSM_REQ(IN, C, BACK, NEXT) { f.req = (IN); f.req_c = (C); f.req_back = (BACK); f.state = (NEXT); return true; }

// Runs until the next extension is needed (returns true, request in f.req/req_c/req_back; the caller stores the
// child interval in f.res and calls again) or the read is done (returns false; out holds the sorted intervals,
// or out.overflow is set).  A plain loop around a switch: every state is a short straight-line piece of the
// reference's loops, named after the place in bwt_smem1a / bwt_seed_strategy1 it comes from.
HD bool seed_step(const DevIndex &ix, const Opt &opt, int len, const u8 *seq, IntvSink &out, SeedFsm &f)
{
    for (;;) {
        switch (f.state) {
        case S_START:
            f.split_len = (int)(opt.min_seed_len * opt.split_factor + .499);
            out.n = 0;
            f.pass = 1; f.x = 0; f.k = 0; f.old_n = 0;
            f.state = S_TASK;
            break;
        case S_TASK: {      // next SMEM task: pass 1 walks x over the read, pass 2 re-seeds inside long, rare SMEMs
            bool have = false;
            if (f.pass == 1) {
                while (f.x < len && seq[f.x] >= 4) ++f.x;
                if (f.x >= len) { f.pass = 2; f.old_n = out.n; f.k = 0; }
                else { f.sx = f.x; f.min_intv = 1; have = true; }
            }
            if (f.pass == 2) {
                for (; f.k < f.old_n; ++f.k) {
                    const Intv p = out.a[f.k];
                    int start = (int)(p.info >> 32), end = (int)(i32)p.info;
                    if (end - start < f.split_len || p.x2 > (u64)opt.split_width) continue;
                    f.sx = (start + end) >> 1; f.min_intv = (int)p.x2 + 1;
                    have = true; ++f.k;
                    break;
                }
            }
            if (!have) { f.x = 0; f.state = opt.max_mem_intv > 0 ? S_P3_NEXT : S_SORT; break; }
            f.from = out.n;
            f.ret = f.sx + 1;
            if (seq[f.sx] > 3) { f.base = out.n; f.state = S_BWD_END; break; }       // bwt_smem1a returns x + 1 at once
            if (f.min_intv < 1) f.min_intv = 1;
            set_intv(ix, seq[f.sx], f.ik);
            f.ik.info = f.sx + 1;
            f.ncurr = 0;
            f.i = f.sx + 1;
            f.state = S_FWD;
            break;
        }
        case S_FWD:         // head of the forward loop (bwa/bwt.c:303-319)
            if (f.i >= len) { f.curr[f.ncurr++] = f.ik; f.state = S_FWD_END; break; }
            if (seq[f.i] < 4) { f.c = 3 - seq[f.i]; FSM_REQ(f.ik, f.c, 0, S_FWD_RES) }
            f.curr[f.ncurr++] = f.ik;
            f.state = S_FWD_END;
            break;
        case S_FWD_RES:
            if (f.res.x2 != f.ik.x2) {
                f.curr[f.ncurr++] = f.ik;
                if (f.res.x2 < (u64)f.min_intv) { f.state = S_FWD_END; break; }
            }
            f.ik = f.res; f.ik.info = f.i + 1;
            ++f.i;
            f.state = S_FWD;
            break;
        case S_FWD_END:
            reverse_(f.curr, f.ncurr);
            f.ret = (int)f.curr[0].info;
            { Intv *t = f.curr; f.curr = f.prev; f.prev = t; }
            f.nprev = f.ncurr;
            f.base = out.n;
            f.i = f.sx - 1;
            f.state = S_BWD_ROW;
            break;
        case S_BWD_ROW:     // one position i of the backward loop (bwa/bwt.c:325-345)
            if (f.i < -1) { f.state = S_BWD_END; break; }
            f.c = f.i < 0 ? -1 : seq[f.i] < 4 ? seq[f.i] : -1;
            f.ncurr = 0; f.j = 0;
            f.state = S_BWD_ITEM;
            break;
        case S_BWD_ITEM:
            if (f.j >= f.nprev) {
                if (f.ncurr == 0) { f.state = S_BWD_END; break; }
                { Intv *t = f.curr; f.curr = f.prev; f.prev = t; }
                f.nprev = f.ncurr;
                --f.i;
                f.state = S_BWD_ROW;
                break;
            }
            if (f.c >= 0) FSM_REQ(f.prev[f.j], f.c, 1, S_BWD_RES)
            f.state = S_BWD_RES;
            break;
        case S_BWD_RES:
            if (f.c < 0 || f.res.x2 < (u64)f.min_intv) {
                if (f.ncurr == 0) {
                    if (out.n == f.base || (u64)(f.i + 1) < (out.a[out.n - 1].info >> 32)) {
                        Intv t = f.prev[f.j]; t.info |= (u64)(f.i + 1) << 32;
                        out.push(t);
                        if (out.overflow) { f.state = S_DONE; return false; }
                    }
                }
            } else if (f.ncurr == 0 || f.res.x2 != f.curr[f.ncurr - 1].x2) {
                Intv t = f.res; t.info = f.prev[f.j].info;
                f.curr[f.ncurr++] = t;
            }
            ++f.j;
            f.state = S_BWD_ITEM;
            break;
        case S_BWD_END:
            reverse_(out.a + f.base, out.n - f.base);
            keep_long_(out, f.from, opt.min_seed_len);
            if (f.pass == 1) f.x = f.ret;
            f.state = S_TASK;
            break;
        case S_P3_NEXT:     // bwt_seed_strategy1 (bwa/bwt.c:358-379) restarted where the previous call stopped
            while (f.x < len && seq[f.x] >= 4) ++f.x;
            if (f.x >= len) { f.state = S_SORT; break; }
            set_intv(ix, seq[f.x], f.ik);
            f.i = f.x + 1;
            f.state = S_P3_FWD;
            break;
        case S_P3_FWD:
            if (f.i >= len) { f.x = len; f.state = S_SORT; break; }
            if (seq[f.i] < 4) { f.c = 3 - seq[f.i]; FSM_REQ(f.ik, f.c, 0, S_P3_RES) }
            f.x = f.i + 1;
            f.state = S_P3_NEXT;
            break;
        case S_P3_RES:
            if (f.res.x2 < (u64)(int)opt.max_mem_intv && f.i - f.x >= opt.min_seed_len) {
                if (f.res.x2 > 0) {
                    Intv m = f.res;
                    m.info = (u64)f.x << 32 | (u64)(f.i + 1);
                    out.push(m);
                    if (out.overflow) { f.state = S_DONE; return false; }
                }
                f.x = f.i + 1;
                f.state = S_P3_NEXT;
                break;
            }
            f.ik = f.res;
            ++f.i;
            f.state = S_P3_FWD;
            break;
        case S_SORT:
            introsort((size_t)out.n, out.a, IntvLess());
            f.state = S_DONE;
            return false;
        default:
            return false;
        }
    }
}

#undef FSM_REQ

// Host / scalar driver: same results as collect_intv().
template <class Ctr>
HD void collect_intv_fsm(const DevIndex &ix, const Opt &opt, int len, const u8 *seq, IntvSink &out, Intv *prev, Intv *curr, Ctr &ctr)
{
    SeedFsm f;
    seed_fsm_init(f, prev, curr);
    while (seed_step(ix, opt, len, seq, out, f)) extend_c(ix, f.req, f.req_c, f.req_back, f.res, ctr);
}

} // namespace b200
