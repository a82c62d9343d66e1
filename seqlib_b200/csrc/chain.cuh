// chain.cuh -- seed chaining and chain filtering for one read.
//   build_chains <- mem_chain      (bwa/bwamem.c:277-341) incl. test_and_merge (:216-237)
//   chain_weight <- mem_chain_weight (bwa/bwamem.c:239-258)
//   filter_chains<- mem_chain_flt  (bwa/bwamem.c:353-411)
// The reference keeps chains in a klib B-tree keyed by `pos` (kbtree.h, node
// size 512 => t = 5, at most 9 keys per node).  With duplicate keys the chain
// returned by kb_intervalp and the in-order traversal depend on the tree
// shape, so the same B-tree (pre-emptive split on the way down, lower-bound
// search inside a node) is kept here over chain indices.
#pragma once
#include "common.cuh"
#include "fmindex.cuh"
#include "sort.cuh"

namespace b200 {

enum { BT_T = 5, BT_MAXK = 2 * BT_T - 1 };

struct BtNode {
    i32 n, internal;
    i32 key[BT_MAXK];       // chain indices
    i32 ptr[BT_MAXK + 1];   // node indices
};

struct ChainWork {          // per-read working set, all in HBM scratch
    Chain *chains; int n_chains, cap_chains;
    Seed *seeds;   int n_seeds, cap_seeds;
    BtNode *nodes; int n_nodes, cap_nodes;
    int root;
    int n_keys;
    u32 ovf;
};

HD int bt_new_node(ChainWork &w, int internal)
{
    if (w.n_nodes >= w.cap_nodes) { w.ovf |= OVF_CHAIN; return 0; }
    BtNode &z = w.nodes[w.n_nodes];
    z.n = 0; z.internal = internal;
    for (int i = 0; i <= BT_MAXK; ++i) z.ptr[i] = -1;
    return w.n_nodes++;
}

// __kb_getp_aux (bwa/kbtree.h:123-138): position of k inside node x
HD int bt_find(const ChainWork &w, const BtNode &x, i64 k, int *r)
{
    int begin = 0, end = x.n;
    if (x.n == 0) return -1;
    while (begin < end) {
        int mid = (begin + end) >> 1;
        if (w.chains[x.key[mid]].pos < k) begin = mid + 1; else end = mid;
    }
    if (begin == x.n) { *r = 1; return x.n - 1; }
    i64 kp = w.chains[x.key[begin]].pos;
    *r = (kp < k) - (k < kp);
    if (*r < 0) --begin;
    return begin;
}

// kb_intervalp (bwa/kbtree.h:159-178): only `lower` is used by mem_chain
HD int bt_lower(const ChainWork &w, i64 k)
{
    int lower = -1, r = 0, xi = w.root;
    while (xi >= 0) {
        const BtNode &x = w.nodes[xi];
        int i = bt_find(w, x, k, &r);
        if (i >= 0 && r == 0) return x.key[i];
        if (i >= 0) lower = x.key[i];
        if (!x.internal) return lower;
        xi = x.ptr[i + 1];
    }
    return lower;
}

// __kb_split (bwa/kbtree.h:187-204)
HD void bt_split(ChainWork &w, int xi, int i, int yi)
{
    int zi = bt_new_node(w, w.nodes[yi].internal);
    if (w.ovf) return;
    BtNode &x = w.nodes[xi], &y = w.nodes[yi], &z = w.nodes[zi];
    z.n = BT_T - 1;
    for (int j = 0; j < BT_T - 1; ++j) z.key[j] = y.key[BT_T + j];
    if (y.internal) for (int j = 0; j < BT_T; ++j) z.ptr[j] = y.ptr[BT_T + j];
    y.n = BT_T - 1;
    for (int j = x.n; j > i; --j) x.ptr[j + 1] = x.ptr[j];
    x.ptr[i + 1] = zi;
    for (int j = x.n - 1; j >= i; --j) x.key[j + 1] = x.key[j];
    x.key[i] = y.key[BT_T - 1];
    ++x.n;
}

// kb_putp (bwa/kbtree.h:205-243)
HD void bt_put(ChainWork &w, int chain_idx)
{
    i64 k = w.chains[chain_idx].pos;
    ++w.n_keys;
    int ri = w.root;
    if (w.nodes[ri].n == BT_MAXK) {
        int si = bt_new_node(w, 1);
        if (w.ovf) return;
        w.root = si;
        w.nodes[si].ptr[0] = ri;
        bt_split(w, si, 0, ri);
        if (w.ovf) return;
        ri = si;
    }
    int xi = ri, r;
    for (;;) {
        BtNode &x = w.nodes[xi];
        if (!x.internal) {
            int i = bt_find(w, x, k, &r);
            for (int j = x.n - 1; j > i; --j) x.key[j + 1] = x.key[j];
            x.key[i + 1] = chain_idx;
            ++x.n;
            return;
        }
        int i = bt_find(w, x, k, &r) + 1;
        if (w.nodes[x.ptr[i]].n == BT_MAXK) {
            bt_split(w, xi, i, x.ptr[i]);
            if (w.ovf) return;
            i64 kp = w.chains[w.nodes[xi].key[i]].pos;
            if (((kp < k) - (k < kp)) > 0) ++i;
        }
        xi = w.nodes[xi].ptr[i];
    }
}

// in-order traversal (__kb_traverse, bwa/kbtree.h:346-370) into order[]
HD int bt_traverse(const ChainWork &w, i32 *order)
{
    // explicit stack: depth <= log_5(n)+2
    int sx[24], si[24], sp = 0, n = 0;
    sx[0] = w.root; si[0] = 0;
    for (;;) {
        while (sx[sp] >= 0 && si[sp] <= w.nodes[sx[sp]].n) {
            const BtNode &x = w.nodes[sx[sp]];
            sx[sp + 1] = x.internal ? x.ptr[si[sp]] : -1;
            si[sp + 1] = 0;
            ++sp;
        }
        --sp;
        if (sp < 0) break;
        if (sx[sp] >= 0 && si[sp] < w.nodes[sx[sp]].n) order[n++] = w.nodes[sx[sp]].key[si[sp]];
        ++si[sp];
    }
    return n;
}

HD const Seed &chain_seed0(const ChainWork &w, const Chain &c) { return w.seeds[c.head]; }
HD const Seed &chain_seedL(const ChainWork &w, const Chain &c) { return w.seeds[c.tail]; }

// test_and_merge (bwa/bwamem.c:216-237)
HD int test_and_merge(const Opt &opt, i64 l_pac, ChainWork &w, Chain &c, const Seed &p, int seed_rid)
{
    const Seed &last = chain_seedL(w, c), &first = chain_seed0(w, c);
    i64 qend = last.qbeg + last.len, rend = last.rbeg + last.len;
    if (seed_rid != c.rid) return 0;
    if (p.qbeg >= first.qbeg && p.qbeg + p.len <= qend && p.rbeg >= first.rbeg && p.rbeg + p.len <= rend) return 1;
    if ((last.rbeg < l_pac || first.rbeg < l_pac) && p.rbeg >= l_pac) return 0;
    i64 x = p.qbeg - last.qbeg, y = p.rbeg - last.rbeg;
    if (y >= 0 && x - y <= opt.w && y - x <= opt.w && x - last.len < opt.max_chain_gap && y - last.len < opt.max_chain_gap) {
        if (w.n_seeds >= w.cap_seeds) { w.ovf |= OVF_SEED; return 1; }
        int si = w.n_seeds++;
        w.seeds[si] = p; w.seeds[si].next = -1;
        w.seeds[c.tail].next = si;
        c.tail = si; ++c.n;
        return 1;
    }
    return 0;
}

// Returns l_rep (for frac_rep) and fills w.chains / order[] (in-order chain indices).
template <class Ctr>
HD int build_chains(const DevIndex &ix, const Opt &opt, int len, const Intv *intv, int n_intv,
                    ChainWork &w, i32 *order, int *n_order, Ctr &ctr)
{
    int b = 0, e = 0, l_rep = 0;
    *n_order = 0;
    w.n_chains = w.n_seeds = w.n_nodes = w.n_keys = 0; w.ovf = 0;
    if (len < opt.min_seed_len) return 0;
    w.root = bt_new_node(w, 0);
    for (int i = 0; i < n_intv; ++i) {
        const Intv &p = intv[i];
        int sb = (int)(p.info >> 32), se = (int)(u32)p.info;
        if (p.x2 <= (u64)opt.max_occ) continue;
        if (sb > e) { l_rep += e - b; b = sb; e = se; }
        else e = e > se ? e : se;
    }
    l_rep += e - b;
    for (int i = 0; i < n_intv; ++i) {
        const Intv &p = intv[i];
        int slen = (int)((u32)p.info - (u32)(p.info >> 32));
        i64 step = p.x2 > (u64)opt.max_occ ? (i64)(p.x2 / opt.max_occ) : 1;
        int count = 0;
        for (i64 k = 0; k < (i64)p.x2 && count < opt.max_occ; k += step, ++count) {
            Seed s;
            // an interval the seeding machine followed through the text carries its position (seed2.cuh), else bwt_sa
            s.rbeg = (p.x0 >> 63) ? (i64)(p.x0 & ~(1ull << 63)) : (i64)sa_lookup(ix, p.x0 + k, ctr);
            s.qbeg = (i32)(p.info >> 32);
            s.score = s.len = slen;
            s.next = -1;
            int rid = intv2rid(ix, s.rbeg, s.rbeg + s.len);
            if (rid < 0) continue;
            bool to_add = false;
            if (w.n_keys) {
                int lower = bt_lower(w, s.rbeg);
                if (lower < 0 || !test_and_merge(opt, ix.l_pac, w, w.chains[lower], s, rid)) to_add = true;
            } else to_add = true;
            if (w.ovf) return l_rep;
            if (to_add) {
                if (w.n_chains >= w.cap_chains) { w.ovf |= OVF_CHAIN; return l_rep; }
                if (w.n_seeds >= w.cap_seeds) { w.ovf |= OVF_SEED; return l_rep; }
                int si = w.n_seeds++, ci = w.n_chains++;
                w.seeds[si] = s;
                Chain &c = w.chains[ci];
                c.pos = s.rbeg; c.n = 1; c.first = -1; c.rid = rid; c.w = 0; c.kept = 0;
                c.is_alt = ix.contig_alt[rid] ? 1 : 0;
                c.head = c.tail = si;
                bt_put(w, ci);
                if (w.ovf) return l_rep;
            }
        }
    }
    *n_order = bt_traverse(w, order);
    return l_rep;
}

// mem_chain_weight (bwa/bwamem.c:239-258)
HD int chain_weight(const ChainWork &w, const Chain &c)
{
    i64 end = 0;
    int wt = 0, tmp;
    for (int s = c.head; s >= 0; s = w.seeds[s].next) {
        const Seed &sd = w.seeds[s];
        if (sd.qbeg >= end) wt += sd.len;
        else if (sd.qbeg + sd.len > end) wt += (int)(sd.qbeg + sd.len - end);
        end = end > sd.qbeg + sd.len ? end : sd.qbeg + sd.len;
    }
    tmp = wt; wt = 0; end = 0;
    for (int s = c.head; s >= 0; s = w.seeds[s].next) {
        const Seed &sd = w.seeds[s];
        if (sd.rbeg >= end) wt += sd.len;
        else if (sd.rbeg + sd.len > end) wt += (int)(sd.rbeg + sd.len - end);
        end = end > sd.rbeg + sd.len ? end : sd.rbeg + sd.len;
    }
    wt = wt < tmp ? wt : tmp;
    return wt < (1 << 30) ? wt : (1 << 30) - 1;
}

struct ChainFltLess {
    const Chain *c;
    HD bool operator()(i32 a, i32 b) const { return c[a].w > c[b].w; }
};

// mem_chain_flt (bwa/bwamem.c:353-411) over the index array a[0..n); `kept_idx`
// is scratch for the list of non-overlapping chains.  Returns the new n.
HD int filter_chains(const Opt &opt, ChainWork &w, i32 *a, int n_chn, i32 *kept_idx)
{
    if (n_chn == 0) return 0;
    Chain *C = w.chains;
    int i, k;
    for (i = k = 0; i < n_chn; ++i) {
        Chain &c = C[a[i]];
        c.first = -1; c.kept = 0;
        c.w = chain_weight(w, c);
        if (c.w < opt.min_chain_weight) continue;
        a[k++] = a[i];
    }
    n_chn = k;
    if (n_chn == 0) return 0;   // (the reference would index a[0] here; min_chain_weight = 0 never drops a chain)
    ChainFltLess lt; lt.c = C;
    introsort((size_t)n_chn, a, lt);
#define CB(ci) (w.seeds[C[ci].head].qbeg)
#define CE(ci) (w.seeds[C[ci].tail].qbeg + w.seeds[C[ci].tail].len)
    int nk = 0;
    C[a[0]].kept = 3;
    kept_idx[nk++] = 0;
    for (i = 1; i < n_chn; ++i) {
        int large_ovlp = 0;
        for (k = 0; k < nk; ++k) {
            int j = kept_idx[k];
            int bi = CB(a[i]), ei = CE(a[i]), bj = CB(a[j]), ej = CE(a[j]);
            int b_max = bj > bi ? bj : bi;
            int e_min = ej < ei ? ej : ei;
            if (e_min > b_max && (!C[a[j]].is_alt || C[a[i]].is_alt)) {
                int li = ei - bi, lj = ej - bj;
                int min_l = li < lj ? li : lj;
                if (e_min - b_max >= min_l * opt.mask_level && min_l < opt.max_chain_gap) {
                    large_ovlp = 1;
                    if (C[a[j]].first < 0) C[a[j]].first = i;
                    if (C[a[i]].w < C[a[j]].w * opt.drop_ratio && C[a[j]].w - C[a[i]].w >= opt.min_seed_len << 1) break;
                }
            }
        }
        if (k == nk) {
            kept_idx[nk++] = i;
            C[a[i]].kept = large_ovlp ? 2 : 3;
        }
    }
    for (i = 0; i < nk; ++i) {
        Chain &c = C[a[kept_idx[i]]];
        if (c.first >= 0) C[a[c.first]].kept = 1;
    }
    for (i = k = 0; i < n_chn; ++i) {
        if (C[a[i]].kept == 0 || C[a[i]].kept == 3) continue;
        if (++k >= opt.max_chain_extend) break;
    }
    for (; i < n_chn; ++i)
        if (C[a[i]].kept < 3) C[a[i]].kept = 0;
    for (i = k = 0; i < n_chn; ++i)
        if (C[a[i]].kept != 0) a[k++] = a[i];
#undef CB
#undef CE
    return k;
}

} // namespace b200
