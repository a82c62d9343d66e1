// finalize_group.cuh -- the gapped part of mem_reg2aln (bwa/bwamem.c:1145-1156: up to three bwa_gen_cigar2 calls with
// a doubling band) for hits that need a real global alignment, G lanes per hit (device only).
// stage_finalize (one thread per read) resolves every hit whose alignment is provably ungapped inline and queues the
// rest as DpJob records; k_finalize_dp drains that queue: DP rows and traceback by the group (global2_group, traceback_group), NM/MD +
// clipping by lane 0, then the CIGAR/MD are appended to the pools and the hit record is completed.
#pragma once
#include "pipeline.cuh"
#include "ksw_group.cuh"

namespace b200 {

__host__ __device__ inline size_t findp_smem_bytes(int maxlen) { return (size_t)(maxlen + 2) * 8 + (size_t)((maxlen + 4) & ~3); }
__host__ __device__ inline size_t findp_scratch_bytes(const Caps &c) { return (size_t)c.z + 8 + sizeof(u32) * (size_t)c.cigar + (size_t)c.md + 64; }

// traceback of ksw_global2 (bwa/ksw.c:620-637), cooperative: the walk is one dependent load per step, and it mostly runs along a
// diagonal.  The lanes of the group fetch the next G cells of the current diagonal at once; every lane then replays the same steps
// on the shuffled bytes (uniform control flow) until the path leaves the diagonal, where the group refetches.  Lane 0 writes the
// CIGAR words.  Returns the number of words (group-uniform), -1 on overflow.
template <int G>
__device__ inline int traceback_group(const GroupCtx<G> &g, const u8 *z, int n_col, int qlen, int tlen, int w, u32 *cigar, int cap_cigar)
{
    int n = 0, which = 0, i = tlen - 1, k = (i + w + 1 < qlen ? i + w + 1 : qlen) - 1;
    bool ovf = false;
    u32 last = 0;                 // the CIGAR word being accumulated (kept in registers by every lane, stored by lane 0)
#define PUSH_OP(op_, len_) do { \
    if (n == 0 || (int)(last & 0xf) != (op_)) { \
        if (n > 0 && g.gl == 0) cigar[n - 1] = last; \
        if (n < cap_cigar) { last = (u32)(len_) << 4 | (op_); ++n; } else ovf = true; } \
    else last += (u32)(len_) << 4; } while (0)
    while (i >= 0 && k >= 0 && !ovf) {
        const int ti = i - g.gl, tk = k - g.gl;
        int zz = 0;
        if (ti >= 0 && tk >= 0) zz = z[(i64)ti * n_col + (tk - (ti > w ? ti - w : 0))];
        for (int t = 0; t < G; ++t) {
            if (i < 0 || k < 0) break;
            const int zt = g.bcast(zz, t);
            which = zt >> (which << 1) & 3;
            if (which == 0) { PUSH_OP(0, 1); --i; --k; }
            else if (which == 1) { PUSH_OP(2, 1); --i; break; }
            else { PUSH_OP(1, 1); --k; break; }
            if (ovf) break;
        }
    }
    if (!ovf && i >= 0) PUSH_OP(2, i + 1);
    if (!ovf && k >= 0) PUSH_OP(1, k + 1);
#undef PUSH_OP
    if (ovf) return -1;
    if (g.gl == 0) {
        if (n > 0) cigar[n - 1] = last;
        for (i = 0; i < n >> 1; ++i) swap_(cigar[i], cigar[n - 1 - i]);
    }
    return n;
}

// bwa_gen_cigar2 (bwa/bwa.c:148-234) for one hit, cooperative.  q: the hit's query slice staged in shared memory in DP
// orientation (already reversed for the reverse strand).  Results are group-uniform except cigar/md contents (lane 0 wrote them).
template <int G>
__device__ GenCigarOut gen_cigar2_group(const GroupCtx<G> &g, const DevIndex &ix, const Opt &opt, const i8 *smat, int w_, int l_query, const u8 *q,
                                        i64 rb, i64 re, int *H, int *E, u8 *z, i64 z_cap, u32 *cigar, int cap_cigar, char *md, int cap_md, CtrLocal &ctr)
{
    GenCigarOut R; R.score = 0; R.n_cigar = 0; R.NM = -1; R.md_len = 0; R.ok = false; R.overflow = false;
    i64 l_pac = ix.l_pac;
    if (l_query <= 0 || rb >= re || (rb < l_pac && re > l_pac)) return R;
    if (re > l_pac << 1 || rb < 0) return R;
    i64 rlen = re - rb;
    bool rev = rb >= l_pac;
    BytesSeq qs; qs.p = q; qs.step = 1;
    TextSeqC ts(&ix, rev ? re - 1 : rb, rev ? -1 : 1);        // the current 32-base text word stays in registers
    R.ok = true;
    int nc = 0;
    if (l_query == rlen && w_ == 0) {
        int sc = 0;
        for (int i = g.gl; i < l_query; i += G) sc += smat[ts[i] * 5 + q[i]];
        for (int o = G >> 1; o > 0; o >>= 1) sc += __shfl_xor_sync(g.mask, sc, o, G);
        R.score = sc;
        if (cap_cigar < 1) { R.overflow = true; return R; }
        if (g.gl == 0) cigar[0] = (u32)l_query << 4;
        nc = 1;
    } else {
        int w, max_gap, max_ins, max_del, min_w;
        max_ins = (int)((double)(((l_query + 1) >> 1) * smat[0] - opt.o_ins) / opt.e_ins + 1.);
        max_del = (int)((double)(((l_query + 1) >> 1) * smat[0] - opt.o_del) / opt.e_del + 1.);
        max_gap = max_ins > max_del ? max_ins : max_del;
        max_gap = max_gap > 1 ? max_gap : 1;
        int dl = (int)rlen - l_query; dl = dl < 0 ? -dl : dl;
        w = (max_gap + dl + 1) >> 1;
        w = w < w_ ? w : w_;
        min_w = dl + 3;
        w = w > min_w ? w : min_w;
        int n_col = l_query < 2 * w + 1 ? l_query : 2 * w + 1;
        if ((i64)n_col * rlen > z_cap) { R.overflow = true; return R; }
        unsigned long long cells = 0;
        R.score = global2_group(g, l_query, qs, (int)rlen, ts, smat, opt.o_del, opt.e_del, opt.o_ins, opt.e_ins, w, H, E, z, &cells);
        if (g.gl == 0) { ctr.sw_cells += cells; ctr.n_global++; }
        g.sync();                      // the direction bytes of all lanes are visible
        nc = traceback_group<G>(g, z, n_col, l_query, (int)rlen, w, cigar, cap_cigar);
        if (nc < 0) { R.overflow = true; return R; }
    }
    R.n_cigar = nc;
    g.sync();
    // NM / MD by lane 0 (bwa/bwa.c:199-226)
    int res[3] = {0, 0, 0};
    if (g.gl == 0) {
        int k, x, y, u, n_mm = 0, n_gap = 0, l = 0;
        bool ovf = false;
        const char *int2base = rb < l_pac ? "ACGTN" : "TGCAN";
#define MD_PUTC(ch_) do { char ch__ = (ch_); if (l < cap_md - 1) md[l++] = ch__; else ovf = true; } while (0)
#define MD_PUTW(v_) do { int vv = (v_); char tb[12]; int tl = 0; if (vv == 0) tb[tl++] = '0'; \
        while (vv > 0) { tb[tl++] = (char)('0' + vv % 10); vv /= 10; } while (tl > 0) MD_PUTC(tb[--tl]); } while (0)
        for (k = 0, x = y = u = 0; k < nc; ++k) {
            int op = cigar[k] & 0xf, len = (int)(cigar[k] >> 4);
            if (op == 0) {
                for (int i = 0; i < len; ++i) {
                    int rbase = ts[y + i];
                    if (q[x + i] != rbase) { MD_PUTW(u); MD_PUTC(int2base[rbase]); ++n_mm; u = 0; }
                    else ++u;
                }
                x += len; y += len;
            } else if (op == 2) {
                if (k > 0 && k < nc - 1) {
                    MD_PUTW(u); MD_PUTC('^');
                    for (int i = 0; i < len; ++i) MD_PUTC(int2base[ts[y + i]]);
                    u = 0; n_gap += len;
                }
                y += len;
            } else if (op == 1) { x += len; n_gap += len; }
        }
        MD_PUTW(u);
#undef MD_PUTC
#undef MD_PUTW
        md[l < cap_md ? l : cap_md - 1] = 0;
        res[0] = ovf ? -1 : l; res[1] = n_mm + n_gap;
    }
    res[0] = __shfl_sync(g.mask, res[0], 0, G);
    res[1] = __shfl_sync(g.mask, res[1], 0, G);
    if (res[0] < 0) { R.overflow = true; return R; }
    R.md_len = res[0]; R.NM = res[1];
    return R;
}

// One queued hit: run the band-doubling loop of mem_reg2aln, finish the record, publish CIGAR/MD.
template <int G>
__device__ void finalize_dp_job(const GroupCtx<G> &g, const DevIndex &ix, const Opt &opt, const Caps &caps, const Batch &B, const DpJob &job,
                                u8 *scratch, u8 *smem, const i8 *smat, CtrLocal &ctr)
{
    i64 rid = job.rid;
    if (B.ovf[rid]) return;
    b200_hit_t *h = B.pool.hits + job.hit;
    int l_query = (int)(B.seq_off[rid + 1] - B.seq_off[rid]);
    const u8 *seq = B.seq + B.seq_off[rid];
    int *H = (int *)smem, *E = H + (caps.maxlen + 2);
    u8 *q = (u8 *)(E + (caps.maxlen + 2));
    u8 *p = scratch;
    u8 *z = p; p = align8(p + caps.z);
    u32 *cg = (u32 *)p; p += sizeof(u32) * (size_t)caps.cigar;
    char *md = (char *)p;
    Reg ar;   // only the fields reg2aln reads
    ar.rb = h->rb; ar.re = h->re; ar.qb = h->qb; ar.qe = h->qe; ar.truesc = h->truesc; ar.w = h->w; ar.score = h->score; ar.sub = h->sub; ar.csub = h->csub;
    int qb = ar.qb, qe = ar.qe, lq = qe - qb;
    bool rev = ar.rb >= ix.l_pac;
    for (int j = g.gl; j < lq; j += G) q[j] = rev ? seq[qe - 1 - j] : seq[qb + j];
    g.sync();
    int i = 0, w2 = job.w2, score = 0, last_sc = -(1 << 30);
    GenCigarOut gc; gc.n_cigar = 0; gc.NM = -1; gc.md_len = 0; gc.overflow = false;
    do {
        w2 = w2 < opt.w << 2 ? w2 : opt.w << 2;
        gc = gen_cigar2_group(g, ix, opt, smat, w2, lq, q, ar.rb, ar.re, H, E, z, caps.z, cg, caps.cigar - 2, md, caps.md, ctr);
        if (gc.overflow) { if (g.gl == 0) atomicOr(&B.ovf[rid], (u32)OVF_OUT); return; }
        score = gc.score;
        if (score == last_sc || w2 == opt.w << 2) break;
        last_sc = score;
        w2 <<= 1;
        g.sync();
    } while (++i < 3 && score < ar.truesc - opt.a);
    if (g.gl == 0) {
        AlnOut a;
        a.n_cigar = gc.n_cigar; a.md_len = gc.md_len; a.NM = gc.NM;
        reg2aln_finish(ix, l_query, &ar, cg, a);
        i64 co = pool_alloc(B.pool, POOL_CIGAR, a.n_cigar), mo = pool_alloc(B.pool, POOL_MD, a.md_len + 1);
        if (co < 0 || mo < 0) { atomicOr(&B.ovf[rid], (u32)OVF_POOL); }
        else {
            for (int k = 0; k < a.n_cigar; ++k) B.pool.cigar[co + k] = cg[k];
            for (int k = 0; k < a.md_len; ++k) B.pool.md[mo + k] = md[k];
            B.pool.md[mo + a.md_len] = 0;
            h->pos = a.pos; h->is_rev = a.is_rev; h->NM = a.NM; h->aln_sub = a.sub; h->n_cigar = a.n_cigar; h->md_len = a.md_len;
            h->cigar_off = co; h->md_off = mo;
            if (a.rid != h->rid) h->rid = -1000;
            atomicAdd(&B.rec[rid].n_cigar, a.n_cigar);
            atomicAdd(&B.rec[rid].n_md, a.md_len + 1);
        }
    }
    g.sync();
}

} // namespace b200
