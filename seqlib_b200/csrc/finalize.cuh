// finalize.cuh -- regions -> final alignments for one read.
//   sort_dedup_patch <- mem_sort_dedup_patch (bwa/bwamem.c:463-515), mem_patch_reg (:432-461)
//   mark_primary_se  <- mem_mark_primary_se(_core) (bwa/bwamem.c:519-584), hash_64 (bwa/utils.h:98-109)
//   approx_mapq_se   <- mem_approx_mapq_se (bwa/bwamem.c:982-1006)
//   gen_cigar2       <- bwa_gen_cigar2 (bwa/bwa.c:148-234)
//   reg2aln          <- mem_reg2aln (bwa/bwamem.c:1119-1189), infer_bw (:818-825)
// Floating point: every double/float expression below is written so that nvcc
// emits the same IEEE-754 operations gcc -O2 emits for the reference on x86-64
// (no FMA contraction: the file is compiled with -fmad=false); log() of small
// integers comes from a table filled by the host libm (log_tab[i] = log(i)).
#pragma once
#include "common.cuh"
#include "fmindex.cuh"
#include "ksw.cuh"
#include "sort.cuh"

namespace b200 {

HD u64 hash_64(u64 key)
{
    key += ~(key << 32);
    key ^= (key >> 22);
    key += ~(key << 13);
    key ^= (key >> 8);
    key += (key << 3);
    key ^= (key >> 15);
    key += ~(key << 27);
    key ^= (key >> 31);
    return key;
}

struct FinScratch {
    EH *eh;           // max(l_query, ...)+1 cells
    u8 *z;            // direction bytes
    i64 z_cap;
    u8 *qbuf;         // l_query bytes: reversed query copy for reverse-strand DP
    i32 *zidx;        // n_regs ints (mark primary lists)
    const double *log_tab; int n_log;
};

struct GenCigarOut { int score, n_cigar, NM, md_len; bool ok, overflow; };

// Global alignment of query[0..l_query) against text [rb, re).  With cigar == NULL only
// the score is produced (the mem_patch_reg use).  MD is written to md[] (NUL-terminated).
template <class Ctr>
HD GenCigarOut gen_cigar2(const DevIndex &ix, const Opt &opt, int w_, int l_query, const u8 *query, i64 rb, i64 re,
                          FinScratch &fs, u32 *cigar, int cap_cigar, char *md, int cap_md, Ctr &ctr)
{
    GenCigarOut R; R.score = 0; R.n_cigar = 0; R.NM = -1; R.md_len = 0; R.ok = false; R.overflow = false;
    i64 l_pac = ix.l_pac;
    if (l_query <= 0 || rb >= re || (rb < l_pac && re > l_pac)) return R;
    // bns_get_seq clamps to [0, 2*l_pac); out of range => rlen != re-rb => no cigar
    if (re > l_pac << 1 || rb < 0) return R;
    i64 rlen = re - rb;
    ctr.ref_bytes += (unsigned long long)((rlen + 3) >> 2);
    // reverse strand: both sequences are walked backwards so that gaps are left-aligned (bwa/bwa.c:163-168)
    bool rev = rb >= l_pac;
    BytesSeq qs;
    TextSeqC ts(&ix, rev ? re - 1 : rb, rev ? -1 : 1);        // the current 32-base text word stays in registers
    if (rev) { qs.p = query + l_query - 1; qs.step = -1; }
    else { qs.p = query; qs.step = 1; }
    R.ok = true;
    if (l_query == rlen && w_ == 0) {
        if (cigar) { if (cap_cigar < 1) { R.overflow = true; return R; } cigar[0] = (u32)l_query << 4; R.n_cigar = 1; }
        int sc = 0;
        for (int i = 0; i < l_query; ++i) sc += opt.mat[ts[i] * 5 + qs[i]];
        R.score = sc;
    } else {
        int w, max_gap, max_ins, max_del, min_w;
        max_ins = (int)((double)(((l_query + 1) >> 1) * opt.mat[0] - opt.o_ins) / opt.e_ins + 1.);
        max_del = (int)((double)(((l_query + 1) >> 1) * opt.mat[0] - opt.o_del) / opt.e_del + 1.);
        max_gap = max_ins > max_del ? max_ins : max_del;
        max_gap = max_gap > 1 ? max_gap : 1;
        int dl = (int)rlen - l_query; dl = dl < 0 ? -dl : dl;
        w = (max_gap + dl + 1) >> 1;
        w = w < w_ ? w : w_;
        min_w = dl + 3;
        w = w > min_w ? w : min_w;
        u8 *z = 0;
        if (cigar) {
            int n_col = l_query < 2 * w + 1 ? l_query : 2 * w + 1;
            if ((i64)n_col * rlen > fs.z_cap) { R.overflow = true; return R; }
            z = fs.z;
        }
        int nc = 0;
        R.score = global2(l_query, qs, (int)rlen, ts, opt.mat, opt.o_del, opt.e_del, opt.o_ins, opt.e_ins, w,
                          fs.eh, z, cigar, cap_cigar, cigar ? &nc : 0, ctr);
        if (cigar) { if (nc < 0) { R.overflow = true; return R; } R.n_cigar = nc; }
    }
    if (cigar && md) {   // NM and MD (bwa/bwa.c:199-226)
        int k, x, y, u, n_mm = 0, n_gap = 0, l = 0;
        bool ovf = false;
        const char *int2base = rb < l_pac ? "ACGTN" : "TGCAN";
#define MD_PUTC(ch_) do { char ch__ = (ch_); if (l < cap_md - 1) md[l++] = ch__; else ovf = true; } while (0)
#define MD_PUTW(v_) do { int vv = (v_); char tb[12]; int tl = 0; if (vv == 0) tb[tl++] = '0'; \
        while (vv > 0) { tb[tl++] = (char)('0' + vv % 10); vv /= 10; } while (tl > 0) MD_PUTC(tb[--tl]); } while (0)
        for (k = 0, x = y = u = 0; k < R.n_cigar; ++k) {
            int op = cigar[k] & 0xf, len = (int)(cigar[k] >> 4);
            if (op == 0) {
                for (int i = 0; i < len; ++i) {
                    int rbase = ts[y + i];
                    if (qs[x + i] != rbase) { MD_PUTW(u); MD_PUTC(int2base[rbase]); ++n_mm; u = 0; }
                    else ++u;
                }
                x += len; y += len;
            } else if (op == 2) {
                if (k > 0 && k < R.n_cigar - 1) {
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
        if (ovf) { R.overflow = true; return R; }
        md[l] = 0;
        R.md_len = l; R.NM = n_mm + n_gap;
    }
    return R;
}

#define B200_PATCH_MAX_R_BW 0.05f
#define B200_PATCH_MIN_SC_RATIO 0.90f

template <class Ctr>
HD int patch_reg(const DevIndex &ix, const Opt &opt, const u8 *query, const Reg *a, const Reg *b, int *_w, FinScratch &fs, Ctr &ctr)
{
    int w, score, q_s, r_s;
    double r;
    if (a->rb < ix.l_pac && b->rb >= ix.l_pac) return 0;
    if (a->qb >= b->qb || a->qe >= b->qe || a->re >= b->re) return 0;
    w = (int)((a->re - b->rb) - (a->qe - b->qb));
    w = w > 0 ? w : -w;
    r = (double)(a->re - b->rb) / (b->re - a->rb) - (double)(a->qe - b->qb) / (b->qe - a->qb);
    r = r > 0. ? r : -r;
    if (a->re < b->rb || a->qe < b->qb) {
        if (w > opt.w << 1 || r >= B200_PATCH_MAX_R_BW) return 0;
    } else if (w > opt.w << 2 || r >= B200_PATCH_MAX_R_BW * 2) return 0;
    w += a->w + b->w;
    w = w < opt.w << 2 ? w : opt.w << 2;
    GenCigarOut g = gen_cigar2(ix, opt, w, b->qe - a->qb, query + a->qb, a->rb, b->re, fs, (u32 *)0, 0, (char *)0, 0, ctr);
    score = g.score;   // (left at 0 when bwa_gen_cigar2 bails out early; the reference reads an unset int there)
    q_s = (int)((double)(b->qe - a->qb) / ((b->qe - b->qb) + (a->qe - a->qb)) * (b->score + a->score) + .499);
    r_s = (int)((double)(b->re - a->rb) / ((b->re - b->rb) + (a->re - a->rb)) * (b->score + a->score) + .499);
    if ((double)score / (q_s > r_s ? q_s : r_s) < B200_PATCH_MIN_SC_RATIO) return 0;
    *_w = w;
    return score;
}

struct RegLessRe { HD bool operator()(const Reg &a, const Reg &b) const { return a.re < b.re; } };
struct RegLessScore {
    HD bool operator()(const Reg &a, const Reg &b) const {
        return a.score > b.score || (a.score == b.score && (a.rb < b.rb || (a.rb == b.rb && a.qb < b.qb)));
    }
};
struct RegLessHash {
    HD bool operator()(const Reg &a, const Reg &b) const {
        return a.score > b.score || (a.score == b.score && (a.is_alt < b.is_alt || (a.is_alt == b.is_alt && a.hash < b.hash)));
    }
};
struct RegLessHash2 {
    HD bool operator()(const Reg &a, const Reg &b) const {
        return a.is_alt < b.is_alt || (a.is_alt == b.is_alt && (a.score > b.score || (a.score == b.score && a.hash < b.hash)));
    }
};

template <class Ctr>
HD int sort_dedup_patch(const DevIndex &ix, const Opt &opt, const u8 *query, int n, Reg *a, FinScratch &fs, Ctr &ctr)
{
    int m, i, j;
    if (n <= 1) return n;
    introsort((size_t)n, a, RegLessRe());
    for (i = 0; i < n; ++i) a[i].n_comp = 1;
    for (i = 1; i < n; ++i) {
        Reg *p = &a[i];
        if (p->rid != a[i - 1].rid || p->rb >= a[i - 1].re + opt.max_chain_gap) continue;
        for (j = i - 1; j >= 0 && p->rid == a[j].rid && p->rb < a[j].re + opt.max_chain_gap; --j) {
            Reg *q = &a[j];
            i64 orr, oq, mr, mq;
            int score, w;
            if (q->qe == q->qb) continue;
            orr = q->re - p->rb;
            oq = q->qb < p->qb ? q->qe - p->qb : p->qe - q->qb;
            mr = q->re - q->rb < p->re - p->rb ? q->re - q->rb : p->re - p->rb;
            mq = q->qe - q->qb < p->qe - p->qb ? q->qe - q->qb : p->qe - p->qb;
            if (orr > opt.mask_level_redun * mr && oq > opt.mask_level_redun * mq) {
                if (p->score < q->score) { p->qe = p->qb; break; }
                else q->qe = q->qb;
            } else if (q->rb < p->rb && (score = patch_reg(ix, opt, query, q, p, &w, fs, ctr)) > 0) {
                p->n_comp += q->n_comp + 1;
                p->seedcov = p->seedcov > q->seedcov ? p->seedcov : q->seedcov;
                p->sub = p->sub > q->sub ? p->sub : q->sub;
                p->csub = p->csub > q->csub ? p->csub : q->csub;
                p->qb = q->qb; p->rb = q->rb;
                p->truesc = p->score = score;
                p->w = w;
                q->qb = q->qe;
            }
        }
    }
    for (i = 0, m = 0; i < n; ++i)
        if (a[i].qe > a[i].qb) { if (m != i) a[m++] = a[i]; else ++m; }
    n = m;
    introsort((size_t)n, a, RegLessScore());
    for (i = 1; i < n; ++i)
        if (a[i].score == a[i - 1].score && a[i].rb == a[i - 1].rb && a[i].qb == a[i - 1].qb) a[i].qe = a[i].qb;
    for (i = 1, m = 1; i < n; ++i)
        if (a[i].qe > a[i].qb) { if (m != i) a[m++] = a[i]; else ++m; }
    return m;
}

HD void mark_primary_core(const Opt &opt, int n, Reg *a, i32 *z)
{
    int i, k, tmp, nz = 0;
    tmp = opt.a + opt.b;
    tmp = opt.o_del + opt.e_del > tmp ? opt.o_del + opt.e_del : tmp;
    tmp = opt.o_ins + opt.e_ins > tmp ? opt.o_ins + opt.e_ins : tmp;
    z[nz++] = 0;
    for (i = 1; i < n; ++i) {
        for (k = 0; k < nz; ++k) {
            int j = z[k];
            int b_max = a[j].qb > a[i].qb ? a[j].qb : a[i].qb;
            int e_min = a[j].qe < a[i].qe ? a[j].qe : a[i].qe;
            if (e_min > b_max) {
                int min_l = a[i].qe - a[i].qb < a[j].qe - a[j].qb ? a[i].qe - a[i].qb : a[j].qe - a[j].qb;
                if (e_min - b_max >= min_l * opt.mask_level) {
                    if (a[j].sub == 0) a[j].sub = a[i].score;
                    if (a[j].score - a[i].score <= tmp && (a[j].is_alt || !a[i].is_alt)) ++a[j].sub_n;
                    break;
                }
            }
        }
        if (k == nz) z[nz++] = i;
        else a[i].secondary = z[k];
    }
}

// z: 2*n ints of scratch
HD int mark_primary_se(const Opt &opt, int n, Reg *a, i64 id, i32 *z)
{
    int i, n_pri;
    if (n == 0) return 0;
    for (i = n_pri = 0; i < n; ++i) {
        a[i].sub = a[i].alt_sc = 0; a[i].secondary = a[i].secondary_all = -1; a[i].hash = hash_64((u64)(id + i));
        if (!a[i].is_alt) ++n_pri;
    }
    introsort((size_t)n, a, RegLessHash());
    mark_primary_core(opt, n, a, z);
    for (i = 0; i < n; ++i) {
        Reg *p = &a[i];
        p->secondary_all = i;
        if (!p->is_alt && p->secondary >= 0 && a[p->secondary].is_alt) p->alt_sc = a[p->secondary].score;
    }
    if (n_pri >= 0 && n_pri < n) {
        if (n_pri > 0) introsort((size_t)n, a, RegLessHash2());
        for (i = 0; i < n; ++i) z[a[i].secondary_all] = i;
        for (i = 0; i < n; ++i) {
            if (a[i].secondary >= 0) {
                a[i].secondary_all = z[a[i].secondary];
                if (a[i].is_alt) a[i].secondary = 0x7fffffff;
            } else a[i].secondary_all = -1;
        }
        if (n_pri > 0) {
            for (i = 0; i < n_pri; ++i) { a[i].sub = 0; a[i].secondary = -1; }
            mark_primary_core(opt, n_pri, a, z + n);
        }
    } else {
        for (i = 0; i < n; ++i) a[i].secondary_all = a[i].secondary;
    }
    return n_pri;
}

HD double tab_log(const FinScratch &fs, int v) { return v >= 0 && v < fs.n_log ? fs.log_tab[v] : 0.0; }

HD int approx_mapq_se(const Opt &opt, const Reg *a, const FinScratch &fs, bool *need_host)
{
    int mapq, l, sub = a->sub ? a->sub : opt.min_seed_len * opt.a;
    double identity;
    sub = a->csub > sub ? a->csub : sub;
    if (sub >= a->score) return 0;
    l = a->qe - a->qb > a->re - a->rb ? a->qe - a->qb : (int)(a->re - a->rb);
    identity = 1. - (double)(l * opt.a - a->score) / (opt.a + opt.b) / l;
    if (a->score == 0) mapq = 0;
    else if (opt.mapQ_coef_len > 0) {
        double tmp;
        if (l >= fs.n_log) *need_host = true;
        tmp = l < opt.mapQ_coef_len ? 1. : opt.mapQ_coef_fac / tab_log(fs, l);
        tmp *= identity * identity;
        mapq = (int)(6.02 * (a->score - sub) / opt.a * tmp * tmp + .499);
    } else {
        if (a->seedcov >= fs.n_log) *need_host = true;
        mapq = (int)(30.0 * (1. - (double)sub / a->score) * tab_log(fs, a->seedcov) + .499);   // MEM_MAPQ_COEF = 30.0
        mapq = identity < 0.95 ? (int)(mapq * identity * identity + .499) : mapq;
    }
    if (a->sub_n > 0) {
        if (a->sub_n + 1 >= fs.n_log) *need_host = true;
        mapq -= (int)(4.343 * tab_log(fs, a->sub_n + 1) + .499);
    }
    if (mapq > 60) mapq = 60;
    if (mapq < 0) mapq = 0;
    mapq = (int)(mapq * (1. - a->frac_rep) + .499);
    return mapq;
}

HD int infer_bw(int l1, int l2, int score, int a, int q, int r)
{
    int w;
    if (l1 == l2 && l1 * a - score < (q + r - a) << 1) return 0;
    w = (int)((double)((l1 < l2 ? l1 : l2) * a - score - q) / r + 2.);
    int d = l1 - l2; d = d < 0 ? -d : d;
    if (w < d) w = d;
    return w;
}

struct AlnOut {      // mem_aln_t essentials
    i64 pos; int rid, flag, is_rev, mapq, NM, n_cigar, md_len, score, sub;
    bool overflow, need_host;
};

// first part of mem_reg2aln (bwa/bwamem.c:1136-1144): MAPQ, secondary flag, inferred band width
struct AlnPlan { int mapq, flag, w2; bool need_host; };

HD AlnPlan reg2aln_plan(const Opt &opt, const Reg *ar, const FinScratch &fs)
{
    AlnPlan p; p.need_host = false; p.flag = 0;
    int qb = ar->qb, qe = ar->qe, tmp, w2;
    i64 rb = ar->rb, re = ar->re;
    p.mapq = ar->secondary < 0 ? approx_mapq_se(opt, ar, fs, &p.need_host) : 0;
    if (ar->secondary >= 0) p.flag |= 0x100;
    tmp = infer_bw(qe - qb, (int)(re - rb), ar->truesc, opt.a, opt.o_del, opt.e_del);
    w2 = infer_bw(qe - qb, (int)(re - rb), ar->truesc, opt.a, opt.o_ins, opt.e_ins);
    w2 = w2 > tmp ? w2 : tmp;
    if (w2 > opt.w) w2 = w2 < ar->w ? w2 : ar->w;
    p.w2 = w2;
    return p;
}

// true when every bwa_gen_cigar2 call of the retry loop takes the no-DP branch (bwa/bwa.c:169-178): w2 = 0 stays 0 when doubled
HD bool reg2aln_is_ungapped(const Reg *ar, int w2) { return w2 == 0 && (i64)(ar->qe - ar->qb) == ar->re - ar->rb; }

// last part of mem_reg2aln (bwa/bwamem.c:1157-1187): strand/pos, squeeze a leading/trailing deletion, add clips
HD void reg2aln_finish(const DevIndex &ix, int l_query, const Reg *ar, u32 *cigar, AlnOut &a)
{
    int qb = ar->qb, qe = ar->qe, i, is_rev;
    i64 rb = ar->rb, re = ar->re;
    i64 pos = depos(ix, rb < ix.l_pac ? rb : re - 1, &is_rev);
    a.is_rev = is_rev;
    if (a.n_cigar > 0) {
        if ((cigar[0] & 0xf) == 2) {
            pos += cigar[0] >> 4;
            --a.n_cigar;
            for (i = 0; i < a.n_cigar; ++i) cigar[i] = cigar[i + 1];
        } else if ((cigar[a.n_cigar - 1] & 0xf) == 2) --a.n_cigar;
    }
    if (qb != 0 || qe != l_query) {
        int clip5 = is_rev ? l_query - qe : qb;
        int clip3 = is_rev ? qb : l_query - qe;
        if (clip5) {
            for (i = a.n_cigar; i > 0; --i) cigar[i] = cigar[i - 1];
            cigar[0] = (u32)clip5 << 4 | 3;
            ++a.n_cigar;
        }
        if (clip3) cigar[a.n_cigar++] = (u32)clip3 << 4 | 3;
    }
    a.rid = pos2rid(ix, pos);
    a.pos = pos - ix.contig_off[a.rid];
    a.score = ar->score; a.sub = ar->sub > ar->csub ? ar->sub : ar->csub;
}

// cigar: capacity cap_cigar words (incl. room for two clips); md: cap_md bytes.
template <class Ctr>
HD AlnOut reg2aln(const DevIndex &ix, const Opt &opt, int l_query, const u8 *query, const Reg *ar,
                  FinScratch &fs, u32 *cigar, int cap_cigar, char *md, int cap_md, Ctr &ctr)
{
    AlnOut a;
    a.pos = -1; a.rid = -1; a.flag = 0; a.is_rev = 0; a.mapq = 0; a.NM = 0; a.n_cigar = 0; a.md_len = 0; a.score = 0; a.sub = 0;
    a.overflow = false; a.need_host = false;
    if (ar->rb < 0 || ar->re < 0) { a.flag |= 0x4; return a; }
    AlnPlan pl = reg2aln_plan(opt, ar, fs);
    a.mapq = pl.mapq; a.flag = pl.flag; a.need_host = pl.need_host;
    int qb = ar->qb, qe = ar->qe, i = 0, w2 = pl.w2, score = 0, last_sc = -(1 << 30);
    i64 rb = ar->rb, re = ar->re;
    GenCigarOut g;
    g.n_cigar = 0; g.NM = -1; g.md_len = 0; g.ok = false;
    do {
        w2 = w2 < opt.w << 2 ? w2 : opt.w << 2;
        g = gen_cigar2(ix, opt, w2, qe - qb, query + qb, rb, re, fs, cigar, cap_cigar - 2, md, cap_md, ctr);
        if (g.overflow) { a.overflow = true; return a; }
        score = g.score;
        if (score == last_sc || w2 == opt.w << 2) break;
        last_sc = score;
        w2 <<= 1;
    } while (++i < 3 && score < ar->truesc - opt.a);
    a.NM = g.NM; a.n_cigar = g.n_cigar; a.md_len = g.md_len;
    reg2aln_finish(ix, l_query, ar, cigar, a);
    return a;
}

} // namespace b200
