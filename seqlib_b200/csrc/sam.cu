// sam.cu -- SAM text of single-end alignments, host code over the hits b200_mem_align_batch returns.
//   b200_results_to_sam <- mem_reg2sam (bwa/bwamem.c:1034-1086): which regions become records, supplementary flag and
//                          MAPQ cap, the unaligned record;
//                          mem_gen_alt (bwa/bwamem_extra.c:117-173): the XA / XB strings of the secondary hits;
//                          mem_aln2sam + add_cigar (bwa/bwamem.c:837-976): the eleven columns, S -> H for supplementary
//                          records with the bases trimmed to match, NM MD AS XS SA pa XA tags, the read comment.
// Nothing is recomputed: every region already carries what mem_reg2aln gave it (position, strand, MAPQ, NM, CIGAR, MD),
// because SeqLib's aligner asks for all of them (src/BWAAligner.cpp:117-128).  Records of different reads are independent:
// the reads are formatted on all host threads, each into its own buffer, and concatenated in read order.
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <functional>
#include <thread>
#include <vector>
#include "../../include/seqlib_b200.h"

namespace b200 { void set_error(const std::string &msg); }

namespace {

enum { F_ALL = 0x8, F_NO_MULTI = 0x10, F_REF_HDR = 0x100, F_SOFTCLIP = 0x200, F_KEEP_SUPP_MAPQ = 0x1000, F_XB = 0x2000 };   // bwa/bwamem.h:40-50

// append-only byte buffer; the hot appends (bases, qualities, numbers) write through a raw pointer after one capacity check
struct Out {
    char *p = nullptr; size_t n = 0, cap = 0;
    Out() {}
    Out(const Out &) = delete;
    Out &operator=(const Out &) = delete;
    Out(Out &&o) noexcept : p(o.p), n(o.n), cap(o.cap) { o.p = nullptr; o.n = o.cap = 0; }
    ~Out() { free(p); }
    void reserve(size_t extra)
    {
        if (n + extra <= cap) return;
        size_t c = cap ? cap : 1024;
        while (c < n + extra) c += c >> 1;
        p = (char *)realloc(p, c); cap = c;
    }
    char *grow(size_t k) { reserve(k); char *w = p + n; n += k; return w; }
    void num(long long v)
    {
        reserve(24);
        char b[24]; int k = 0;
        unsigned long long u = v < 0 ? 0ull - (unsigned long long)v : (unsigned long long)v;
        do { b[k++] = (char)('0' + u % 10); u /= 10; } while (u);
        if (v < 0) p[n++] = '-';
        while (k) p[n++] = b[--k];
    }
    void ch(char c) { reserve(1); p[n++] = c; }
    void strn(const char *q, size_t k) { memcpy(grow(k), q, k); }
    void str(const char *q) { strn(q, strlen(q)); }
    bool empty() const { return n == 0; }
};

struct Aln {            // mem_aln_t (bwa/bwamem.h:115-126) assembled from a b200_hit_t
    const b200_hit_t *h;
    int flag, mapq, sub;
    const Out *xa;
};

struct Ctx {
    const char *const *rnames; int n_rnames; const b200_mem_opt_t *opt; b200_results_view_t v;
    const char *seqs; const int64_t *seq_off; const char *quals; const int64_t *qual_off;
    const char *names; const int64_t *name_off; const char *comments; const int64_t *comment_off;
};

inline int nt4(unsigned char c)
{
    switch (c) { case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2; case 'T': case 't': return 3; default: return c < 4 ? c : 4; }
}

// "ACGTN"[code] / "TGCAN"[code] of every input byte (mem_aln2sam prints the nt4 codes of the read, bwa/bwamem.c:902-921)
struct BaseTabs {
    char fwd[256], rev[256];
    BaseTabs() { for (int c = 0; c < 256; ++c) { fwd[c] = "ACGTN"[nt4((unsigned char)c)]; rev[c] = "TGCAN"[nt4((unsigned char)c)]; } }
};
const BaseTabs g_tabs;

void add_cigar(const Ctx &C, const b200_hit_t &h, Out &o, int which)
{
    if (!h.n_cigar) { o.ch('*'); return; }
    const uint32_t *cg = C.v.cigar + h.cigar_off;
    for (int i = 0; i < h.n_cigar; ++i) {
        int c = cg[i] & 0xf;
        if (!(C.opt->flag & F_SOFTCLIP) && !h.is_alt && (c == 3 || c == 4)) c = which ? 4 : 3;
        o.num(cg[i] >> 4); o.ch("MIDSH"[c]);
    }
}

// one record; h == nullptr: the unaligned record (mem_reg2aln with ar == 0: everything zero, rid = pos = -1, flag 4)
void aln2sam(const Ctx &C, int64_t read, const Aln *list, size_t n_list, int which, Out &o)
{
    static const b200_hit_t zero = {};
    const Aln &p = list[(size_t)which];
    const b200_hit_t &h = p.h ? *p.h : zero;
    const int rid = p.h ? h.rid : -1;
    const bool is_rev = p.h && h.is_rev;
    const int n_cigar = p.h ? h.n_cigar : 0;
    int flag = p.flag;
    flag |= rid < 0 ? 0x4 : 0;
    flag |= is_rev ? 0x10 : 0;
    o.reserve((size_t)(C.name_off[read + 1] - C.name_off[read]) + 2 * (size_t)(C.seq_off[read + 1] - C.seq_off[read]) + 256);   // the common case never reallocates below
    o.strn(C.names + C.name_off[read], (size_t)(C.name_off[read + 1] - C.name_off[read])); o.ch('\t');
    o.num((flag & 0xffff) | (flag & 0x10000 ? 0x100 : 0)); o.ch('\t');
    if (rid >= 0) {
        o.str(C.rnames[rid]); o.ch('\t');
        o.num((long long)h.pos + 1); o.ch('\t');
        o.num(p.mapq); o.ch('\t');
        add_cigar(C, h, o, which);
    } else o.str("*\t0\t0\t*");
    o.ch('\t');
    o.str("*\t0\t0");
    o.ch('\t');
    const char *seq = C.seqs + C.seq_off[read];
    const int l_seq = (int)(C.seq_off[read + 1] - C.seq_off[read]);
    const char *qual = C.quals && C.qual_off[read + 1] > C.qual_off[read] ? C.quals + C.qual_off[read] : nullptr;
    if (flag & 0x100) o.str("*\t*");
    else {
        int qb = 0, qe = l_seq;
        const uint32_t *cg = p.h ? C.v.cigar + h.cigar_off : nullptr;
        if (n_cigar && which && !(C.opt->flag & F_SOFTCLIP) && !h.is_alt) {
            int first = cg[0] & 0xf, last = cg[n_cigar - 1] & 0xf;
            if (!is_rev) {
                if (first == 4 || first == 3) qb += cg[0] >> 4;
                if (last == 4 || last == 3) qe -= cg[n_cigar - 1] >> 4;
            } else {
                if (first == 4 || first == 3) qe -= cg[0] >> 4;
                if (last == 4 || last == 3) qb += cg[n_cigar - 1] >> 4;
            }
        }
        const int len = qe > qb ? qe - qb : 0;
        if (!is_rev) {
            char *w = o.grow((size_t)len + 1);
            for (int i = qb; i < qe; ++i) *w++ = g_tabs.fwd[(unsigned char)seq[i]];
            *w = '\t';
            if (qual) o.strn(qual + qb, (size_t)len); else o.ch('*');
        } else {
            char *w = o.grow((size_t)len + 1);
            for (int i = qe - 1; i >= qb; --i) *w++ = g_tabs.rev[(unsigned char)seq[i]];
            *w = '\t';
            if (qual) { char *x = o.grow((size_t)len); for (int i = qe - 1; i >= qb; --i) *x++ = qual[i]; } else o.ch('*');
        }
    }
    if (n_cigar) {
        o.str("\tNM:i:"); o.num(h.NM);
        o.str("\tMD:Z:"); o.str(C.v.md + h.md_off);
    }
    const int score = p.h ? h.score : 0;
    if (score >= 0) { o.str("\tAS:i:"); o.num(score); }
    if (p.sub >= 0) { o.str("\tXS:i:"); o.num(p.sub); }
    if (!(flag & 0x100)) {
        size_t i;
        for (i = 0; i < n_list; ++i) if ((int)i != which && !(list[i].flag & 0x100)) break;
        if (i < n_list) {
            o.str("\tSA:Z:");
            for (i = 0; i < n_list; ++i) {
                const Aln &r = list[i];
                if ((int)i == which || (r.flag & 0x100)) continue;
                o.str(C.rnames[r.h->rid]); o.ch(',');
                o.num((long long)r.h->pos + 1); o.ch(',');
                o.ch("+-"[r.h->is_rev ? 1 : 0]); o.ch(',');
                const uint32_t *rc = C.v.cigar + r.h->cigar_off;
                for (int k = 0; k < r.h->n_cigar; ++k) { o.num(rc[k] >> 4); o.ch("MIDSH"[rc[k] & 0xf]); }
                o.ch(','); o.num(r.mapq);
                o.ch(','); o.num(r.h->NM);
                o.ch(';');
            }
        }
        if (p.h && h.alt_sc > 0) { char b[64]; snprintf(b, sizeof b, "\tpa:f:%.3f", (double)h.score / h.alt_sc); o.str(b); }
    }
    if (p.xa && !p.xa->empty()) { o.str((C.opt->flag & F_XB) ? "\tXB:Z:" : "\tXA:Z:"); o.strn(p.xa->p, p.xa->n); }
    if (C.comments && C.comment_off[read + 1] > C.comment_off[read]) {
        o.ch('\t'); o.strn(C.comments + C.comment_off[read], (size_t)(C.comment_off[read + 1] - C.comment_off[read]));
    }
    o.ch('\n');
}

int pri_idx(double drop, const b200_hit_t *a, int i)        // get_pri_idx (bwa/bwamem_extra.c:117-122)
{
    int k = a[i].secondary_all;
    if (k >= 0 && a[i].score >= a[k].score * drop) return k;
    return -1;
}

void read2sam(const Ctx &C, int64_t read, Out &o)
{
    const b200_mem_opt_t &opt = *C.opt;
    const b200_hit_t *a = C.v.hits + C.v.hit_off[read];
    const int n = (int)(C.v.hit_off[read + 1] - C.v.hit_off[read]);
    std::vector<Out> xa;
    if (!(opt.flag & F_ALL) && n > 1) {                          // mem_gen_alt (a lone region has no secondary_all >= 0)
        int tot = 0;
        for (int i = 0; i < n; ++i) if (pri_idx(opt.XA_drop_ratio, a, i) >= 0) ++tot;
        if (tot) {
            std::vector<int> cnt((size_t)n, 0); std::vector<char> has_alt((size_t)n, 0);
            for (int i = 0; i < n; ++i) {
                int r = pri_idx(opt.XA_drop_ratio, a, i);
                if (r >= 0) { ++cnt[(size_t)r]; if (a[i].is_alt) has_alt[(size_t)r] = 1; }
            }
            xa.resize((size_t)n);
            for (int i = 0; i < n; ++i) {
                int r = pri_idx(opt.XA_drop_ratio, a, i);
                if (r < 0) continue;
                if (cnt[(size_t)r] > opt.max_XA_hits_alt || (!has_alt[(size_t)r] && cnt[(size_t)r] > opt.max_XA_hits)) continue;
                const b200_hit_t &t = a[i];
                Out &s = xa[(size_t)r];
                s.str(C.rnames[t.rid]);
                s.ch(','); s.ch("+-"[t.is_rev ? 1 : 0]); s.num((long long)t.pos + 1);
                s.ch(',');
                const uint32_t *cg = C.v.cigar + t.cigar_off;
                for (int k = 0; k < t.n_cigar; ++k) { s.num(cg[k] >> 4); s.ch("MIDSHN"[cg[k] & 0xf]); }
                s.ch(','); s.num(t.NM);
                if (opt.flag & F_XB) { s.ch(','); s.num(t.score); s.ch(','); s.num(t.mapq); }
                s.ch(';');
            }
        }
    }
    Aln small[8];
    std::vector<Aln> big;
    Aln *aa = small;
    if (n > 8) { big.resize((size_t)n); aa = big.data(); }
    size_t n_aa = 0;
    int l = 0;
    for (int k = 0; k < n; ++k) {
        const b200_hit_t &p = a[k];
        if (p.score < opt.T) continue;
        if (p.secondary >= 0 && (p.is_alt || !(opt.flag & F_ALL))) continue;
        if (p.secondary >= 0 && p.secondary < INT_MAX && p.score < a[p.secondary].score * opt.drop_ratio) continue;
        Aln q; q.h = &p; q.flag = p.flag; q.mapq = p.mapq; q.sub = p.aln_sub;
        q.xa = xa.empty() ? nullptr : &xa[(size_t)k];
        if (p.secondary >= 0) q.sub = -1;
        if (l && p.secondary < 0) q.flag |= (opt.flag & F_NO_MULTI) ? 0x10000 : 0x800;
        if (!(opt.flag & F_KEEP_SUPP_MAPQ) && l && !p.is_alt && q.mapq > aa[0].mapq) q.mapq = aa[0].mapq;
        aa[n_aa++] = q;
        ++l;
    }
    if (n_aa == 0) {
        Aln t; t.h = nullptr; t.flag = 0; t.mapq = 0; t.sub = 0; t.xa = nullptr;
        aa[0] = t;
        aln2sam(C, read, aa, 1, 0, o);
    } else {
        for (size_t k = 0; k < n_aa; ++k) aln2sam(C, read, aa, n_aa, (int)k, o);
    }
}

} // namespace

extern "C" int b200_results_to_sam(const b200_results_view_t *view, const b200_mem_opt_t *opt, const char *const *rnames, int n_rnames,
                                   const char *seqs, const int64_t *seq_off, const char *quals, const int64_t *qual_off,
                                   const char *names, const int64_t *name_off, const char *comments, const int64_t *comment_off,
                                   char **sam, int64_t *sam_len)
{
    if (!view || !opt || !rnames || !seqs || !seq_off || !names || !name_off || !sam || !sam_len || (quals && !qual_off) || (comments && !comment_off)) {
        b200::set_error("b200_results_to_sam: null argument"); return B200_ERR_ARG;
    }
    if (opt->flag & F_REF_HDR) { b200::set_error("b200_results_to_sam: MEM_F_REF_HDR (XR tag) is not supported: the index image keeps no contig annotations"); return B200_ERR_LIMIT; }
    Ctx C; C.rnames = rnames; C.n_rnames = n_rnames; C.opt = opt; C.v = *view;
    for (int64_t i = 0; i < view->n_hits; ++i)
        if (view->hits[i].rid >= n_rnames) { b200::set_error("b200_results_to_sam: a hit refers to a contig beyond rnames[]"); return B200_ERR_ARG; }
    C.seqs = seqs; C.seq_off = seq_off; C.quals = quals; C.qual_off = qual_off; C.names = names; C.name_off = name_off;
    C.comments = comments; C.comment_off = comment_off;
    const int64_t n = C.v.n_reads;
    unsigned nt = std::thread::hardware_concurrency();
    if (nt == 0) nt = 1;
    if (nt > 64) nt = 64;
    if ((int64_t)nt > (n + 4095) / 4096) nt = (unsigned)((n + 4095) / 4096);
    if (nt == 0) nt = 1;
    std::vector<Out> parts(nt);
    auto run = [&](const std::function<void(unsigned)> &f) {
        if (nt == 1) { f(0); return; }
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nt; ++t) th.emplace_back(f, t);
        for (auto &x : th) x.join();
    };
    run([&](unsigned t) {
        int64_t b = n * t / nt, e = n * (t + 1) / nt;
        int64_t bytes = (C.seq_off[e] - C.seq_off[b]) * 2 + (C.name_off[e] - C.name_off[b]) + (e - b) * 96;
        parts[t].reserve((size_t)bytes);
        for (int64_t i = b; i < e; ++i) read2sam(C, i, parts[t]);
    });
    size_t tot = 0;
    std::vector<size_t> at(nt);
    for (unsigned t = 0; t < nt; ++t) { at[t] = tot; tot += parts[t].n; }
    char *out = (char *)malloc(tot + 1);
    if (!out) { b200::set_error("b200_results_to_sam: out of memory"); return B200_ERR_NOMEM; }
    run([&](unsigned t) { if (parts[t].n) memcpy(out + at[t], parts[t].p, parts[t].n); });
    out[tot] = 0;
    *sam = out; *sam_len = (int64_t)tot;
    return 0;
}
