// oracle_wrap.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
// CPU restatement of the host-side record assembly of BWAAligner::alignSequence, src/BWAAligner.cpp:111-248, on a
// minimal bam1_t: which regions become hits (Q1: the int `secondary` is tested as a boolean, :118-120), the
// std::sort by (mapq desc, rid, pos) (:7-11,133), the keepSecFrac / maxSecondary filters with their quirks (Q2: rank in
// the sorted list, not a secondary counter, :140; Q3: primaryScore follows the sorted order, :147-148), the hard-clip
// substring (:164-176), the N -> S/H cigar rewrite (:191-200), the 4-bit sequence with the reverse-strand table of
// :209-219 (A<->T swapped, C and G left alone), qual[0] = 0xff (:235) and the NA / NM / AS integer tags (:237-241; XA is
// never set by mem_reg2aln).  htslib is not available here: bam_aux_append is restated as "append tag, type 'i', 4 bytes".
// The input is one read's regions in mem_align1 order with the mem_aln_t fields mem_reg2aln produced for them (the
// committed golden vectors or the live oracle/_ref library provide those).
// Output: a flat serialisation of the emitted records that tests/cxx/wraptest.cpp also produces from the product's
// BamRecords, so the two can be compared byte for byte.  Parity: pinned by construction on the reference's source
// (no fixture of the reference covers this layer: seq_test.cpp:893-911 only checks ChrID / Position / Sequence).
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include "../include/seqlib_b200.h"

namespace {

struct Aln {            // the mem_aln_t fields the wrapper reads (bwa/bwamem.h:107-121)
    int64_t pos; int rid, flag, is_rev, mapq, NM, n_cigar, score; const uint32_t *cigar;
};

bool aln_sort(const Aln &a, const Aln &b)        // src/BWAAligner.cpp:7-11
{
    if (a.mapq != b.mapq) return a.mapq > b.mapq;
    if (a.rid != b.rid) return a.rid < b.rid;
    return a.pos < b.pos;
}

template <class T> void put(std::vector<uint8_t> &o, T v) { const uint8_t *p = (const uint8_t *)&v; o.insert(o.end(), p, p + sizeof(T)); }

} // namespace

// returns the number of bytes the serialisation needs (written when it fits `cap`); *n_rec = records emitted
extern "C" int64_t oracle_wrap_records(const char *seq_, int l_seq, const char *name_, int n_regs, const b200_hit_t *regs,
                                       const uint32_t *cigar_pool, int hardclip, double keepSecFrac, int maxSecondary,
                                       uint8_t *out, int64_t cap, int *n_rec)
{
    const std::string seq(seq_, (size_t)l_seq), name(name_);
    double primaryScore = 0;
    std::vector<Aln> hits;
    for (int i = 0; i < n_regs; ++i) {
        const b200_hit_t &r = regs[i];
        if (r.secondary && (keepSecFrac < 0.0 || keepSecFrac > 1.0)) continue;
        Aln a; a.pos = r.pos; a.rid = r.rid; a.flag = r.flag; a.is_rev = r.is_rev; a.mapq = r.mapq; a.NM = r.NM; a.n_cigar = r.n_cigar;
        a.score = r.score; a.cigar = cigar_pool + r.cigar_off;
        hits.push_back(a);
    }
    std::sort(hits.begin(), hits.end(), aln_sort);
    std::vector<uint8_t> o;
    int emitted = 0;
    for (size_t i = 0; i < hits.size(); ++i) {
        const Aln &h = hits[i];
        bool isSec = (h.flag & 256);
        bool tooLow = isSec && (primaryScore * keepSecFrac > h.score);
        bool tooMany = isSec && (int(i) > maxSecondary);
        if (tooLow || tooMany) continue;
        if (!isSec) primaryScore = h.score;
        int32_t tid = h.rid; int64_t pos = h.pos; uint8_t qual = (uint8_t)h.mapq; uint16_t flag = (uint16_t)h.flag;
        uint32_t n_cigar = (uint32_t)h.n_cigar;
        if (h.is_rev) flag |= 16;
        std::string clipped = seq;
        if (hardclip) {
            size_t tstart = 0, clen = 0;
            for (int c = 0; c < h.n_cigar; ++c) {
                uint32_t op = h.cigar[c] & 0xf;
                if (c == 0 && op == 3) tstart = h.cigar[c] >> 4;
                else if ((0x3C1A7 >> (op << 1) & 3) & 1) clen += h.cigar[c] >> 4;
            }
            clipped = seq.substr(tstart, clen);
        }
        uint16_t l_qname = (uint16_t)(name.size() + 1);
        int32_t l_qseq = (int32_t)clipped.size();
        int l_data = l_qname + (h.n_cigar << 2) + ((l_qseq + 1) >> 1) + l_qseq;
        std::vector<uint8_t> data((size_t)l_data, 0);
        memcpy(data.data(), name.c_str(), name.size() + 1);
        memcpy(data.data() + l_qname, h.cigar, (size_t)h.n_cigar << 2);
        uint32_t newOp = hardclip ? 5 : 4;
        for (uint32_t k = 0; k < n_cigar; ++k) {
            uint32_t c; memcpy(&c, data.data() + l_qname + 4 * k, 4);
            if ((c & 0xf) == 3) { c = (c & ~0xfu) | newOp; memcpy(data.data() + l_qname + 4 * k, &c, 4); }
        }
        uint8_t *seqbuf = data.data() + l_qname + (n_cigar << 2);
        int sl = (int)clipped.size();
        if (h.is_rev) {
            int j = 0;
            for (int p = sl - 1; p >= 0; --p, ++j) {
                uint8_t v = 15;
                switch (clipped[p]) { case 'A': v = 8; break; case 'C': v = 2; break; case 'G': v = 4; break; case 'T': v = 1; break; }
                seqbuf[j >> 1] &= ~(0xF << ((~j & 1) << 2));
                seqbuf[j >> 1] |= v << ((~j & 1) << 2);
            }
        } else {
            for (int p = 0; p < sl; ++p) {
                uint8_t v = 15;
                switch (clipped[p]) { case 'A': v = 1; break; case 'C': v = 2; break; case 'G': v = 4; break; case 'T': v = 8; break; }
                seqbuf[p >> 1] &= ~(0xF << ((~p & 1) << 2));
                seqbuf[p >> 1] |= v << ((~p & 1) << 2);
            }
        }
        uint8_t *q = seqbuf + ((l_qseq + 1) >> 1);
        if (l_qseq) q[0] = 0xff;
        // serialise: core, then data with qual[1..] masked (uninitialised in the reference), then the three tags
        put(o, tid); put(o, pos); put(o, qual); put(o, flag); put(o, n_cigar); put(o, l_qname); put(o, l_qseq);
        put<int32_t>(o, -1); put<int64_t>(o, -1); put<int64_t>(o, 0);
        for (int k = 1; k < l_qseq; ++k) q[k] = 0;
        put<int32_t>(o, l_data);
        o.insert(o.end(), data.begin(), data.end());
        const char *tags[3] = {"NA", "NM", "AS"};
        int32_t vals[3] = {n_regs, h.NM, h.score};
        for (int t = 0; t < 3; ++t) { o.push_back((uint8_t)tags[t][0]); o.push_back((uint8_t)tags[t][1]); o.push_back('i'); put(o, vals[t]); }
        ++emitted;
    }
    if (n_rec) *n_rec = emitted;
    if ((int64_t)o.size() <= cap && out) memcpy(out, o.data(), o.size());
    return (int64_t)o.size();
}
