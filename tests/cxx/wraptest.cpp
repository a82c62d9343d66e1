// wraptest.cpp -- TEST HARNESS: serialises the BamRecords the product's wrapper code (seqlib_b200/cxx/BWA.cpp
// emit_records, reached through the test seam SeqLib::detail::RecordsFromRegions) makes from one read's regions, in the
// format of oracle/oracle_wrap.cpp, so the Python test can compare the two byte for byte without a GPU.
#include <cstring>
#include <vector>
#include "SeqLib/BWAAligner.h"
#include "seqlib_b200.h"

template <class T> static void put(std::vector<uint8_t> &o, T v) { const uint8_t *p = (const uint8_t *)&v; o.insert(o.end(), p, p + sizeof(T)); }

extern "C" int64_t wraptest_records(const char *seq_, int l_seq, const char *name_, int n_regs, const b200_hit_t *regs,
                                    const uint32_t *cigar_pool, int hardclip, double keepSecFrac, int maxSecondary,
                                    uint8_t *out, int64_t cap, int *n_rec)
{
    int64_t hit_off[2] = {0, n_regs};
    b200_results_view_t v; memset(&v, 0, sizeof(v));
    v.n_reads = 1; v.hit_off = hit_off; v.hits = regs; v.cigar = cigar_pool; v.n_hits = n_regs;
    SeqLib::BamRecordPtrVector recs;
    try {
        SeqLib::detail::RecordsFromRegions(std::string(seq_, (size_t)l_seq), std::string(name_), v, 0, hardclip != 0, keepSecFrac, maxSecondary, recs);
    } catch (const std::out_of_range &) { if (n_rec) *n_rec = -1; return -1; }
    {   // the batch path's arena packing must make the same records, and a record must survive getting memory of its own
        SeqLib::BamRecordPtrVector recs2;
        SeqLib::detail::RecordsFromRegions(std::string(seq_, (size_t)l_seq), std::string(name_), v, 0, hardclip != 0, keepSecFrac, maxSecondary, recs2, true);
        if (recs2.size() != recs.size()) { if (n_rec) *n_rec = -2; return -2; }
        for (size_t i = 0; i < recs.size(); ++i) {
            const bam1_t *a = recs[i]->b.get(), *c = recs2[i]->b.get();
            if (memcmp(&a->core, &c->core, sizeof(a->core)) || a->l_data != c->l_data || memcmp(a->data, c->data, (size_t)a->l_data) ||
                !(c->mempolicy & BAM_USER_OWNS_DATA) || a->mempolicy) { if (n_rec) *n_rec = -2; return -2; }
            std::vector<uint8_t> before(c->data, c->data + c->l_data);
            recs2[i]->AddZTag("BC", "xyz");
            const bam1_t *d = recs2[i]->b.get();
            std::string got;
            if (d->mempolicy || d->l_data != (int)before.size() + 7 || memcmp(d->data, before.data(), before.size()) ||
                !recs2[i]->GetZTag("BC", got) || got != "xyz") { if (n_rec) *n_rec = -3; return -3; }
        }
    }
    std::vector<uint8_t> o;
    for (auto &r : recs) {
        const bam1_t *b = r->b.get();
        put<int32_t>(o, b->core.tid); put<int64_t>(o, b->core.pos); put<uint8_t>(o, b->core.qual); put<uint16_t>(o, b->core.flag);
        put<uint32_t>(o, b->core.n_cigar); put<uint16_t>(o, b->core.l_qname); put<int32_t>(o, b->core.l_qseq);
        put<int32_t>(o, b->core.mtid); put<int64_t>(o, b->core.mpos); put<int64_t>(o, b->core.isize);
        const int core_len = b->core.l_qname + ((int)b->core.n_cigar << 2) + ((b->core.l_qseq + 1) >> 1) + b->core.l_qseq;
        put<int32_t>(o, core_len);
        std::vector<uint8_t> d(b->data, b->data + b->l_data);
        uint8_t *q = d.data() + b->core.l_qname + (b->core.n_cigar << 2) + ((b->core.l_qseq + 1) >> 1);
        for (int k = 1; k < b->core.l_qseq; ++k) q[k] = 0;          // the reference leaves qual[1..] uninitialised
        o.insert(o.end(), d.begin(), d.end());                      // core part + the aux tags appended after it
    }
    if (n_rec) *n_rec = (int)recs.size();
    if ((int64_t)o.size() <= cap && out) memcpy(out, o.data(), o.size());
    return (int64_t)o.size();
}
