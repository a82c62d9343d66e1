// BWA.cpp -- host side of the drop-in classes BWAIndex, BWAAligner and the legacy BWAWrapper.
// Compute goes through the C ABI (include/seqlib_b200.h) into the CUDA engine; what stays here is the
// reference's own host-side record assembly, restated from src/BWAIndex.cpp:19-180,382-418 and
// src/BWAAligner.cpp:7-265 (argument checks, exception types, hit sorting/filtering quirks Q1-Q6 of
// SURVEY.md 8b, bam1_t packing, NA/NM/AS tags).
#include <algorithm>
#include <cstring>
#include <sstream>
#include <stdexcept>
#include <functional>
#include <memory>
#include <thread>
#include <vector>
#include <exception>
#include <algorithm>
#include "SeqLib/BWAIndex.h"
#include "SeqLib/BWAAligner.h"
#include "SeqLib/BWAWrapper.h"
#include "seqlib_b200.h"

namespace SeqLib {

// ------------------------------------------------------------------ BWAIndex
BWAIndex::~BWAIndex() { if (idx_) b200_index_destroy(idx_); }

void BWAIndex::LoadIndex(const std::string &prefix)
{
    b200_index *n = nullptr;
    if (b200_index_load(prefix.c_str(), &n) != B200_OK || !n) throw std::runtime_error("Failed to load BWA index");
    if (idx_) b200_index_destroy(idx_);
    idx_ = n;
}

BamHeader BWAIndex::HeaderFromIndex() const { return BamHeader(printSamHeader()); }

int BWAIndex::NumSequences() const { return idx_ ? b200_index_n_seqs(idx_) : 0; }

std::string BWAIndex::ChrIDToName(int id) const
{
    if (!idx_) throw std::runtime_error("Index has not be loaded / constructed");
    int n = b200_index_n_seqs(idx_);
    if (id < 0 || id >= n) {
        std::ostringstream ss;
        ss << "BWAIndex::ChrIDToName - id out of bounds of refs in index for id of " << id << " on IDX of size " << n;
        throw std::out_of_range(ss.str());
    }
    return std::string(b200_index_seq_name(idx_, id));
}

std::string BWAIndex::printSamHeader() const
{
    if (!idx_) return "";
    std::ostringstream out;
    int n = b200_index_n_seqs(idx_);
    for (int i = 0; i < n; ++i) out << "@SQ\tSN:" << b200_index_seq_name(idx_, i) << "\tLN:" << b200_index_seq_len(idx_, i) << "\n";
    return out.str();
}

void BWAIndex::ConstructIndex(const UnalignedSequenceVector &refs)
{
    if (refs.empty()) return;
    for (auto const &r : refs)
        if (r.Name.empty() || r.Seq.empty())
            throw std::invalid_argument("BWAIndex::Construct each reference must have non-empty Name and Seq");
    if (idx_) { b200_index_destroy(idx_); idx_ = nullptr; }
    std::vector<const char *> names, seqs;
    for (auto const &r : refs) { names.push_back(r.Name.c_str()); seqs.push_back(r.Seq.c_str()); }
    b200_index *n = nullptr;
    int rc = b200_index_construct((int)refs.size(), names.data(), seqs.data(), 1, &n);
    if (rc != B200_OK || !n) throw std::runtime_error(std::string("BWAIndex::Construct BWT construction failed: ") + b200_last_error());
    idx_ = n;
}

void BWAIndex::WriteIndex(const std::string &prefix) const
{
    if (!idx_) throw std::runtime_error("BWAIndex::writeIndex: no index loaded");
    int rc = b200_index_write(idx_, prefix.c_str());
    if (rc != B200_OK) throw std::runtime_error(std::string("BWAIndex::writeIndex: ") + b200_last_error());
}

std::ostream &operator<<(std::ostream &os, const BWAIndex &idx)
{
    if (!idx.idx_) os << "[BWAIndex] <no index loaded>";
    else os << "[BWAIndex] #seqs=" << b200_index_n_seqs(idx.idx_) << " pac_len=" << b200_index_l_pac(idx.idx_);
    return os;
}

// ------------------------------------------------------------------ BWAAligner
BWAAligner::BWAAligner(BWAIndexPtr idx) : index_(std::move(idx))
{
    b200_mem_opt_init(&opt_);
    opt_.flag |= 0x200;                               // MEM_F_SOFTCLIP (SeqLib/BWAAligner.h:17)
}

#define NONNEG(v, what) do { if ((v) < 0) throw std::invalid_argument(what); } while (0)
void BWAAligner::SetGapOpen(int v) { NONNEG(v, "SetGapOpen: gap_open must be >= 0"); opt_.o_ins = opt_.o_del = v; }
void BWAAligner::SetGapExtension(int v) { NONNEG(v, "SetGapExtension: gap_ext must be >= 0"); opt_.e_ins = opt_.e_del = v; }
void BWAAligner::SetMismatchPenalty(int v) { NONNEG(v, "SetMismatchPenalty: mismatch must be >= 0"); opt_.b = v; b200_fill_scmat(opt_.a, opt_.b, opt_.mat); }
void BWAAligner::SetZDropoff(int v) { NONNEG(v, "SetZDropoff: zdrop must be >= 0"); opt_.zdrop = v; }
void BWAAligner::SetAScore(int a)
{   // scales the penalties but, like the reference, leaves `mat` alone (src/BWAAligner.cpp:43-59)
    NONNEG(a, "SetAScore: a must be >= 0");
    opt_.a = a; opt_.b *= a; opt_.T *= a; opt_.o_ins *= a; opt_.o_del *= a; opt_.e_ins *= a; opt_.e_del *= a;
    opt_.zdrop *= a; opt_.pen_clip5 *= a; opt_.pen_clip3 *= a; opt_.pen_unpaired *= a;
}
void BWAAligner::Set3primeClippingPenalty(int v) { NONNEG(v, "Set3primeClippingPenalty: penalty must be >= 0"); opt_.pen_clip3 = v; }
void BWAAligner::Set5primeClippingPenalty(int v) { NONNEG(v, "Set5primeClippingPenalty: penalty must be >= 0"); opt_.pen_clip5 = v; }
void BWAAligner::SetBandwidth(int v) { NONNEG(v, "SetBandwidth: bandwidth must be >= 0"); opt_.w = v; }
void BWAAligner::SetReseedTrigger(float v) { if (v < 0.0f) throw std::invalid_argument("SetReseedTrigger: trigger must be >= 0"); opt_.split_factor = v; }

namespace {

struct HitRef { const b200_hit_t *h; };

// sort by descending MAPQ, then (rid, pos)  (src/BWAAligner.cpp:7-11)
bool aln_sort(const HitRef &a, const HitRef &b)
{
    if (a.h->mapq != b.h->mapq) return a.h->mapq > b.h->mapq;
    if (a.h->rid != b.h->rid) return a.h->rid < b.h->rid;
    return a.h->pos < b.h->pos;
}

// Turns the hits of one read into BamRecords (src/BWAAligner.cpp:111-248).
// Batch packing: the bam1_t structs and data blocks of many records in a few large blocks (one arena per packing thread).  A record
// points into its arena through an aliasing shared_ptr, so the arena lives as long as any of its records and a record costs no
// allocation of its own; bam1_t::mempolicy says so, and BamRecord::Own() gives a record private memory before its data could grow.
struct RecordArena {
    std::vector<void *> blocks;
    uint8_t *cur = nullptr; size_t left = 0;
    void *alloc(size_t n)
    {
        n = (n + 15) & ~(size_t)15;
        if (n > left) {
            size_t sz = std::max<size_t>(n, (size_t)1 << 20);
            void *p = std::malloc(sz);
            if (!p) throw std::bad_alloc();
            blocks.push_back(p); cur = (uint8_t *)p; left = sz;
        }
        void *r = cur; cur += n; left -= n;
        return r;
    }
    ~RecordArena() { for (void *p : blocks) std::free(p); }
};

void emit_records(const std::string &seq, const std::string &name, const b200_results_view_t &v, int64_t read, bool hardclip,
                  double keepSecFrac, int maxSecondary, bool primary_first, BamRecordPtrVector &out,
                  const std::shared_ptr<RecordArena> &arena = std::shared_ptr<RecordArena>())
{
    int64_t b0 = v.hit_off[read], b1 = v.hit_off[read + 1];
    int n_regs = (int)(b1 - b0);
    double primaryScore = 0;
    HitRef small[8];                                     // nearly every read has one or two regions: no heap traffic for those
    std::vector<HitRef> big;
    HitRef *hits = small;
    if (n_regs > 8) { big.resize((size_t)n_regs); hits = big.data(); }
    size_t n_hits = 0;
    for (int64_t i = b0; i < b1; ++i) {
        const b200_hit_t &r = v.hits[i];
        // Q1: tests the int `secondary` (-1 = primary is non-zero too)
        if (r.secondary && (keepSecFrac < 0.0 || keepSecFrac > 1.0)) continue;
        hits[n_hits++].h = &r;
    }
    if (n_hits > 1) std::sort(hits, hits + n_hits, aln_sort);
    if (primary_first && n_hits > 1)   // legacy BWAWrapper ordering: primaries ahead of secondaries, otherwise as sorted
        std::stable_partition(hits, hits + n_hits, [](const HitRef &x) { return !(x.h->flag & BAM_FSECONDARY); });
    out.reserve(out.size() + n_hits);
    for (size_t i = 0; i < n_hits; ++i) {
        const b200_hit_t &h = *hits[i].h;
        bool isSec = (h.flag & BAM_FSECONDARY);
        bool tooLow = isSec && (primaryScore * keepSecFrac > h.score);
        bool tooMany = isSec && (int(i) > maxSecondary);                 // Q2: rank, not a secondary counter
        if (tooLow || tooMany) continue;
        if (!isSec) primaryScore = h.score;                             // Q3
        std::shared_ptr<BamRecord> rec;
        if (arena) {
            bam1_t *nb = (bam1_t *)arena->alloc(sizeof(bam1_t));
            std::memset(nb, 0, sizeof(bam1_t));
            nb->mempolicy = BAM_USER_OWNS_STRUCT | BAM_USER_OWNS_DATA;
            rec = std::make_shared<BamRecord>(std::shared_ptr<bam1_t>(arena, nb));
        } else rec = std::make_shared<BamRecord>();
        bam1_t *b = rec->b.get();
        b->core.tid = h.rid; b->core.pos = h.pos; b->core.qual = (uint8_t)h.mapq; b->core.flag = (uint16_t)h.flag;
        b->core.n_cigar = (uint32_t)h.n_cigar; b->core.mtid = -1; b->core.mpos = -1; b->core.isize = 0;
        if (h.is_rev) b->core.flag |= BAM_FREVERSE;
        const uint32_t *cig = v.cigar + h.cigar_off;
        const char *clipped = seq.data();                               // the bases that go into the record (no copy)
        size_t clipped_len = seq.size();
        if (hardclip) {                                                  // Q6
            size_t tstart = 0, clen = 0;
            for (int c = 0; c < h.n_cigar; ++c) {
                uint32_t op = bam_cigar_op(cig[c]);
                if (c == 0 && op == BAM_CREF_SKIP) tstart = bam_cigar_oplen(cig[c]);
                else if (bam_cigar_type(op) & 1) clen += bam_cigar_oplen(cig[c]);
            }
            if (tstart > seq.size()) throw std::out_of_range("basic_string::substr");   // what seq.substr(tstart, clen) does
            clipped = seq.data() + tstart;
            clipped_len = std::min(clen, seq.size() - tstart);
        }
        b->core.l_qname = (uint16_t)(name.size() + 1);
        b->core.l_qseq = (int32_t)clipped_len;
        b->l_data = b->core.l_qname + (h.n_cigar << 2) + ((b->core.l_qseq + 1) >> 1) + b->core.l_qseq;
        {   // room for the three integer tags appended below (3 x 7 bytes), rounded like bam_aux_append's kroundup32 would
            // leave it after the third append: one allocation instead of four
            size_t m = (size_t)b->l_data + 21; --m; m |= m >> 1; m |= m >> 2; m |= m >> 4; m |= m >> 8; m |= m >> 16; ++m;
            b->data = arena ? (uint8_t *)arena->alloc(m) : (uint8_t *)std::malloc(m);
            if (!b->data) throw std::bad_alloc();
            b->m_data = (uint32_t)m;
        }
        std::memset(b->data, 0, b->l_data);
        std::memcpy(b->data, name.c_str(), name.size() + 1);
        uint32_t *dst = bam_get_cigar(b);
        uint32_t newOp = hardclip ? BAM_CHARD_CLIP : BAM_CSOFT_CLIP;
        for (int k = 0; k < h.n_cigar; ++k) {
            uint32_t c = cig[k];
            if ((c & BAM_CIGAR_MASK) == BAM_CREF_SKIP) c = (c & ~(uint32_t)BAM_CIGAR_MASK) | newOp;
            std::memcpy((uint8_t *)dst + 4 * k, &c, 4);
        }
        uint8_t *sb = bam_get_seq(b);
        int sl = (int)clipped_len;
        {   // 4-bit codes, two bases per byte, high nibble first.  Reverse strand: Q11 -- src/BWAAligner.cpp:213-218 maps A->T
            // and T->A but leaves C and G as they are; BWAAligner mirrors that, the legacy BWAWrapper facade writes the true
            // reverse complement (what seq_test/seq_test.cpp:904 asserts).
            static const struct Tabs {
                uint8_t fwd[256], rev_ref[256], rev_true[256];
                Tabs()
                {
                    for (int c = 0; c < 256; ++c) fwd[c] = rev_ref[c] = rev_true[c] = 15;
                    fwd['A'] = 1; fwd['C'] = 2; fwd['G'] = 4; fwd['T'] = 8;
                    rev_ref['A'] = 8; rev_ref['C'] = 2; rev_ref['G'] = 4; rev_ref['T'] = 1;
                    rev_true['A'] = 8; rev_true['C'] = 4; rev_true['G'] = 2; rev_true['T'] = 1;
                }
            } T;
            const uint8_t *tab = !h.is_rev ? T.fwd : (primary_first ? T.rev_true : T.rev_ref);
            const unsigned char *src = (const unsigned char *)clipped;
            int j = 0;
            if (!h.is_rev) {
                for (; j + 1 < sl; j += 2) sb[j >> 1] = (uint8_t)(tab[src[j]] << 4 | tab[src[j + 1]]);
                if (j < sl) sb[j >> 1] = (uint8_t)(tab[src[j]] << 4);
            } else {
                for (; j + 1 < sl; j += 2) sb[j >> 1] = (uint8_t)(tab[src[sl - 1 - j]] << 4 | tab[src[sl - 2 - j]]);
                if (j < sl) sb[j >> 1] = (uint8_t)(tab[src[sl - 1 - j]] << 4);
            }
        }
        // Q4: the reference sets qual[0] = 0xff and leaves the rest uninitialised; here the whole string is "absent"
        if (sl) std::memset(bam_get_qual(b), 0xff, sl);
        {   // the three integer tags, as bam_aux_append would write them; the room was allocated above (Q5: no XA tag is ever produced)
            const int32_t vals[3] = {n_regs, h.NM, h.score};
            static const char tags[3][3] = {"NA", "NM", "AS"};
            uint8_t *t = b->data + b->l_data;
            for (int k = 0; k < 3; ++k, t += 7) { t[0] = (uint8_t)tags[k][0]; t[1] = (uint8_t)tags[k][1]; t[2] = 'i'; std::memcpy(t + 3, &vals[k], 4); }
            b->l_data += 21;
        }
        out.push_back(rec);
    }
}

} // namespace

namespace detail {
// test seam: the record assembly of alignSequence on regions that are already computed (tests/cxx/wraptest.cpp feeds it the
// reference's own regions and compares the records with oracle/oracle_wrap.cpp)
void RecordsFromRegions(const std::string &seq, const std::string &name, const b200_results_view_t &v, int64_t read, bool hardclip,
                        double keepSecFrac, int maxSecondary, BamRecordPtrVector &out, bool batch_packing)
{
    // batch_packing: the arena packing of alignSequences instead of the per-record allocation of alignSequence
    if (batch_packing) emit_records(seq, name, v, read, hardclip, keepSecFrac, maxSecondary, false, out, std::make_shared<RecordArena>());
    else emit_records(seq, name, v, read, hardclip, keepSecFrac, maxSecondary, false, out);
}
} // namespace detail

void BWAAligner::alignSequence(const std::string &seq, const std::string &name, BamRecordPtrVector &out, bool hardclip,
                               double keepSecFrac, int maxSecondary) const
{
    if (index_->IsEmpty()) return;
    int64_t off[2] = {0, (int64_t)seq.size()};
    int64_t id = lrand48();                                            // mem_align1's tie-break draw (bwa/bwamem_extra.c:112)
    b200_results_t *res = nullptr;
    int rc = b200_mem_align_batch(index_->handle(), &opt_, 1, seq.data(), off, &id, &res);
    if (rc != B200_OK) throw std::runtime_error(std::string("BWAAligner::alignSequence: ") + b200_last_error());
    b200_results_view_t v;
    b200_results_view(res, &v);
    try { emit_records(seq, name, v, 0, hardclip, keepSecFrac, maxSecondary, false, out); }
    catch (...) { b200_results_free(res); throw; }
    b200_results_free(res);
}

void BWAAligner::alignSequence(const UnalignedSequence &us, BamRecordPtrVector &out, bool hardclip, double keepSecFrac,
                               int maxSecondary) const
{
    alignSequence(us.Seq, us.Name, out, hardclip, keepSecFrac, maxSecondary);
    if (!copyComment_) return;
    for (auto &rec : out) rec->AddZTag("BC", us.Com);
}

void BWAAligner::alignSequences(const UnalignedSequenceVector &reads, std::vector<BamRecordPtrVector> &out, bool hardclip,
                                double keepSecFrac, int maxSecondary) const
{
    const size_t n = reads.size();
    const unsigned nt = n >= 4096 ? std::max(1u, std::min(std::thread::hardware_concurrency(), 32u)) : 1u;
    auto on_threads = [&](const std::function<void(unsigned)> &f) {
        if (nt == 1) { f(0); return; }
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nt; ++t) th.emplace_back(f, t);
        for (auto &x : th) x.join();
    };
    if (out.size() >= 4096 && nt > 1) {      // records of a previous batch: millions of small frees, spread over the threads
        const size_t m = out.size();
        on_threads([&](unsigned t) { for (size_t i = m * t / nt, e = m * (t + 1) / nt; i < e; ++i) BamRecordPtrVector().swap(out[i]); });
    }
    out.clear();
    out.resize(n);
    if (index_->IsEmpty() || reads.empty()) return;
    // offsets, tie-break ids and the flattened bases live in page-locked memory of the library's pool: the call below then moves
    // them by DMA instead of through the driver's staging copies of pageable memory
    // (small batches keep to the heap: page-locking costs more than it saves there)
    struct HostBuf {
        void *p = nullptr; bool pinned;
        HostBuf(size_t bytes, bool pin) : pinned(pin)
        {
            if (!pin) { p = std::malloc(bytes ? bytes : 1); if (!p) throw std::bad_alloc(); }
            else if (b200_host_alloc(bytes ? bytes : 1, &p) != B200_OK) throw std::runtime_error(std::string("BWAAligner::alignSequences: ") + b200_last_error());
        }
        ~HostBuf() { if (pinned) b200_host_free(p); else std::free(p); }
        HostBuf(const HostBuf &) = delete; HostBuf &operator=(const HostBuf &) = delete;
    };
    const bool pin = n >= 4096;
    HostBuf off_b((n + 1) * sizeof(int64_t), pin), ids_b(n * sizeof(int64_t), pin);
    int64_t *off = (int64_t *)off_b.p, *ids = (int64_t *)ids_b.p;
    off[0] = 0;
    for (size_t i = 0; i < n; ++i) { off[i + 1] = off[i] + (int64_t)reads[i].Seq.size(); ids[i] = lrand48(); }
    HostBuf all_b((size_t)off[n] + 1, pin);
    char *all = (char *)all_b.p;
    on_threads([&](unsigned t) { for (size_t i = n * t / nt, e = n * (t + 1) / nt; i < e; ++i) std::memcpy(all + off[i], reads[i].Seq.data(), reads[i].Seq.size()); });
    b200_results_t *res = nullptr;
    int rc = b200_mem_align_batch(index_->handle(), &opt_, (int64_t)n, all, off, ids, &res);
    if (rc != B200_OK) throw std::runtime_error(std::string("BWAAligner::alignSequences: ") + b200_last_error());
    b200_results_view_t v;
    b200_results_view(res, &v);
    // bam1_t packing (src/BWAAligner.cpp:151-247) is independent per read: large batches are packed by all host threads
    std::vector<std::exception_ptr> errs(nt);
    auto work = [&](unsigned t) {
        try {
            auto arena = std::make_shared<RecordArena>();
            for (size_t i = n * t / nt, e = n * (t + 1) / nt; i < e; ++i)
                emit_records(reads[i].Seq, reads[i].Name, v, (int64_t)i, hardclip, keepSecFrac, maxSecondary, false, out[i], arena);
        } catch (...) { errs[t] = std::current_exception(); }
    };
    on_threads(work);
    b200_results_free(res);
    for (auto &e : errs) if (e) std::rethrow_exception(e);
}

// ------------------------------------------------------------------ BWAWrapper (legacy facade)
BWAWrapper::BWAWrapper() : index_(std::make_shared<BWAIndex>()), aligner_(std::make_shared<BWAAligner>(index_)) {}

void BWAWrapper::ConstructIndex(const UnalignedSequenceVector &v) { index_->ConstructIndex(v); }

bool BWAWrapper::LoadIndex(const std::string &file)
{
    try { index_->LoadIndex(file); } catch (const std::runtime_error &) { return false; }
    return true;
}

bool BWAWrapper::WriteIndex(const std::string &index_name) const
{
    if (index_->IsEmpty()) return false;
    try { index_->WriteIndex(index_name); } catch (const std::runtime_error &) { return false; }
    return true;
}

void BWAWrapper::AlignSequence(const std::string &seq, const std::string &name, BamRecordVector &vec, bool hardclip,
                               double keepSecFrac, int maxSecondary) const
{
    if (index_->IsEmpty()) return;
    int64_t off[2] = {0, (int64_t)seq.size()};
    int64_t id = lrand48();
    b200_results_t *res = nullptr;
    int rc = b200_mem_align_batch(index_->handle(), &aligner_->options(), 1, seq.data(), off, &id, &res);
    if (rc != B200_OK) throw std::runtime_error(std::string("BWAWrapper::AlignSequence: ") + b200_last_error());
    b200_results_view_t v;
    b200_results_view(res, &v);
    BamRecordPtrVector tmp;
    try { emit_records(seq, name, v, 0, hardclip, keepSecFrac, maxSecondary, true, tmp); }
    catch (...) { b200_results_free(res); throw; }
    b200_results_free(res);
    for (auto &p : tmp) vec.push_back(*p);
}

void BWAWrapper::AlignSequence(const UnalignedSequence &us, BamRecordVector &vec, bool hardclip, double keepSecFrac,
                               int maxSecondary) const
{
    AlignSequence(us.Seq, us.Name, vec, hardclip, keepSecFrac, maxSecondary);
}

std::ostream &operator<<(std::ostream &out, const BWAWrapper &b) { out << *b.index_; return out; }

} // namespace SeqLib
