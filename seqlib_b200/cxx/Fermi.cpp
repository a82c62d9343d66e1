// Fermi.cpp -- SeqLib::FermiAssembler and SeqLib::BFC over the CUDA engine's C ABI (host code only).
// Follows src/FermiAssembler.cpp:7-205 and src/BFC.cpp:40-362 of the reference: same ownership of the fseq1_t array
// (strdup'd C strings, mutated in place by correction), same option handling, same quirks where they are observable.
#include <cstring>
#include <cassert>
#include <algorithm>
#include <stdexcept>
#include "SeqLib/FermiAssembler.h"
#include "SeqLib/BFC.h"

namespace SeqLib {

static void check(int rc, const char *what)
{
    if (rc != B200_OK) throw std::runtime_error(std::string(what) + ": " + b200_last_error());
}

FermiAssembler::FermiAssembler() : m_seqs(0), m(0), size(0), n_seqs(0), n_utg(0), m_utgs(0) { b200_fml_opt_init(&opt); }

FermiAssembler::FermiAssembler(fml_opt_t &_opt) : m_seqs(0), m(0), size(0), n_seqs(0), n_utg(0), opt(_opt), m_utgs(0) {}

FermiAssembler::~FermiAssembler()
{
    ClearReads();
    ClearContigs();
}

void FermiAssembler::push(const std::string &name, const std::string &seq, const std::string &qual, bool keep_empty_qual)
{
    if (m <= n_seqs) {
        m = m <= 0 ? 32 : (m * 2);
        m_seqs = (fseq1_t *)realloc(m_seqs, m * sizeof(fseq1_t));
        if (!m_seqs) throw std::bad_alloc();
    }
    m_names.push_back(name);
    fseq1_t *s = &m_seqs[n_seqs];
    s->seq = strdup(seq.c_str());
    // The reference's AddReads() strdup()s an empty quality string to "" and fermi-lite then reads l_seq bytes of it
    // (src/FermiAssembler.cpp:69-109, SURVEY 8b Q8: undefined behaviour).  A quality string that does not cover the read is
    // passed on as "no qualities" here.
    if (qual.size() == seq.size() && !qual.empty()) s->qual = strdup(qual.c_str());
    else s->qual = nullptr;
    (void)keep_empty_qual;
    s->l_seq = (int32_t)seq.length();
    size += m_seqs[n_seqs++].l_seq;
}

void FermiAssembler::AddRead(const BamRecord &r) { AddRead(UnalignedSequence(r.Qname(), r.Sequence(), r.Qualities())); }

void FermiAssembler::AddRead(const UnalignedSequence &r)
{
    if (r.Seq.empty()) return;
    if (r.Name.empty()) return;
    push(r.Name, r.Seq, r.Qual, true);
}

void FermiAssembler::AddReads(const UnalignedSequenceVector &v)
{
    for (UnalignedSequenceVector::const_iterator r = v.begin(); r != v.end(); ++r) push(r->Name, r->Seq, r->Qual, false);
}

void FermiAssembler::AddReads(const BamRecordVector &brv)
{
    for (BamRecordVector::const_iterator r = brv.begin(); r != brv.end(); ++r) push(r->Qname(), r->Sequence(), r->Qualities(), false);
}

std::vector<std::vector<std::string> > FermiAssembler::AssembleWindows(const std::vector<UnalignedSequenceVector> &windows,
                                                                      const fml_opt_t *opt_, int n_threads)
{
    fml_opt_t o;
    if (opt_) o = *opt_; else b200_fml_opt_init(&o);
    std::vector<std::vector<std::string> > out(windows.size());
    // a window has qualities if any of its reads has a quality string covering the read (push() above); windows with and
    // without qualities go in two calls because the flat quality pool is all or nothing
    for (int pass = 0; pass < 2; ++pass) {
        std::vector<size_t> which;
        std::vector<char> seqs, quals;
        std::vector<int64_t> off(1, 0), win_off(1, 0);
        for (size_t w = 0; w < windows.size(); ++w) {
            bool has_q = false;
            for (size_t i = 0; i < windows[w].size(); ++i)
                if (!windows[w][i].Seq.empty() && windows[w][i].Qual.size() == windows[w][i].Seq.size()) has_q = true;
            if ((int)has_q != pass) continue;
            which.push_back(w);
            for (size_t i = 0; i < windows[w].size(); ++i) {
                const UnalignedSequence &r = windows[w][i];
                seqs.insert(seqs.end(), r.Seq.begin(), r.Seq.end());
                if (pass) {
                    if (r.Qual.size() == r.Seq.size()) quals.insert(quals.end(), r.Qual.begin(), r.Qual.end());
                    else quals.insert(quals.end(), r.Seq.size(), '~');
                }
                off.push_back((int64_t)seqs.size());
            }
            win_off.push_back((int64_t)off.size() - 1);
        }
        if (which.empty()) continue;
        seqs.push_back(0); quals.push_back(0);
        std::vector<b200_utgs_t *> h(which.size(), (b200_utgs_t *)0);
        int rc = b200_fml_assemble_windows(&o, (int64_t)which.size(), win_off.data(), seqs.data(), pass ? quals.data() : 0, off.data(), n_threads, h.data());
        for (size_t k = 0; k < which.size(); ++k) {
            if (!h[k]) continue;
            int n = 0; const b200_utg_t *u = 0;
            if (b200_utgs_view(h[k], &n, &u) == B200_OK) for (int i = 0; i < n; ++i) out[which[k]].push_back(std::string(u[i].seq));
            b200_utgs_free(h[k]);
        }
        check(rc, "FermiAssembler::AssembleWindows");
    }
    return out;
}

void FermiAssembler::ClearContigs()
{
    b200_fml_utg_destroy(n_utg, m_utgs);
    m_utgs = 0;
    n_utg = 0;
}

void FermiAssembler::ClearReads()
{
    if (!m_seqs) return;
    for (size_t i = 0; i < n_seqs; ++i) {
        fseq1_t *s = &m_seqs[i];
        if (s->qual) free(s->qual);
        s->qual = NULL;
        if (s->seq) free(s->seq);
        s->seq = NULL;
    }
    free(m_seqs);
    m_seqs = NULL;
    // (the reference leaves n_seqs / m untouched here; they are reset so the object can be refilled)
    n_seqs = 0; m = 0; size = 0;
    m_names.clear();
}

// fml_correct on options straight from the caller: with ec_k == 0 (the default) the reference's call changes nothing
//  and so does the engine's.
void FermiAssembler::CorrectReads()
{
    float kcov = 0;
    check(b200_fml_correct(&opt, (int)n_seqs, m_seqs, &kcov), "FermiAssembler::CorrectReads");
}

void FermiAssembler::CorrectAndFilterReads()
{
    float kcov = 0;
    check(b200_fml_fltuniq(&opt, (int)n_seqs, m_seqs, &kcov), "FermiAssembler::CorrectAndFilterReads");
}

void FermiAssembler::PerformAssembly()
{
    b200_fml_utg_destroy(n_utg, m_utgs);
    m_utgs = 0; n_utg = 0;
    check(b200_fml_assemble(&opt, (int)n_seqs, m_seqs, &n_utg, &m_utgs), "FermiAssembler::PerformAssembly");
}

void FermiAssembler::DirectAssemble(float kcov)
{
    // src/FermiAssembler.cpp:24-39: no fml_opt_adjust, min_ensr raised from kcov without the min_cnt / max_cnt clamps
    opt.mag_opt.min_ensr = opt.mag_opt.min_ensr > kcov * .1 ? opt.mag_opt.min_ensr : (int)(kcov * .1 + .499);
    opt.mag_opt.min_insr = opt.mag_opt.min_ensr - 1;
    std::vector<int64_t> off(n_seqs + 1, 0);
    for (size_t i = 0; i < n_seqs; ++i) off[i + 1] = off[i] + (m_seqs[i].l_seq > 0 ? m_seqs[i].l_seq : 0);
    std::vector<char> pool((size_t)off[n_seqs] + 1);
    for (size_t i = 0; i < n_seqs; ++i) if (m_seqs[i].l_seq > 0) memcpy(pool.data() + off[i], m_seqs[i].seq, m_seqs[i].l_seq);
    b200_utgs_t *U = 0;
    check(b200_fml_seqs2utg_flat(&opt, (int64_t)n_seqs, pool.data(), off.data(), &U), "FermiAssembler::DirectAssemble");
    int n = 0; const b200_utg_t *v = 0;
    b200_utgs_view(U, &n, &v);
    b200_fml_utg_destroy(n_utg, m_utgs);
    m_utgs = n ? (fml_utg_t *)calloc(n, sizeof(fml_utg_t)) : 0; n_utg = n;
    for (int i = 0; i < n; ++i) {
        m_utgs[i] = v[i];
        m_utgs[i].seq = strndup(v[i].seq, v[i].len);
        m_utgs[i].cov = strndup(v[i].cov, v[i].len);
        int no = v[i].n_ovlp[0] + v[i].n_ovlp[1];
        m_utgs[i].ovlp = (fml_ovlp_t *)calloc(no ? no : 1, sizeof(fml_ovlp_t));
        if (no) memcpy(m_utgs[i].ovlp, v[i].ovlp, no * sizeof(fml_ovlp_t));
    }
    b200_utgs_free(U);
}

std::vector<std::string> FermiAssembler::GetContigs() const
{
    std::vector<std::string> c;
    for (int i = 0; i < n_utg; ++i) c.push_back(std::string(m_utgs[i].seq));
    return c;
}

UnalignedSequenceVector FermiAssembler::GetSequences() const
{
    UnalignedSequenceVector r;
    for (size_t i = 0; i < n_seqs; ++i) {
        fseq1_t *s = &m_seqs[i];
        UnalignedSequence read;
        if (s->seq) read.Seq = (std::string(s->seq));
        read.Name = m_names[i];
        r.push_back(read);
    }
    return r;
}

void FermiAssembler::WriteGFA(std::ostream &out)
{
    out << "H\tVN:Z:1.0" << std::endl;
    for (int i = 0; i < n_utg; ++i) {
        const fml_utg_t *u = m_utgs + i;
        out << "S\t" << i << "\t";
        out << u->seq << "\tLN:i:" << u->len << "\tRC:i:" << u->nsr << "\tPD:Z:";
        out << u->cov << std::endl;
        for (int j = 0; j < u->n_ovlp[0] + u->n_ovlp[1]; ++j) {
            fml_ovlp_t *o = &u->ovlp[j];
            if ((uint32_t)i < o->id)
                out << "L\t" << i << "\t" << "+-"[!o->from] << "\t" << o->id << "\t" << "+-"[o->to] << "\t" << o->len << "M" << std::endl;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------- BFC
BFC::BFC() : m_idx(0), flt_uniq(0), kmer(0), kcov(0), q(20), l_pre(-1), ch(nullptr) { b200_fml_opt_init(&fml_opt); }

BFC::~BFC()
{
    ClearReads();
    if (ch) { b200_kmer_table_destroy(ch); ch = nullptr; }
}

bool BFC::AddSequence(std::string_view seq, std::string_view qual, std::string_view name)
{
    if (seq.empty() || (!qual.empty() && qual.size() != seq.size())) return false;
    fseq1_t s{};
    s.l_seq = (int32_t)seq.size();
    s.seq = strndup(seq.data(), seq.size());
    s.qual = qual.empty() ? nullptr : strndup(qual.data(), qual.size());
    m_seqs.push_back(s);
    m_names.emplace_back(name);
    return true;
}

bool BFC::GetSequence(std::string &s, std::string &qn)
{
    if (m_idx >= m_seqs.size()) return false;
    s = std::string(m_seqs.at(m_idx).seq);
    qn = m_names.at(m_idx);
    std::transform(s.begin(), s.end(), s.begin(), ::toupper);
    ++m_idx;
    return true;
}

void BFC::ClearReads()
{
    for (size_t i = 0; i < m_seqs.size(); ++i) { free(m_seqs[i].seq); free(m_seqs[i].qual); m_seqs[i].seq = m_seqs[i].qual = nullptr; }
    m_seqs.clear();
    m_names.clear();
    m_idx = 0;
}

namespace {
struct Flat { std::vector<int64_t> off; std::vector<char> seq, qual; bool has_qual; };
Flat flatten(const std::vector<fseq1_t> &v)
{
    Flat f; f.off.assign(v.size() + 1, 0); f.has_qual = false;
    for (size_t i = 0; i < v.size(); ++i) { f.off[i + 1] = f.off[i] + v[i].l_seq; if (v[i].qual) f.has_qual = true; }
    f.seq.resize((size_t)f.off[v.size()] + 1);
    if (f.has_qual) f.qual.assign((size_t)f.off[v.size()] + 1, '~');      // reads without qualities count as high quality
    for (size_t i = 0; i < v.size(); ++i) {
        memcpy(f.seq.data() + f.off[i], v[i].seq, v[i].l_seq);
        if (v[i].qual) memcpy(f.qual.data() + f.off[i], v[i].qual, v[i].l_seq);
    }
    return f;
}
}

void BFC::Train()
{
    if (ch) { b200_kmer_table_destroy(ch); ch = nullptr; }
    b200_fml_opt_init(&fml_opt);
    if (kmer <= 0) {
        b200_fml_opt_adjust(&fml_opt, (int)m_seqs.size(), m_seqs.data());
        kmer = fml_opt.ec_k;
    }
    uint64_t tot_len = 0;
    for (auto const &s : m_seqs) tot_len += s.l_seq;
    l_pre = (tot_len > 8) ? (int)std::min<uint64_t>(tot_len - 8, 20) : 0;
    Flat f = flatten(m_seqs);
    check(b200_fml_count((int64_t)m_seqs.size(), f.seq.data(), f.has_qual ? f.qual.data() : nullptr, f.off.data(), kmer, q, l_pre, &ch), "BFC::Train");
}

void BFC::ErrorCorrect()
{
    assert(kmer > 0);
    if (!ch) throw std::runtime_error("BFC::ErrorCorrect: Train() was not called");
    uint64_t hist[256], hist_high[64];
    int mode = -1;
    check(b200_kmer_table_hist(ch, hist, hist_high, &mode), "BFC::ErrorCorrect");
    uint64_t sum_k = 0, tot_k = 0;
    for (int i = fml_opt.min_cnt; i < 256; ++i) { sum_k += hist[i]; tot_k += i * hist[i]; }
    kcov = sum_k ? static_cast<float>(tot_k) / sum_k : 0.0f;
    int raw_min = static_cast<int>(.1 * kcov + 0.499f);
    int min_cov = std::clamp(raw_min, fml_opt.min_cnt, fml_opt.max_cnt);
    Flat f = flatten(m_seqs);
    std::vector<int32_t> len(m_seqs.size() + 1);
    check(b200_kmer_correct_flat(ch, min_cov, mode, flt_uniq, (int64_t)m_seqs.size(), f.seq.data(), f.has_qual ? f.qual.data() : nullptr,
                                 f.off.data(), len.data()), "BFC::ErrorCorrect");
    // what worker_ec does with the outcome (fermi-lite/bfc.c:492-509): corrected bases copied back; with flt_uniq set a read is
    // trimmed to len[i] (seq and qual NUL-terminated there) or, when nothing of it is left, freed and NULLed with l_seq = 0
    for (size_t i = 0; i < m_seqs.size(); ++i) {
        const int32_t keep = flt_uniq ? len[i] : m_seqs[i].l_seq;
        if (flt_uniq && keep <= 0) {
            free(m_seqs[i].seq); free(m_seqs[i].qual);
            m_seqs[i].seq = m_seqs[i].qual = nullptr; m_seqs[i].l_seq = 0;
            continue;
        }
        memcpy(m_seqs[i].seq, f.seq.data() + f.off[i], keep);
        if (m_seqs[i].qual) memcpy(m_seqs[i].qual, f.qual.data() + f.off[i], keep);
        if (keep < m_seqs[i].l_seq) { m_seqs[i].seq[keep] = 0; if (m_seqs[i].qual) m_seqs[i].qual[keep] = 0; m_seqs[i].l_seq = keep; }
    }
}

} // namespace SeqLib
