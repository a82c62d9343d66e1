// BamRecord.cpp -- minimal bam1_t-backed record (see include/SeqLib/BamRecord.h).
// Follows SeqLib/BamRecord.h:258-601 and src/BamRecord.cpp:99-106,646-664,960-970 for the members the
// alignment path uses; aux tags are appended exactly like htslib's bam_aux_append (tag[2], type, payload).
#include <cstring>
#include <sstream>
#include <stdexcept>
#include "SeqLib/BamRecord.h"
#include "SeqLib/BamHeader.h"

namespace SeqLib {

namespace {
struct Bam1Free {
    void operator()(bam1_t *p) const { if (p) { free(p->data); free(p); } }
};
void aux_append(bam1_t *b, const char tag[2], char type, int len, const uint8_t *data)
{
    int ori = b->l_data;
    size_t need = (size_t)b->l_data + 3 + len;
    if (b->m_data < need) {
        size_t m = need; --m; m |= m >> 1; m |= m >> 2; m |= m >> 4; m |= m >> 8; m |= m >> 16; ++m;   // kroundup32
        uint8_t *nd = (uint8_t *)realloc(b->data, m);
        if (!nd) throw std::bad_alloc();
        b->data = nd; b->m_data = (uint32_t)m;
    }
    b->l_data = (int)need;
    b->data[ori] = (uint8_t)tag[0]; b->data[ori + 1] = (uint8_t)tag[1]; b->data[ori + 2] = (uint8_t)type;
    memcpy(b->data + ori + 3, data, len);
}
// walks the aux block; returns pointer to the type byte of `tag` or NULL
const uint8_t *aux_find(const bam1_t *b, const char tag[2])
{
    const uint8_t *s = bam_get_aux(b), *e = b->data + b->l_data;
    while (s + 3 <= e) {
        const uint8_t *t = s + 2;
        if (s[0] == (uint8_t)tag[0] && s[1] == (uint8_t)tag[1]) return t;
        char ty = (char)*t; s = t + 1;
        switch (ty) {
        case 'A': case 'c': case 'C': s += 1; break;
        case 's': case 'S': s += 2; break;
        case 'i': case 'I': case 'f': s += 4; break;
        case 'd': s += 8; break;
        case 'Z': case 'H': while (s < e && *s) ++s; ++s; break;
        case 'B': {
            if (s + 5 > e) return nullptr;
            char sub = (char)s[0]; uint32_t n; memcpy(&n, s + 1, 4); s += 5;
            int w = (sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : 4;
            s += (size_t)w * n; break;
        }
        default: return nullptr;
        }
    }
    return nullptr;
}
} // namespace

CigarField::CigarField(char t, uint32_t len)
{
    int op = -1;
    for (int i = 0; BAM_CIGAR_STR[i]; ++i) if (BAM_CIGAR_STR[i] == t) op = i;
    if (op < 0) throw std::invalid_argument("Cigar type must be one of MIDNSHP=XB");
    data = len << BAM_CIGAR_SHIFT | (uint32_t)op;
}

std::ostream &operator<<(std::ostream &out, const CigarField &c) { out << c.Length() << c.Type(); return out; }
std::ostream &operator<<(std::ostream &out, const Cigar &c) { for (auto &f : c) out << f; return out; }
int Cigar::NumQueryConsumed() const { int n = 0; for (auto &f : m_data) if (f.ConsumesQuery()) n += f.Length(); return n; }
int Cigar::NumReferenceConsumed() const { int n = 0; for (auto &f : m_data) if (f.ConsumesReference()) n += f.Length(); return n; }
bool Cigar::operator==(const Cigar &c) const { return m_data == c.m_data; }

BamRecord::BamRecord()
{
    bam1_t *p = (bam1_t *)calloc(1, sizeof(bam1_t));
    if (!p) throw std::bad_alloc();
    b = std::shared_ptr<bam1_t>(p, Bam1Free());
}

std::string BamRecord::Qname() const { return b && b->data ? std::string(bam_get_qname(b)) : std::string(); }

std::string BamRecord::Sequence() const
{
    static const char tab[] = "=ACMGRSVTWYHKDBN";
    const uint8_t *p = bam_get_seq(b);
    std::string out((size_t)b->core.l_qseq, 'N');
    for (int32_t i = 0; i < b->core.l_qseq; ++i) out[i] = tab[bam_seqi(p, i)];
    return out;
}

std::string BamRecord::Qualities(int offset) const
{
    const uint8_t *p = bam_get_qual(b);
    if (!b->core.l_qseq || p[0] == 0xff) return std::string();
    std::string out((size_t)b->core.l_qseq, ' ');
    for (int32_t i = 0; i < b->core.l_qseq; ++i) out[i] = (char)(p[i] + offset);
    return out;
}

Cigar BamRecord::GetCigar() const
{
    Cigar cig;
    const uint8_t *raw = (const uint8_t *)bam_get_cigar(b);
    for (uint32_t k = 0; k < b->core.n_cigar; ++k) { uint32_t c; memcpy(&c, raw + 4 * k, 4); cig.add(CigarField(c)); }
    return cig;
}

std::string BamRecord::CigarString() const { std::ostringstream ss; ss << GetCigar(); return ss.str(); }

int32_t BamRecord::PositionEnd() const { return b ? (int32_t)b->core.pos + GetCigar().NumReferenceConsumed() : -1; }

// A record packed into a batch block (mempolicy says the struct and the data are not its own) becomes an ordinary malloc'ed bam1_t
// before anything may reallocate its data.
void BamRecord::Own()
{
    if (!b || !(b->mempolicy & (BAM_USER_OWNS_STRUCT | BAM_USER_OWNS_DATA))) return;
    bam1_t *n = (bam1_t *)calloc(1, sizeof(bam1_t));
    if (!n) throw std::bad_alloc();
    *n = *b;
    n->mempolicy = 0;
    n->data = nullptr;
    if (b->data) {
        n->m_data = b->m_data ? b->m_data : (uint32_t)b->l_data;
        n->data = (uint8_t *)malloc(n->m_data ? n->m_data : 1);
        if (!n->data) { free(n); throw std::bad_alloc(); }
        memcpy(n->data, b->data, (size_t)b->l_data);
    }
    b = std::shared_ptr<bam1_t>(n, Bam1Free());
}

void BamRecord::AddIntTag(const std::string &tag, int32_t val)
{
    if (tag.size() < 2) return;
    Own();
    aux_append(b.get(), tag.data(), 'i', 4, (const uint8_t *)&val);
}

void BamRecord::AddZTag(std::string tag, std::string val)
{
    if (tag.size() < 2 || val.empty()) return;
    Own();
    aux_append(b.get(), tag.data(), 'Z', (int)val.size() + 1, (const uint8_t *)val.c_str());
}

bool BamRecord::GetIntTag(const std::string &tag, int32_t &t) const
{
    if (tag.size() < 2 || !b || !b->data) return false;
    const uint8_t *p = aux_find(b.get(), tag.data());
    if (!p) return false;
    switch ((char)*p) {
    case 'c': t = (int8_t)p[1]; return true; case 'C': t = p[1]; return true;
    case 's': { int16_t v; memcpy(&v, p + 1, 2); t = v; return true; }
    case 'S': { uint16_t v; memcpy(&v, p + 1, 2); t = v; return true; }
    case 'i': { int32_t v; memcpy(&v, p + 1, 4); t = v; return true; }
    case 'I': { uint32_t v; memcpy(&v, p + 1, 4); t = (int32_t)v; return true; }
    default: return false;
    }
}

bool BamRecord::GetZTag(const std::string &tag, std::string &s) const
{
    if (tag.size() < 2 || !b || !b->data) return false;
    const uint8_t *p = aux_find(b.get(), tag.data());
    if (!p || (char)*p != 'Z') return false;
    s = std::string((const char *)p + 1);
    return true;
}

std::ostream &operator<<(std::ostream &out, const BamRecord &r)
{
    if (r.isEmpty()) { out << "empty read"; return out; }
    out << r.Qname() << "\t" << r.AlignmentFlag() << "\t" << (r.ChrID() + 1) << "\t" << (r.Position() + 1) << "\t"
        << r.MapQuality() << "\t" << r.CigarString() << "\t*\t0\t0\t" << r.Sequence() << "\t*";
    int32_t v;
    if (r.GetIntTag("NM", v)) out << "\tNM:i:" << v;
    if (r.GetIntTag("AS", v)) out << "\tAS:i:" << v;
    return out;
}

// ---- BamHeader ---------------------------------------------------------------------------
BamHeader::BamHeader(const HeaderSequenceVector &hsv) : seqs_(hsv)
{
    std::ostringstream ss;
    for (auto &s : hsv) ss << "@SQ\tSN:" << s.Name << "\tLN:" << s.Length << "\n";
    text_ = ss.str();
}

BamHeader::BamHeader(const std::string &text) : text_(text)
{
    std::istringstream in(text);
    std::string line;
    while (std::getline(in, line)) {
        if (line.compare(0, 3, "@SQ") != 0) continue;
        std::string name; uint32_t len = 0;
        std::istringstream ls(line);
        std::string tok;
        while (std::getline(ls, tok, '\t')) {
            if (tok.compare(0, 3, "SN:") == 0) name = tok.substr(3);
            else if (tok.compare(0, 3, "LN:") == 0) len = (uint32_t)std::strtoul(tok.c_str() + 3, nullptr, 10);
        }
        if (!name.empty()) seqs_.push_back(HeaderSequence(name, len));
    }
}

std::string BamHeader::IDtoName(int id) const
{
    if (id < 0) throw std::invalid_argument("BamHeader::IDtoName - ID must be >= 0");
    if (seqs_.empty()) throw std::out_of_range("BamHeader::IDtoName - Header is uninitialized");
    if (id >= (int)seqs_.size()) throw std::out_of_range("BamHeader::IDtoName - Requested ID is higher than number of sequences");
    return seqs_[id].Name;
}

int BamHeader::Name2ID(const std::string &name) const
{
    for (size_t i = 0; i < seqs_.size(); ++i) if (seqs_[i].Name == name) return (int)i;
    return -1;
}

} // namespace SeqLib
