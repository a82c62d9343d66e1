// SeqLib::BamRecord / Cigar / CigarField -- the output container of the aligner.
//
// htslib is not vendored by the reference and not installed here, so bam1_t is
// restated with htslib's public layout (htslib/sam.h, >= 1.10: 64-bit pos) so a
// real htslib can be substituted later; "bam1_t layout: parity unpinned" (the
// reference does not pin an htslib version), the values placed in it are pinned
// by bwa.  Covers what src/BWAAligner.cpp:136-248 writes and what the
// reference's BWA test reads back (seq_test/seq_test.cpp:893-911;
// SeqLib/BamRecord.h:49-186,258-601).
#pragma once
#include <cstdint>
#include <cstdlib>
#include <memory>
#include <string>
#include <vector>
#include <ostream>

extern "C" {
typedef int64_t hts_pos_t;
typedef struct bam1_core_t {
    hts_pos_t pos;
    int32_t tid;
    uint16_t bin;
    uint8_t qual;
    uint8_t l_extranul;
    uint16_t flag;
    uint16_t l_qname;
    uint32_t n_cigar;
    int32_t l_qseq;
    int32_t mtid;
    hts_pos_t mpos;
    hts_pos_t isize;
} bam1_core_t;
typedef struct bam1_t {
    bam1_core_t core;
    uint64_t id;
    uint8_t *data;
    int l_data;
    uint32_t m_data;
    uint32_t mempolicy : 2, : 30;
} bam1_t;
}

#define BAM_USER_OWNS_STRUCT 1
#define BAM_USER_OWNS_DATA 2
#define BAM_FPAIRED 1
#define BAM_FUNMAP 4
#define BAM_FREVERSE 16
#define BAM_FSECONDARY 256
#define BAM_FSUPPLEMENTARY 2048
#define BAM_CMATCH 0
#define BAM_CINS 1
#define BAM_CDEL 2
#define BAM_CREF_SKIP 3
#define BAM_CSOFT_CLIP 4
#define BAM_CHARD_CLIP 5
#define BAM_CIGAR_MASK 0xf
#define BAM_CIGAR_SHIFT 4
#define BAM_CIGAR_STR "MIDNSHP=XB"
#define bam_cigar_op(c) ((c) & BAM_CIGAR_MASK)
#define bam_cigar_oplen(c) ((c) >> BAM_CIGAR_SHIFT)
#define bam_cigar_type(o) (0x3C1A7 >> ((o) << 1) & 3)   // bit 1: consumes query; bit 2: consumes reference
#define bam_get_qname(b) ((char *)(b)->data)
#define bam_get_cigar(b) ((uint32_t *)((b)->data + (b)->core.l_qname))
#define bam_get_seq(b) ((b)->data + ((b)->core.n_cigar << 2) + (b)->core.l_qname)
#define bam_get_qual(b) ((b)->data + ((b)->core.n_cigar << 2) + (b)->core.l_qname + (((b)->core.l_qseq + 1) >> 1))
#define bam_get_aux(b) ((b)->data + ((b)->core.n_cigar << 2) + (b)->core.l_qname + (((b)->core.l_qseq + 1) >> 1) + (b)->core.l_qseq)
#define bam_get_l_aux(b) ((b)->l_data - ((b)->core.n_cigar << 2) - (b)->core.l_qname - (b)->core.l_qseq - (((b)->core.l_qseq + 1) >> 1))
#define bam_seqi(s, i) ((s)[(i) >> 1] >> ((~(i) & 1) << 2) & 0xf)

namespace SeqLib {

class CigarField {
public:
    CigarField(char t, uint32_t len);
    explicit CigarField(uint32_t f) : data(f) {}
    uint32_t raw() const { return data; }
    char Type() const { return BAM_CIGAR_STR[data & BAM_CIGAR_MASK]; }
    uint8_t RawType() const { return data & BAM_CIGAR_MASK; }
    uint32_t Length() const { return data >> BAM_CIGAR_SHIFT; }
    bool ConsumesReference() const { return bam_cigar_type(data & BAM_CIGAR_MASK) & 2; }
    bool ConsumesQuery() const { return bam_cigar_type(data & BAM_CIGAR_MASK) & 1; }
    friend std::ostream &operator<<(std::ostream &out, const CigarField &c);
    bool operator==(const CigarField &c) const { return c.data == data; }
    bool operator!=(const CigarField &c) const { return c.data != data; }

private:
    uint32_t data;
};

class Cigar {
public:
    Cigar() {}
    typedef std::vector<CigarField>::iterator iterator;
    typedef std::vector<CigarField>::const_iterator const_iterator;
    iterator begin() { return m_data.begin(); }
    iterator end() { return m_data.end(); }
    const_iterator begin() const { return m_data.begin(); }
    const_iterator end() const { return m_data.end(); }
    const CigarField &back() const { return m_data.back(); }
    const CigarField &front() const { return m_data.front(); }
    size_t size() const { return m_data.size(); }
    const CigarField &operator[](size_t i) const { return m_data[i]; }
    CigarField &operator[](size_t i) { return m_data[i]; }
    void add(const CigarField &c) { m_data.push_back(c); }
    int NumQueryConsumed() const;
    int NumReferenceConsumed() const;
    bool operator==(const Cigar &c) const;
    friend std::ostream &operator<<(std::ostream &out, const Cigar &c);

private:
    std::vector<CigarField> m_data;
};

class BamRecord;
typedef std::shared_ptr<BamRecord> BamRecordPtr;
typedef std::vector<BamRecordPtr> BamRecordPtrVector;
typedef std::vector<BamRecord> BamRecordVector;

class BamRecord {
    friend class BWAAligner;
    friend class BWAWrapper;

public:
    BamRecord();                                   ///< allocates an empty bam1_t (SeqLib/BamRecord.cpp:99-106)
    /// wraps a record somebody else owns (htslib's BAM_USER_OWNS_STRUCT / BAM_USER_OWNS_DATA in bam1_t::mempolicy): the batch path
    /// of BWAAligner packs the records of a batch into a few large blocks that live as long as any of their records
    explicit BamRecord(std::shared_ptr<bam1_t> p) : b(std::move(p)) {}
    bool isEmpty() const { return !b || b->data == nullptr; }
    std::string Qname() const;
    int32_t ChrID() const { return b ? b->core.tid : -1; }
    int32_t Position() const { return b ? (int32_t)b->core.pos : -1; }
    int32_t MapQuality() const { return b ? b->core.qual : -1; }
    uint32_t AlignmentFlag() const { return b->core.flag; }
    bool ReverseFlag() const { return b && (b->core.flag & BAM_FREVERSE); }
    bool SecondaryFlag() const { return b && (b->core.flag & BAM_FSECONDARY); }
    bool MappedFlag() const { return b && !(b->core.flag & BAM_FUNMAP); }
    int32_t Length() const { return b->core.l_qseq; }
    std::string Sequence() const;
    std::string Qualities(int offset = 33) const;
    Cigar GetCigar() const;
    std::string CigarString() const;
    size_t CigarSize() const { return b->core.n_cigar; }
    int32_t PositionEnd() const;
    void AddIntTag(const std::string &tag, int32_t val);
    void AddZTag(std::string tag, std::string val);
    bool GetIntTag(const std::string &tag, int32_t &t) const;
    bool GetZTag(const std::string &tag, std::string &s) const;
    const bam1_t *raw() const { return b.get(); }
    void Own();                                    ///< makes the record own its memory (a private copy) if it does not; the tag setters call it
    friend std::ostream &operator<<(std::ostream &out, const BamRecord &r);

    std::shared_ptr<bam1_t> b;                     ///< the record (public in the reference too, SeqLib/BamRecord.h:248)
};

} // namespace SeqLib
