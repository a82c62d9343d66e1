// SeqLib::BWAWrapper -- the legacy single-class spelling still used by the reference's README
// (README.md:126-143), seqtools (src/seqtools/seqtools.cpp:184-204) and its only BWA test
// (seq_test/seq_test.cpp:793-915).  Thin facade over BWAIndex + BWAAligner; AlignSequence
// returns BamRecordVector with the primary hit first (the ordering that test asserts).
#pragma once
#include "SeqLib/BWAAligner.h"

namespace SeqLib {

class BWAWrapper {
public:
    BWAWrapper();
    void ConstructIndex(const UnalignedSequenceVector &v);
    bool LoadIndex(const std::string &file);
    bool WriteIndex(const std::string &index_name) const;
    void AlignSequence(const std::string &seq, const std::string &name, BamRecordVector &vec, bool hardclip,
                       double keep_sec_with_frac_of_primary_score, int max_secondary) const;
    void AlignSequence(const UnalignedSequence &us, BamRecordVector &vec, bool hardclip,
                       double keep_sec_with_frac_of_primary_score, int max_secondary) const;
    BamHeader HeaderFromIndex() const { return index_->HeaderFromIndex(); }
    std::string ChrIDToName(int id) const { return index_->ChrIDToName(id); }
    int NumSequences() const { return index_->NumSequences(); }
    bool IsEmpty() const { return index_->IsEmpty(); }
    void SetGapOpen(int v) { aligner_->SetGapOpen(v); }
    void SetGapExtension(int v) { aligner_->SetGapExtension(v); }
    void SetMismatchPenalty(int v) { aligner_->SetMismatchPenalty(v); }
    void SetZDropoff(int v) { aligner_->SetZDropoff(v); }
    void SetAScore(int v) { aligner_->SetAScore(v); }
    void Set3primeClippingPenalty(int v) { aligner_->Set3primeClippingPenalty(v); }
    void Set5primeClippingPenalty(int v) { aligner_->Set5primeClippingPenalty(v); }
    void SetBandwidth(int v) { aligner_->SetBandwidth(v); }
    void SetReseedTrigger(float v) { aligner_->SetReseedTrigger(v); }
    friend std::ostream &operator<<(std::ostream &out, const BWAWrapper &b);

private:
    BWAIndexPtr index_;
    std::shared_ptr<BWAAligner> aligner_;
};

} // namespace SeqLib
