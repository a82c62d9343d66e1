// SeqLib::BWAAligner -- drop-in for SeqLib/BWAAligner.h:12-69.  alignSequence keeps the
// reference's signature and per-hit post-processing (src/BWAAligner.cpp:89-250); the
// seed-and-extend itself runs on the GPU through b200_mem_align_batch.  alignSequences
// is the batched entry the GPU wants (the reference's precedent is mem_process_seqs,
// bwa/bwamem.h:161).
#pragma once
#include <vector>
#include "SeqLib/BWAIndex.h"
#include "SeqLib/BamRecord.h"
#include "SeqLib/UnalignedSequence.h"
#include "seqlib_b200.h"

namespace SeqLib {

class BWAAligner {
public:
    explicit BWAAligner(BWAIndexPtr idx);
    ~BWAAligner() {}

    void SetGapOpen(int gap_open);
    void SetGapExtension(int gap_ext);
    void SetMismatchPenalty(int mismatch);
    void SetZDropoff(int zdrop);
    void SetAScore(int a);
    void Set3primeClippingPenalty(int penalty);
    void Set5primeClippingPenalty(int penalty);
    void SetBandwidth(int bw);
    void SetReseedTrigger(float trigger);

    void alignSequence(const std::string &seq, const std::string &name, BamRecordPtrVector &out, bool hardclip,
                       double keepSecFrac, int maxSecondary) const;
    void alignSequence(const UnalignedSequence &us, BamRecordPtrVector &out, bool hardclip, double keepSecFrac,
                       int maxSecondary) const;
    /// Batch form: out[i] receives the records of reads[i]; one lrand48() tie-break id is drawn per read, in order,
    /// exactly as a loop over alignSequence would (bwa/bwamem_extra.c:112).
    void alignSequences(const UnalignedSequenceVector &reads, std::vector<BamRecordPtrVector> &out, bool hardclip,
                        double keepSecFrac, int maxSecondary) const;

    const b200_mem_opt_t &options() const { return opt_; }

private:
    BWAIndexPtr index_;
    b200_mem_opt_t opt_;
    bool copyComment_ = false;
};

namespace detail {
// Test seam (not part of the reference's API): the record assembly of BWAAligner::alignSequence, src/BWAAligner.cpp:111-248,
// applied to regions that are already computed.  tests/test_cpu_wrapper.py checks it against oracle/oracle_wrap.cpp.
void RecordsFromRegions(const std::string &seq, const std::string &name, const b200_results_view_t &v, int64_t read, bool hardclip,
                        double keepSecFrac, int maxSecondary, BamRecordPtrVector &out, bool batch_packing = false);
}

} // namespace SeqLib
