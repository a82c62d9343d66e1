// SeqLib::BFC -- same public surface as the reference's SeqLib/BFC.h:27-76 (behaviour of src/BFC.cpp:40-362), with
//   fml_count + bfc_ch_hist -> b200_fml_count + b200_kmer_table_hist,  kmer_correct -> b200_kmer_correct_flat,
//   bfc_ch_destroy -> b200_kmer_table_destroy.
#pragma once
#include <string>
#include <string_view>
#include <vector>
#include "seqlib_b200.h"
#include "SeqLib/FermiAssembler.h"

namespace SeqLib {

class BFC {
public:
    BFC();
    ~BFC();
    BFC(const BFC &) = delete;
    BFC &operator=(const BFC &) = delete;

    void ErrorCorrect();
    void Train();
    bool AddSequence(std::string_view seq, std::string_view qual, std::string_view name);
    void SetKmer(int k) { kmer = k; }
    void ClearReads();
    float GetKCov() const { return kcov; }
    int GetKMer() const { return kmer; }
    int NumSequences() const { return (int)m_seqs.size(); }
    bool GetSequence(std::string &s, std::string &q);
    void ResetGetSequence() { m_idx = 0; }

private:
    size_t m_idx;
    std::vector<fseq1_t> m_seqs;
    fml_opt_t fml_opt;
    std::vector<std::string> m_names;
    int flt_uniq;
    int kmer;
    float kcov;
    int q, l_pre;                   // bfc_opt_t.q / l_pre (fermi-lite/bfc.c:18-37)
    b200_kmer_table_t *ch;
};

} // namespace SeqLib
