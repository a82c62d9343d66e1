// SeqLib::FermiAssembler -- same public surface as the reference's SeqLib/FermiAssembler.h:25-123 (and the behaviour of
// src/FermiAssembler.cpp:7-205), with fermi-lite's C calls replaced by the CUDA engine's C ABI (include/seqlib_b200.h):
//   fml_opt_init -> b200_fml_opt_init, fml_correct -> b200_fml_correct, fml_fltuniq -> b200_fml_fltuniq,
//   fml_assemble -> b200_fml_assemble, fml_seq2fmi + fml_fmi2mag + fml_mag_clean + fml_mag2utg -> b200_fml_seqs2utg_flat,
//   fml_utg_destroy -> b200_fml_utg_destroy.
#pragma once
#include <string>
#include <vector>
#include <cstdlib>
#include <iostream>
#include <stdint.h>
#include "seqlib_b200.h"
#include "SeqLib/UnalignedSequence.h"
#include "SeqLib/BamRecord.h"

#ifndef FKIT_FML_H__          // fermi-lite/fml.h not included: its names, on the layout-identical ABI types
typedef b200_fseq1_t fseq1_t;
typedef b200_magopt_t magopt_t;
typedef b200_fml_opt_t fml_opt_t;
typedef b200_utg_ovlp_t fml_ovlp_t;
typedef b200_utg_t fml_utg_t;
#define MAG_F_AGGRESSIVE 0x20
#define MAG_F_POPOPEN    0x40
#define MAG_F_NO_SIMPL   0x80
#endif

namespace SeqLib {

class FermiAssembler {
public:
    FermiAssembler();
    FermiAssembler(fml_opt_t &_opt);
    ~FermiAssembler();
    FermiAssembler(const FermiAssembler &) = delete;
    FermiAssembler &operator=(const FermiAssembler &) = delete;

    /** (new) Many windows in one call: windows[w] holds the reads of one genomic window; element w of the result is what
     *  AddReads(windows[w]) + PerformAssembly() + GetContigs() returns for a FermiAssembler of its own (same options; NULL =
     *  fml_opt_init).  The windows are assembled concurrently on the device (b200_fml_assemble_windows). */
    static std::vector<std::vector<std::string> > AssembleWindows(const std::vector<UnalignedSequenceVector> &windows,
                                                                 const fml_opt_t *opt = 0, int n_threads = 0);

    void AddReads(const BamRecordVector &brv);
    void AddReads(const UnalignedSequenceVector &v);
    void AddRead(const UnalignedSequence &r);
    void AddRead(const BamRecord &r);
    void ClearReads();
    void ClearContigs();
    void CorrectReads();
    void CorrectAndFilterReads();
    UnalignedSequenceVector GetSequences() const;
    void PerformAssembly();
    std::vector<std::string> GetContigs() const;
    void DirectAssemble(float kcov);

    void SetMinOverlap(uint32_t m) { opt.min_asm_ovlp = m; }
    void SetAggressiveTrim() { opt.mag_opt.flag |= MAG_F_AGGRESSIVE; }
    void SetSimplifyBubble() { opt.mag_opt.flag &= ~MAG_F_NO_SIMPL; }
    void SetDropOverlapRatio(double ratio) { opt.mag_opt.min_dratio1 = ratio; }
    void SetKmerMinThreshold(int min) { opt.min_cnt = min; }
    void SetKmerMaxThreshold(int max) { opt.max_cnt = max; }
    uint32_t GetMinOverlap() const { return opt.min_asm_ovlp; }
    size_t NumSequences() const { return n_seqs; }
    void WriteGFA(std::ostream &out);

private:
    void push(const std::string &name, const std::string &seq, const std::string &qual, bool keep_empty_qual);
    fseq1_t *m_seqs;
    size_t m;
    std::vector<std::string> m_names;
    uint64_t size;
    size_t n_seqs;
    int n_utg;
    fml_opt_t opt;
    fml_utg_t *m_utgs;
};

} // namespace SeqLib
