// SeqLib::FastqReader -- same public surface as the reference's SeqLib/FastqReader.h:22-60 (behaviour of
// src/FastqReader.cpp:8-59: kseq_read over gzread -> b200_fastq_open / b200_fastq_next_batch), plus a batch form that
// fills an UnalignedSequenceVector for BWAAligner::alignSequences.
#pragma once
#include <string>
#include "seqlib_b200.h"
#include "SeqLib/UnalignedSequence.h"

namespace SeqLib {

class FastqReader {
public:
    /** Construct an empty FASTQ/FASTA reader */
    FastqReader() : m_r(nullptr), m_i(0), m_done(false) { m_b.n = 0; }
    /** Construct a reader and open a FASTQ/FASTA file */
    FastqReader(const std::string &file);
    FastqReader(const FastqReader &) = delete;
    FastqReader &operator=(const FastqReader &) = delete;
    ~FastqReader();

    /** Open a FASTQ/FASTA file ("-" = stdin); false (and a message on stderr) if it cannot be read */
    bool Open(const std::string &file);
    /** Retrieve the next sequence: Name, Com, Seq, Qual.  false at the end of the input or at a truncated record */
    bool GetNextSequence(UnalignedSequence &s);
    /** Batch form: appends up to `max` records to `v`, returns the number appended */
    size_t GetNextSequences(UnalignedSequenceVector &v, size_t max);

private:
    bool fill();
    std::string m_file;
    b200_fastq_t *m_r;
    b200_fastq_batch_t m_b;
    int64_t m_i;
    bool m_done;
};

} // namespace SeqLib
