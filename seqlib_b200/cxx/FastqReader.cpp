// SeqLib::FastqReader over the C ABI (see include/SeqLib/FastqReader.h).
#include <iostream>
#include "SeqLib/FastqReader.h"

namespace SeqLib {

static const int64_t kBatch = 4096;

FastqReader::FastqReader(const std::string &file) : m_file(file), m_r(nullptr), m_i(0), m_done(false)
{
    m_b.n = 0;
    Open(m_file);
}

FastqReader::~FastqReader() { if (m_r) b200_fastq_close(m_r); }

bool FastqReader::Open(const std::string &f)
{
    m_file = f;
    if (m_r) { b200_fastq_close(m_r); m_r = nullptr; }
    m_b.n = 0; m_i = 0; m_done = false;
    if (b200_fastq_open(f.c_str(), &m_r) != 0) {
        std::cerr << b200_last_error() << std::endl;         // the reference prints and returns false (src/FastqReader.cpp:14-26)
        m_r = nullptr;
        return false;
    }
    return true;
}

bool FastqReader::fill()
{
    if (!m_r || m_done) return false;
    if (b200_fastq_next_batch(m_r, kBatch, &m_b) != 0) { m_b.n = 0; m_done = true; return false; }
    m_i = 0;
    // a negative kseq_read ends GetNextSequence's stream (src/FastqReader.cpp:45-46)
    if (m_b.status != 0) m_done = true;
    return m_b.n > 0;
}

bool FastqReader::GetNextSequence(UnalignedSequence &s)
{
    if (!m_r) return false;
    if (m_i >= m_b.n && !fill()) return false;
    const int64_t i = m_i++;
    s.Name.assign(m_b.name + m_b.name_off[i], (size_t)(m_b.name_off[i + 1] - m_b.name_off[i]));
    // Com / Qual are only assigned once kseq has allocated those strings (src/FastqReader.cpp:49-56): a comment-free FASTA
    // leaves the caller's Com and Qual untouched
    if (m_b.has[i] & 1) s.Com.assign(m_b.comment + m_b.comment_off[i], (size_t)(m_b.comment_off[i + 1] - m_b.comment_off[i]));
    if (m_b.has[i] & 2) s.Qual.assign(m_b.qual + m_b.qual_off[i], (size_t)(m_b.qual_off[i + 1] - m_b.qual_off[i]));
    s.Seq.assign(m_b.seq + m_b.seq_off[i], (size_t)(m_b.seq_off[i + 1] - m_b.seq_off[i]));
    return true;
}

size_t FastqReader::GetNextSequences(UnalignedSequenceVector &v, size_t max)
{
    size_t k = 0;
    UnalignedSequence s;
    while (k < max && GetNextSequence(s)) { v.push_back(s); ++k; }
    return k;
}

} // namespace SeqLib
