// SeqLib::BamHeader -- the part of SeqLib/BamHeader.h:37-106 the aligner needs
// (BWAIndex::HeaderFromIndex builds one from "@SQ\tSN:..\tLN:..\n" text,
// src/BWAIndex.cpp:35-78).  htslib is not available here, so the header is kept
// as parsed @SQ records plus the verbatim text.
#pragma once
#include <string>
#include <vector>
#include <stdexcept>

namespace SeqLib {

struct HeaderSequence {
    HeaderSequence(const std::string &n, uint32_t l) : Name(n), Length(l) {}
    std::string Name;
    uint32_t Length;
};
typedef std::vector<HeaderSequence> HeaderSequenceVector;

class BamHeader {
public:
    BamHeader() {}
    explicit BamHeader(const HeaderSequenceVector &hsv);
    explicit BamHeader(const std::string &text);
    int NumSequences() const { return (int)seqs_.size(); }
    int GetSequenceLength(int id) const { return id >= 0 && id < (int)seqs_.size() ? (int)seqs_[id].Length : -1; }
    int GetSequenceLength(const std::string &id) const { int i = Name2ID(id); return i < 0 ? -1 : (int)seqs_[i].Length; }
    bool IsOpen() const { return !text_.empty() || !seqs_.empty(); }
    bool isEmpty() const { return !IsOpen(); }
    std::string AsString() const { return text_; }
    std::string IDtoName(int id) const;
    int Name2ID(const std::string &name) const;
    HeaderSequenceVector GetHeaderSequenceVector() const { return seqs_; }

private:
    HeaderSequenceVector seqs_;
    std::string text_;
};

} // namespace SeqLib
