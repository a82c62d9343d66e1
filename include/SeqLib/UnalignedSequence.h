// SeqLib::UnalignedSequence -- input record of the aligner / assembler
// (same public members as the reference's SeqLib/UnalignedSequence.h:9-60).
#pragma once
#include <string>
#include <vector>
#include <ostream>

namespace SeqLib {

struct UnalignedSequence {
    UnalignedSequence() : Strand('*') {}
    UnalignedSequence(const std::string &n, const std::string &s) : Name(n), Seq(s), Strand('*') {}
    UnalignedSequence(const std::string &n, const std::string &s, const std::string &q) : Name(n), Seq(s), Qual(q), Strand('*') {}
    UnalignedSequence(const std::string &n, const std::string &s, const std::string &q, char t) : Name(n), Seq(s), Qual(q), Strand(t) {}

    std::string Name;  ///< read / contig name
    std::string Com;   ///< comment
    std::string Seq;   ///< bases (ACGTN)
    std::string Qual;  ///< quality string
    char Strand;       ///< '*', '+' or '-'

    friend std::ostream &operator<<(std::ostream &os, const UnalignedSequence &us)
    {
        os << "@" << us.Name << " " << us.Com << "\n" << us.Seq << "\n+\n" << us.Qual << "\n";
        return os;
    }
};

typedef std::vector<UnalignedSequence> UnalignedSequenceVector;

} // namespace SeqLib
