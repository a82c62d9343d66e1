// SeqLib::FastqReader drop-in: mirrors how the reference's README / seq_test use the class (Open, GetNextSequence loop)
// and prints one line per record for the Python test to compare with the reference parser's output.
#include <cstdio>
#include <iostream>
#include "SeqLib/FastqReader.h"

int main(int argc, char **argv)
{
    if (argc < 2) return 2;
    SeqLib::FastqReader none;
    SeqLib::UnalignedSequence s;
    if (none.GetNextSequence(s)) return 3;                     // an empty reader yields nothing
    SeqLib::FastqReader bad("/nonexistent/path.fq");           // prints a message, does not throw
    if (bad.GetNextSequence(s)) return 4;
    SeqLib::FastqReader r;
    if (!r.Open(argv[1])) return 5;
    s.Com = "KEEP"; s.Qual = "KEEPQ";
    size_t n = 0;
    while (r.GetNextSequence(s)) {
        std::printf("%s\x01%s\x01%s\x01%s\n", s.Name.c_str(), s.Com.c_str(), s.Seq.c_str(), s.Qual.c_str());
        ++n;
    }
    std::fprintf(stderr, "%zu records\n", n);
    return 0;
}
