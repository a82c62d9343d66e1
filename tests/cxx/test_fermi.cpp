// C++ drop-in test of SeqLib::FermiAssembler / SeqLib::BFC: the flows of the reference's Boost tests "fermi_assemble",
// "fermi_add_reads" and "correct_and_assemble" (seq_test/seq_test.cpp:111-175,374-389,468-500) on reads given as a
// "name<TAB>seq<TAB>qual" text file; prints the contigs so the Python test can compare them with the golden unitigs.  Needs a GPU.
#include <cstdio>
#include <fstream>
#include <iostream>
#include <sstream>
#include "SeqLib/FermiAssembler.h"
#include "SeqLib/BFC.h"

static int failures = 0;
#define CHECK(c) do { if (!(c)) { std::fprintf(stderr, "CHECK failed %s:%d: %s\n", __FILE__, __LINE__, #c); ++failures; } } while (0)

int main(int argc, char **argv)
{
    using namespace SeqLib;
    if (argc < 3) { std::fprintf(stderr, "usage: test_fermi reads.tsv assemble|bfc\n"); return 2; }
    std::ifstream in(argv[1]);
    UnalignedSequenceVector reads;
    std::string line;
    while (std::getline(in, line)) {
        std::istringstream ss(line);
        std::string n, s, q;
        std::getline(ss, n, '\t'); std::getline(ss, s, '\t'); std::getline(ss, q, '\t');
        reads.push_back(UnalignedSequence(n, s, q));
    }
    std::string mode = argv[2];
    if (mode == "assemble") {
        FermiAssembler f;
        for (size_t i = 0; i < reads.size() / 2; ++i) f.AddRead(reads[i]);                    // one by one ...
        f.AddReads(UnalignedSequenceVector(reads.begin() + reads.size() / 2, reads.end()));      // ... and in bulk
        CHECK(f.NumSequences() == reads.size());
        f.CorrectReads();                                   // default options: ec_k == 0, an observable no-op
        UnalignedSequenceVector got = f.GetSequences();
        CHECK(got.size() == reads.size());
        for (size_t i = 0; i < got.size(); ++i) CHECK(got[i].Seq == reads[i].Seq && got[i].Name == reads[i].Name);
        f.PerformAssembly();
        std::vector<std::string> c = f.GetContigs();
        for (auto &s : c) std::cout << s << "\n";
        std::ostringstream gfa;
        f.WriteGFA(gfa);
        CHECK(gfa.str().rfind("H\tVN:Z:1.0", 0) == 0);
        {   // the same reads as two windows (all / first half) through the batch entry: each equals its own assembler
            std::vector<UnalignedSequenceVector> wins;
            wins.push_back(reads);
            wins.push_back(UnalignedSequenceVector(reads.begin(), reads.begin() + reads.size() / 2));
            wins.push_back(UnalignedSequenceVector());
            std::vector<std::vector<std::string> > wc = FermiAssembler::AssembleWindows(wins, 0, 2);
            CHECK(wc.size() == 3 && wc[0] == c && wc[2].empty());
            FermiAssembler h;
            h.AddReads(wins[1]);
            h.PerformAssembly();
            CHECK(wc[1] == h.GetContigs());
        }
        f.ClearContigs();
        CHECK(f.GetContigs().empty());
        f.ClearReads();
        CHECK(f.NumSequences() == 0);
    } else {
        BFC b;
        CHECK(!b.AddSequence("", "", "x"));
        CHECK(!b.AddSequence("ACGT", "II", "x"));
        for (auto &r : reads) CHECK(b.AddSequence(r.Seq, r.Qual, r.Name));
        CHECK(b.NumSequences() == (int)reads.size());
        b.Train();
        b.ErrorCorrect();
        float kcov = b.GetKCov();
        CHECK(b.GetKMer() > 0 && (b.GetKMer() & 1));
        CHECK(kcov > 0);
        UnalignedSequenceVector v;
        std::string seq, name;
        size_t n_changed = 0, i = 0;
        while (b.GetSequence(seq, name)) { v.push_back({name, seq}); if (seq != reads[i].Seq) ++n_changed; ++i; }
        CHECK(v.size() == reads.size());
        CHECK(n_changed > 0);
        FermiAssembler f;
        f.AddReads(v);
        f.DirectAssemble(kcov);
        std::cerr << "kmer " << b.GetKMer() << " kcov " << kcov << " corrected reads " << n_changed << "\n";
        for (auto &s : f.GetContigs()) std::cout << s << "\n";
    }
    return failures ? 1 : 0;
}
