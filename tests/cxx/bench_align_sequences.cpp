// Throughput of the drop-in C++ batch path: BWAAligner::alignSequences (b200_mem_align_batch + bam1_t packing on all host
// threads, src/BWAAligner.cpp:151-247) next to the bare ABI call on the same reads.  usage: bench_align_sequences [ref_mb] [n_reads]
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include "SeqLib/BWAAligner.h"
#include "SeqLib/BWAIndex.h"

static uint64_t s_state = 0x5EED0001ull;
static inline uint64_t rnd() { uint64_t z = (s_state += 0x9e3779b97f4a7c15ull); z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull; z = (z ^ (z >> 27)) * 0x94d049bb133111ebull; return z ^ (z >> 31); }

int main(int argc, char **argv)
{
    using namespace SeqLib;
    const size_t ref_len = (size_t)(argc > 1 ? atof(argv[1]) : 50.0) * 1000000, n = argc > 2 ? (size_t)atol(argv[2]) : 1000000, L = 150;
    std::string ref(ref_len, 'A');
    for (size_t i = 0; i < ref_len; ++i) ref[i] = "ACGT"[rnd() & 3];
    // contigs of at most 125 Mb (bntann1_t.len is a 32-bit int, bwa/bntseq.h:38-46): 24 of them for the 3 Gb reference
    UnalignedSequenceVector refs;
    const size_t n_ctg = std::max<size_t>(1, (ref_len + 124999999) / 125000000), per = ref_len / n_ctg;
    for (size_t c = 0; c < n_ctg; ++c) {
        const size_t b = c * per, e = c + 1 < n_ctg ? b + per : ref_len;
        refs.push_back(UnalignedSequence("chrS" + std::to_string(c + 1), ref.substr(b, e - b)));
    }
    BWAIndexPtr idx(new BWAIndex());
    auto t0 = std::chrono::steady_clock::now();
    idx->ConstructIndex(refs);
    double t_index = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    BWAAligner aln(idx);
    UnalignedSequenceVector reads(n);
    static const char comp[256] = {0};
    (void)comp;
    for (size_t i = 0; i < n; ++i) {
        size_t p = rnd() % (ref_len - L);
        if (p / per != (p + L - 1) / per && p / per < n_ctg - 1) p -= L;      // keep the read inside one contig
        std::string s = ref.substr(p, L);
        for (size_t k = 0; k < L; ++k) if (rnd() % 100 == 0) s[k] = "ACGT"[(std::string("ACGT").find(s[k]) + 1 + rnd() % 3) & 3];
        if (rnd() & 1) {                // reverse complement
            std::string r(L, 'A');
            for (size_t k = 0; k < L; ++k) { char c = s[L - 1 - k]; r[k] = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : 'A'; }
            s.swap(r);
        }
        reads[i].Name = "r" + std::to_string(i);
        reads[i].Seq.swap(s);
    }
    std::vector<BamRecordPtrVector> out;
    aln.alignSequences(reads, out, false, 0.9, 10);          // warm-up: device pools, pinned buffers
    double best = 1e30, best_abi = 1e30;
    size_t n_rec = 0;
    for (int it = 0; it < 3; ++it) {
        t0 = std::chrono::steady_clock::now();
        aln.alignSequences(reads, out, false, 0.9, 10);
        double t = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (t < best) best = t;
        n_rec = 0;
        for (auto &v : out) n_rec += v.size();
        // the bare ABI call on the same reads
        std::vector<int64_t> off(n + 1, 0), ids(n);
        std::string all;
        all.reserve(n * L);
        for (size_t i = 0; i < n; ++i) { off[i + 1] = off[i] + (int64_t)reads[i].Seq.size(); ids[i] = (int64_t)i; all += reads[i].Seq; }
        b200_mem_opt_t opt; b200_mem_opt_init(&opt);
        b200_results_t *res = 0;
        t0 = std::chrono::steady_clock::now();
        if (b200_mem_align_batch(idx->handle(), &opt, (int64_t)n, all.data(), off.data(), ids.data(), &res) != 0) { std::fprintf(stderr, "%s\n", b200_last_error()); return 1; }
        t = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (t < best_abi) best_abi = t;
        b200_results_free(res);
    }
    std::printf("{\"ref_bp\": %zu, \"reads\": %zu, \"records\": %zu, \"index_build_s\": %.3f, \"alignSequences_s\": %.4f, \"alignSequences_reads_per_s\": %.0f, "
                "\"abi_call_s\": %.4f, \"abi_reads_per_s\": %.0f, \"packing_and_flatten_s\": %.4f}\n",
                ref_len, n, n_rec, t_index, best, n / best, best_abi, n / best_abi, best - best_abi);
    return 0;
}
