// fml_host.h -- host-side arithmetic of fml_correct_core (fermi-lite/bfc.c:513-553) shared by fml.cu and the CPU
// emulation in tests/hostsim: option derivation only, no per-read work.
#pragma once
#include <cstdint>
#include "bfc.cuh"

namespace b200 {

// the tail of bfc_ch_hist (fermi-lite/htab.c:122-126)
inline int fml_hist_mode(const uint64_t cnt[256])
{
    int max_i = -1; uint64_t max = 0;
    for (int i = 3; i < 256; ++i) if (cnt[i] > max) { max = cnt[i]; max_i = i; }
    return max_i;
}

// l_pre as fml_correct_core derives it from the total read length (u64 arithmetic: tot_len < 8 wraps and yields 20)
inline int fml_initial_l_pre(uint64_t tot_len) { return tot_len - 8 < 20 ? (int)(tot_len - 8) : 20; }

// kcov and the solid-k-mer threshold (fermi-lite/bfc.c:538-543)
inline void fml_kcov_min_cov(const uint64_t hist[256], int min_cnt, int max_cnt, float &kcov, int &min_cov)
{
    uint64_t sum_k = 0, tot_k = 0;
    for (int i = min_cnt < 0 ? 0 : min_cnt; i < 256; ++i) { sum_k += hist[i]; tot_k += (uint64_t)i * hist[i]; }
    kcov = (float)tot_k / sum_k;       // NaN when no k-mer reaches min_cnt, like the reference
    // (int)NaN is INT_MIN on x86-64, which the two clamps below turn into min_cnt
    min_cov = sum_k ? (int)(.1 * kcov + .499) : min_cnt;
    min_cov = min_cov < max_cnt ? min_cov : max_cnt;
    min_cov = min_cov > min_cnt ? min_cov : min_cnt;
}

} // namespace b200
