// int_peak.cu -- measured integer issue peaks of a B200 for the instruction kinds the packed 16-bit Smith-Waterman
// wavefront (seqlib_b200/csrc/ksw_wave.cuh) is made of: the roofline denominator of `bench.py --workload ksw`.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a int_peak.cu -o int_peak && ./int_peak
// Every thread runs 8 independent dependency chains so that the pipes, not latencies, bound the rate.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int KIND>
__global__ void __launch_bounds__(256) k_peak(unsigned *out, int iters, unsigned seed)
{
    unsigned a[8], b = seed | 0x00010001u, c = seed * 3u + 7u;
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = threadIdx.x * 8u + k + seed;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (KIND == 0) a[k] = __vimax3_s16x2(a[k], b, c);                       // VIMNMX3.S16x2
                else if (KIND == 1) a[k] = __viaddmax_s16x2_relu(a[k], b, c);           // VIADDMNMX.S16x2.RELU
                else if (KIND == 2) a[k] = (a[k] & b) | (~a[k] & c);                    // LOP3
                else if (KIND == 3) a[k] = __vadd2(a[k], b);                            // VIADD.16x2
                else if (KIND == 4) a[k] = a[k] * b + c;                                // IMAD (FMA pipe)
                else if (KIND == 5) { a[k] = (k & 1) ? a[k] * b + c : __vimax3_s16x2(a[k], b, c); }   // half IMAD, half VIMNMX3
                else a[k] = __byte_perm(a[k], b, 0x5410) ^ c;                           // PRMT + LOP3
            }
            b += 0x00010001u;
        }
    }
    unsigned x = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) x ^= a[k];
    if (x == 0x12345678u) out[0] = x;
}

template <int KIND>
static double run(const char *name, int per_iter, unsigned *out, int sms)
{
    const int iters = 4096, grid = sms * 8;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_peak<KIND><<<grid, 256>>>(out, 64, 1);
    cudaEventRecord(e0);
    k_peak<KIND><<<grid, 256>>>(out, iters, 1);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)grid * 256 * iters * 32 * per_iter;
    double rate = ops / (ms * 1e-3);
    printf("{\"kind\": \"%s\", \"thread_instr_per_s\": %.4g, \"warp_instr_per_clk_per_sm\": %.3f}\n", name, rate, rate / 32 / sms / 1.965e9);
    return rate;
}

int main()
{
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    unsigned *out; cudaMalloc(&out, 64);
    run<0>("VIMNMX3.S16x2", 1, out, sms);
    run<1>("VIADDMNMX.S16x2.RELU", 1, out, sms);
    run<2>("LOP3", 1, out, sms);
    run<3>("VIADD.16x2", 1, out, sms);
    run<4>("IMAD", 1, out, sms);
    run<5>("IMAD+VIMNMX3 1:1", 1, out, sms);
    run<6>("PRMT+LOP3", 2, out, sms);
    return 0;
}
