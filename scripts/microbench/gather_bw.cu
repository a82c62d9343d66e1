// gather_bw.cu -- what does a B200 sustain for dependent-free random 32-byte gathers out of a table far larger than L2?
// This is the roofline of FM-index seeding (one 32-byte Occ block per rank query): the stream-copy peak in
// MEASURED_PEAKS.json is not reachable by 32-byte random reads, so the kernel is judged against this number too.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a gather_bw.cu -o gather_bw && ./gather_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33; return x; }

template <int MLP, int BYTES>
__global__ void k_gather(const uint8_t *__restrict__ tab, uint64_t n_blocks, int iters, uint64_t *out)
{
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t acc = 0, s = mix(t + 1);
    for (int it = 0; it < iters; ++it) {
        uint32_t v[MLP][8];
#pragma unroll
        for (int m = 0; m < MLP; ++m) {
            s = mix(s + m + 1);
            const uint8_t *p = tab + (s % n_blocks) * BYTES;
            if (BYTES == 32)
                asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(v[m][0]), "=r"(v[m][1]), "=r"(v[m][2]), "=r"(v[m][3]), "=r"(v[m][4]), "=r"(v[m][5]), "=r"(v[m][6]), "=r"(v[m][7]) : "l"(p));
            else {
                asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(v[m][0]), "=r"(v[m][1]), "=r"(v[m][2]), "=r"(v[m][3]), "=r"(v[m][4]), "=r"(v[m][5]), "=r"(v[m][6]), "=r"(v[m][7]) : "l"(p));
                uint32_t w[8];
                asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "l"(p + 32));
                v[m][0] ^= w[0] ^ w[7];
            }
        }
#pragma unroll
        for (int m = 0; m < MLP; ++m) { acc += v[m][0] + v[m][7]; s ^= v[m][3]; }   // the next addresses depend on the data: one dependent round per iteration
    }
    if (acc == 0x1234567) out[0] = acc;
}

template <int MLP, int BYTES>
static void run(const uint8_t *tab, uint64_t bytes, int blocks_per_sm, int sms, uint64_t *out)
{
    uint64_t nb = bytes / BYTES;
    int iters = 2000 / MLP;
    int grid = sms * blocks_per_sm;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k_gather<MLP, BYTES><<<grid, 128>>>(tab, nb, iters / 4, out);
    cudaEventRecord(a);
    k_gather<MLP, BYTES><<<grid, 128>>>(tab, nb, iters, out);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms = 0; cudaEventElapsedTime(&ms, a, b);
    double n = (double)grid * 128 * iters * MLP;
    printf("{\"bytes_per_gather\": %d, \"mlp_per_thread\": %d, \"threads_per_sm\": %d, \"gathers_per_s\": %.4g, \"GBps\": %.1f, \"latency_bound_us\": %.3f}\n",
           BYTES, MLP, blocks_per_sm * 128, n / (ms * 1e-3), n * BYTES / (ms * 1e-3) / 1e9, ms * 1e3 / iters);
}

int main(int argc, char **argv)
{
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    uint64_t max_bytes = 24ull << 30;
    uint8_t *tab; uint64_t *out;
    cudaMalloc(&tab, max_bytes); cudaMemset(tab, 1, max_bytes); cudaMalloc(&out, 8);
    // table-size sweep at a fixed, ample concurrency: separates cache / TLB reach from DRAM behaviour
    for (uint64_t mb : {32ull, 96ull, 256ull, 512ull, 1024ull, 3072ull, 12288ull, 24576ull}) {
        printf("{\"table_MB\": %llu}\n", (unsigned long long)mb);
        run<4, 32>(tab, mb << 20, 8, sms, out);
        run<4, 64>(tab, mb << 20, 8, sms, out);
    }
    // concurrency sweep on the 3 GB table (the size of the 3 Gb Occ array)
    uint64_t bytes = 3ull << 30;
    for (int bps : {1, 2, 4, 16}) { run<1, 32>(tab, bytes, bps, sms, out); run<4, 32>(tab, bytes, bps, sms, out); }
    return 0;
}
