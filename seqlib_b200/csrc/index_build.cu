// index_build.cu -- exact FM-index construction on the GPU.
//
// Replaces, for BWAIndex::ConstructIndex (src/BWAIndex.cpp:125-138), the CPU
// chain  is_bwt (bwa/is.c:208, SA-IS, `int` lengths)  ->  bwt_bwtupdate_core
// (bwa/bwtindex.c:149-171)  ->  bwt_cal_sa (bwa/bwt.c:62-84, a sequential
// LF walk over the whole text).  The BWT and suffix array are pure functions
// of the text, so any exact suffix sort gives bit-identical arrays; this one
// is built for HBM: no `int` limit and 6*10^9 suffixes are sorted in place.
//
//   1. key(p, d) = the 29 bases at text[p+29d ..) (58 bits, zero padded) << 5 | min(remaining, 29).
//      The low field makes a suffix that ends inside the window sort before every
//      longer suffix sharing the padded bases (end of text = smallest symbol).
//   2. a 4096-bin histogram on the top 6 bases splits the suffixes into super-buckets
//      of bounded size; each is collected, radix-sorted on key(p,0) (cub), and runs of
//      equal keys are refined with key(p,1), key(p,2), ... (only repeats survive a round).
//   3. in rank order: BWT symbol text[p-1] and the SA sample are emitted; the rank of
//      suffix 0 is `primary`; the Occ blocks get their counts from one scan.
#include <cub/cub.cuh>
#include <vector>
#include <algorithm>
#include <cstdio>
#include "engine.cuh"
#include "tie_sorter.cuh"

namespace b200 {

#define KEY_BASES 29

// 29 bases starting at q (q may be >= N), packed MSB-first into 58 bits, then << 5 | min(max(N-q,0),29)
__device__ __forceinline__ u64 suffix_key(const u64 *__restrict__ text, u64 N, u64 q)
{
    if (q >= N) return 0;
    u64 rem = N - q;
    int r = rem < KEY_BASES ? (int)rem : KEY_BASES;
    u64 wi = q >> 5; int sh = (int)(q & 31) * 2;
    u64 lo = text[wi], hi = text[wi + 1];                  // text has one spare word
    u64 bits = sh ? (lo >> sh) | (hi << (64 - sh)) : lo;   // base j of the window in bits 2j..2j+1
    if (r < KEY_BASES) bits &= (1ull << (2 * r)) - 1;       // (r = 0 cannot happen here)
    else bits &= (1ull << (2 * KEY_BASES)) - 1;
    // reverse the order of the 2-bit groups so that base 0 is most significant
    u64 x = bits;
    x = ((x >> 2) & 0x3333333333333333ull) | ((x & 0x3333333333333333ull) << 2);
    x = ((x >> 4) & 0x0f0f0f0f0f0f0f0full) | ((x & 0x0f0f0f0f0f0f0f0full) << 4);
    x = ((x >> 8) & 0x00ff00ff00ff00ffull) | ((x & 0x00ff00ff00ff00ffull) << 8);
    x = ((x >> 16) & 0x0000ffff0000ffffull) | ((x & 0x0000ffff0000ffffull) << 16);
    x = (x >> 32) | (x << 32);                               // base 0 now in bits 63..62
    x >>= 64 - 2 * KEY_BASES;                                // 58 significant bits
    return x << 5 | (u64)r;
}

__global__ void k_hist(const u64 *__restrict__ text, u64 N, unsigned long long *hist)
{
    __shared__ unsigned int sh[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    for (u64 p = (u64)blockIdx.x * blockDim.x + threadIdx.x; p < N; p += (u64)gridDim.x * blockDim.x)
        atomicAdd(&sh[suffix_key(text, N, p) >> 51], 1u);
    __syncthreads();
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) if (sh[i]) atomicAdd(&hist[i], (unsigned long long)sh[i]);
}

// append every suffix whose top-6-base bin lies in [lo, hi) (order irrelevant: it is sorted next)
__global__ void k_collect(const u64 *__restrict__ text, u64 N, u32 lo, u32 hi, u64 *__restrict__ keys, u64 *__restrict__ pos,
                          unsigned long long *counter)
{
    for (u64 p0 = (u64)blockIdx.x * blockDim.x; p0 < N; p0 += (u64)gridDim.x * blockDim.x) {
        u64 p = p0 + threadIdx.x;
        u64 key = 0; bool take = false;
        if (p < N) { key = suffix_key(text, N, p); u32 bin = (u32)(key >> 51); take = bin >= lo && bin < hi; }
        unsigned m = __ballot_sync(0xffffffffu, take);
        if (m) {
            int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
            unsigned long long base = 0;
            if (lane == leader) base = atomicAdd(counter, (unsigned long long)__popc(m));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (take) { u64 o = base + __popc(m & ((1u << lane) - 1)); keys[o] = key; pos[o] = p; }
        }
    }
}

struct TextKey {          // chunk d of suffix p = the 29 bases at p + 29 d
    const u64 *text; u64 N;
    __device__ __forceinline__ u64 operator()(u64 p, int depth) const { return suffix_key(text, N, p + (u64)KEY_BASES * depth); }
};

// emit BWT symbols (one byte per rank, 4 = '$') and SA samples for ranks [rank0, rank0+n)
__global__ void k_emit(const u64 *__restrict__ text, u64 N, const u64 *__restrict__ pos, u64 n, u64 rank0, u8 *__restrict__ bwt8,
                       u64 *__restrict__ sa, int sa_shift, unsigned long long *primary)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u64 f = rank0 + i, p = pos[i];
    if (p == 0) { bwt8[f] = 4; *primary = f; }
    else { u64 q = p - 1; bwt8[f] = (u8)((text[q >> 5] >> (2 * (q & 31))) & 3); }
    if ((f & ((1ull << sa_shift) - 1)) == 0) sa[f >> sa_shift] = p;
}

struct Cnt4 { u32 c[4]; };
struct Cnt4Add { __device__ __forceinline__ Cnt4 operator()(const Cnt4 &a, const Cnt4 &b) const { Cnt4 r; for (int i = 0; i < 4; ++i) r.c[i] = a.c[i] + b.c[i]; return r; } };

// pack 64 symbols per block from bwt8 (skipping the '$' at `primary`) and count them
__global__ void k_pack_blocks(const u8 *__restrict__ bwt8, u64 N, u64 primary, OccBlock *__restrict__ occ, Cnt4 *__restrict__ cnt, u64 n_occ)
{
    u64 b = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_occ) return;
    u64 s0 = 0, s1 = 0; Cnt4 c; c.c[0] = c.c[1] = c.c[2] = c.c[3] = 0;
    for (int j = 0; j < 64; ++j) {
        u64 x = b * 64 + j;
        if (x >= N) break;
        u64 f = x < primary ? x : x + 1;
        u64 s = bwt8[f];
        ++c.c[s];
        s0 |= (s & 1) << j; s1 |= (s >> 1) << j;
    }
    occ[b].sym[0] = s0; occ[b].sym[1] = s1;
    cnt[b] = c;
}
__global__ void k_store_counts(const Cnt4 *__restrict__ excl, OccBlock *__restrict__ occ, u64 n_occ)
{
    u64 b = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_occ) return;
    for (int i = 0; i < 4; ++i) occ[b].cnt[i] = excl[b].c[i];
}

void build_fm_index_device(b200_index *idx, const BlobHeader &h)
{
    const u64 N = h.seq_len;
    u8 *blob = (u8 *)idx->d_blob;
    const u64 *text = (const u64 *)(blob + h.off_text);
    OccBlock *occ = (OccBlock *)(blob + h.off_occ);
    u64 *sa = (u64 *)(blob + h.off_sa);
    int dev = 0, sms = 148; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);

    // 1. histogram of the top 6 bases
    DevBuf d_hist; d_hist.reserve(4096 * 8 + 64);
    CU_CHECK(cudaMemset(d_hist.p, 0, 4096 * 8 + 64));
    k_hist<<<sms * 8, 256>>>(text, N, d_hist.as<unsigned long long>());
    std::vector<unsigned long long> hist(4096);
    CU_CHECK(cudaMemcpy(hist.data(), d_hist.p, 4096 * 8, cudaMemcpyDeviceToHost));
    // 2. super-buckets
    size_t fr = 0, tot = 0; CU_CHECK(cudaMemGetInfo(&fr, &tot));
    u64 maxb = 1ull << 29;                                    // elements per super-bucket
    u64 by_mem = (u64)(fr / 2) / 48;                          // keys+pos double buffered + scratch
    if (by_mem < maxb) maxb = by_mem;
    if (const char *e = getenv("B200_BUILD_MAXB")) maxb = strtoull(e, 0, 10);
    if (maxb < 1024) maxb = 1024;
    DevBuf bwt8; bwt8.reserve(N + 2);
    DevBuf keys, pos, d_ctr; d_ctr.reserve(64);
    unsigned long long *d_primary = d_ctr.as<unsigned long long>() + 1;
    CU_CHECK(cudaMemset(d_ctr.p, 0, 64));
    Sorter<TextKey> S; S.kf.text = text; S.kf.N = N;
    // rank 0 = the empty suffix: BWT[0] = text[N-1], SA[0] = N (stored as -1 like bwt_cal_sa, bwa/bwt.c:83)
    {
        u64 q = N - 1, w;
        CU_CHECK(cudaMemcpy(&w, text + (q >> 5), 8, cudaMemcpyDeviceToHost));
        u8 s = (u8)((w >> (2 * (q & 31))) & 3);
        CU_CHECK(cudaMemcpy(bwt8.p, &s, 1, cudaMemcpyHostToDevice));
        u64 m1 = ~0ull;
        CU_CHECK(cudaMemcpy(sa, &m1, 8, cudaMemcpyHostToDevice));
    }
    u64 rank0 = 1;
    u32 lo = 0;
    while (lo < 4096) {
        u64 cnt = hist[lo]; u32 hi = lo + 1;
        while (hi < 4096 && cnt + hist[hi] <= maxb) cnt += hist[hi++];
        if (cnt > (1ull << 31) - 1024) throw std::runtime_error("a 6-base prefix bucket exceeds 2^31 suffixes");
        if (cnt) {
            keys.reserve(cnt * 8); pos.reserve(cnt * 8);
            CU_CHECK(cudaMemset(d_ctr.p, 0, 8));
            k_collect<<<sms * 8, 256>>>(text, N, lo, hi, keys.as<u64>(), pos.as<u64>(), d_ctr.as<unsigned long long>());
            u64 *kp = keys.as<u64>(), *pp = pos.as<u64>();
            S.sort_bucket(kp, pp, cnt);
            k_emit<<<nb(cnt, 256), 256>>>(text, N, pp, cnt, rank0, bwt8.as<u8>(), sa, h.sa_shift, d_primary);
            CU_CHECK(cudaDeviceSynchronize());
            rank0 += cnt;
        }
        lo = hi;
    }
    if (rank0 != N + 1) throw std::runtime_error("suffix count mismatch");
    unsigned long long primary = 0;
    CU_CHECK(cudaMemcpy(&primary, d_primary, 8, cudaMemcpyDeviceToHost));
    // 3. Occ blocks
    DevBuf cnt, tmp; cnt.reserve(h.n_occ * sizeof(Cnt4));
    k_pack_blocks<<<nb(h.n_occ, 128), 128>>>(bwt8.as<u8>(), N, primary, occ, cnt.as<Cnt4>(), h.n_occ);
    {
        size_t tb = 0; Cnt4 zero; zero.c[0] = zero.c[1] = zero.c[2] = zero.c[3] = 0;
        cub::DeviceScan::ExclusiveScan(nullptr, tb, cnt.as<Cnt4>(), cnt.as<Cnt4>(), Cnt4Add(), zero, (int)h.n_occ);
        tmp.reserve(tb);
        CU_CHECK(cub::DeviceScan::ExclusiveScan(tmp.p, tb, cnt.as<Cnt4>(), cnt.as<Cnt4>(), Cnt4Add(), zero, (int)h.n_occ));
    }
    k_store_counts<<<nb(h.n_occ, 256), 256>>>(cnt.as<Cnt4>(), occ, h.n_occ);
    CU_CHECK(cudaDeviceSynchronize());
    // totals = counts before the last (empty or partial) block + its symbols: take them from the last block + its content
    Cnt4 last; OccBlock lb;
    CU_CHECK(cudaMemcpy(&last, cnt.as<Cnt4>() + (h.n_occ - 1), sizeof(Cnt4), cudaMemcpyDeviceToHost));
    CU_CHECK(cudaMemcpy(&lb, occ + (h.n_occ - 1), sizeof(OccBlock), cudaMemcpyDeviceToHost));
    u64 totals[4] = {last.c[0], last.c[1], last.c[2], last.c[3]};
    {
        u64 x0 = (h.n_occ - 1) * 64;
        for (u64 x = x0; x < N; ++x) { int j = (int)(x - x0); ++totals[occ_sym(lb, j)]; }
    }
    idx->primary = primary;
    idx->L2[0] = 0;
    for (int c = 0; c < 4; ++c) idx->L2[c + 1] = idx->L2[c] + totals[c];
    if (idx->L2[4] != N) throw std::runtime_error("BWT symbol count mismatch");
    for (int c = 0; c < 4; ++c) if (totals[c] >= (1ull << 32)) throw std::runtime_error("a single base occurs >= 2^32 times: outside the 32-bit Occ block layout");
    idx->seq_len = N;
    idx->dev.primary = primary; for (int c = 0; c < 5; ++c) idx->dev.L2[c] = idx->L2[c];
}

} // namespace b200
