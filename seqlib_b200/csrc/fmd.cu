// fmd.cu -- FMD-index construction on the GPU: replaces fml_seq2fmi / fml_fmi_gen (fermi-lite/misc.c:65-128), i.e.
// mr_insert_multi in MR_SO_RCLO order (fermi-lite/mrope.c:220-307) + rld_enc (fermi-lite/rld0.c:137-221).
//
// The reference inserts the reversed reads column by column into six ropes (BCR).  The BWT it arrives at is a pure
// function of the read set: the collection is {s, revcomp(s)} for every read without N (an even-length reverse palindrome
// loses its last base first), and row (string j, offset i) sorts by
//        s_j[i..) $ revcomp(s_j[0..i)) $        with $ < A < C < G < T,
// which is the suffix followed -- after the sentinel -- by the suffix of the PARTNER strand that starts where this one's
// prefix ends (that is what "reverse-complement lexicographic order" of the sentinels amounts to; identical rows are
// interchangeable).  So the device sorts all rows by 27-symbol chunks of that key (tie_sorter.cuh), emits the preceding
// symbol of every row, and packs the result into 64-byte rank blocks (fmd.cuh).
#include <cstring>
#include <cub/cub.cuh>
#include "engine.cuh"
#include "tie_sorter.cuh"
#include "fmd.cuh"
#include "fmd_dev.h"

namespace b200 {

thread_local int g_fmd_launches = 0;      // kernels launched by the last fmd_build_device on this thread (cub passes counted per call)

static inline unsigned nblk(u64 n, int t) { return (unsigned)((n + t - 1) / t); }

// per read: usable length (0 = not indexed); fml_fmi_gen's filters (fermi-lite/misc.c:85-96)
__global__ void __launch_bounds__(256) k_fmd_plan(const char *seq, const i64 *off, const i32 *len, i64 n, u32 *rows /* 2n + 1 */)
{
    i64 r = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const char *s = seq + off[r];
    int l = len ? len[r] : (int)(off[r + 1] - off[r]);
    bool ok = l > 0;
    for (int i = 0; i < l && ok; ++i) ok = fmd_nt6(s[i]) < 5;
    if (ok && !(l & 1)) {            // is_rev_same (fermi-lite/misc.c:56-63)
        int i;
        for (i = 0; i < l >> 1; ++i) if (fmd_nt6(s[i]) + fmd_nt6(s[l - 1 - i]) != 5) break;
        if (i == l >> 1) --l;
    }
    u32 v = ok ? (u32)l + 1 : 0u;    // rows of each strand (l symbols + sentinel); l may be 0 after the palindrome cut
    rows[2 * r] = v; rows[2 * r + 1] = v;
}

// text of both strands: string 2r = read r, string 2r+1 = its reverse complement, each followed by 0
__global__ void __launch_bounds__(256) k_fmd_text(const char *seq, const i64 *off, i64 n, const u64 *start /* 2n + 1 */, u8 *text)
{
    i64 r = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    u64 a = start[2 * r], b = start[2 * r + 1];
    int rows = (int)(b - a);
    if (rows == 0) return;
    int l = rows - 1;
    const char *s = seq + off[r];
    for (int i = 0; i < l; ++i) {
        int c = fmd_nt6(s[i]);
        text[a + i] = (u8)c;
        text[b + l - 1 - i] = (u8)(5 - c);
    }
    text[a + l] = 0; text[b + l] = 0;
}

struct RotKey {              // chunk `depth` of the key of row p (its string comes from row_str)
    const u8 *text; const u64 *start; const u32 *row_str;
    __device__ __forceinline__ u64 operator()(u64 p, int depth) const
    {
        u64 j = row_str[p];
        u64 sj = start[j], sp = start[j ^ 1];
        int i = (int)(p - sj);
        int L = (int)(start[j + 1] - sj) - 1;
        const u8 *a = text + sj + i;             // chars d = 0 .. L - i   (suffix + its sentinel)
        const u8 *b = text + sp - 1;             // chars d = L - i + 1 .. L + 1  -> partner[d - 1]
        int d0 = depth * FMD_KEY_SYMS, split = L - i, last = L + 1;
        u64 key = 0;
#pragma unroll 1
        for (int t = 0; t < FMD_KEY_SYMS; ++t) {
            int d = d0 + t;
            u64 c = d <= split ? a[d] : (d <= last ? b[d] : 0);
            key = key * 5 + c;
        }
        return key;
    }
};

__global__ void __launch_bounds__(256) k_fmd_row_str(const u64 *start, u64 n_str, u32 *row_str)
{
    u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_str) return;
    for (u64 p = start[j], b = start[j + 1]; p < b; ++p) row_str[p] = (u32)j;
}

__global__ void __launch_bounds__(256) k_fmd_rows(u64 n, RotKey kf, u64 *keys, u32 *ids)
{
    u64 p = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    ids[p] = (u32)p;
    keys[p] = kf(p, 0);
}

// BWT symbol of every sorted row
__global__ void __launch_bounds__(256) k_fmd_emit(const u8 *text, const u64 *start, const u32 *row_str, const u32 *ids, u64 n, u8 *bwt8)
{
    u64 x = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= n) return;
    u64 p = ids[x];
    bwt8[x] = p != start[row_str[p]] ? text[p - 1] : (u8)0;      // the symbol before the row's suffix; rows at offset 0 carry $
}

// ------------------------------------------------------------------------------------------------ prefix doubling
// The row key s[i..) $ revcomp(s[0..i)) $ is (A $, B $) with A = s_j[i..) and B = s_j'[L - i ..) -- both SUFFIXES of strings of
// the collection.  So: (1) rank every row by its own suffix A $ (equal suffixes share a rank): one radix sort on the first
// 27 symbols, then doubling rounds on (rank of row p, rank of row p + h), h = 27, 54, 108, ...; (2) one last sort on
// (suffix rank of the row, suffix rank of the partner's row at offset L - i).  log2(L / 27) + 3 sorts of all rows instead of
// one refinement round per 27 symbols of the deepest tie.
__global__ void __launch_bounds__(256) k_pd_key0(const u8 *text, const u64 *start, const u32 *row_str, u64 n, u64 *keys, u32 *ids)
{
    u64 p = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    u64 e = start[row_str[p] + 1] - 1;            // position of the string's sentinel
    u64 key = 0;
#pragma unroll 1
    for (int t = 0; t < FMD_KEY_SYMS; ++t) { u64 q = p + t; key = key * 5 + (q <= e ? (u64)text[q] : 0ull); }
    keys[p] = key; ids[p] = (u32)p;
}

// head[x] = first element of a run of equal keys (in sorted order); hidx = its own index there, 0 elsewhere
__global__ void __launch_bounds__(256) k_pd_heads(const u64 *keys, u64 n, u32 *hidx, u32 *n_groups_partial)
{
    u64 x = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= n) return;
    bool head = x == 0 || keys[x] != keys[x - 1];
    hidx[x] = head ? (u32)x : 0u;
    if (head) atomicAdd(n_groups_partial, 1u);
}

// rank[row] = index of the head of the row's run (after an inclusive max-scan of hidx)
__global__ void __launch_bounds__(256) k_pd_scatter_rank(const u32 *ids, const u32 *hidx, u64 n, u32 *rank)
{
    u64 x = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= n) return;
    rank[ids[x]] = hidx[x];
}

// doubling key of the row at sorted position x: (rank of p) << 32 | (rank of p + h) + 1, 0 past the string's sentinel
__global__ void __launch_bounds__(256) k_pd_key_double(const u32 *ids, const u32 *rank, const u64 *start, const u32 *row_str, u64 n, u32 h, u64 *keys)
{
    u64 x = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= n) return;
    u64 p = ids[x];
    u64 e = start[row_str[p] + 1] - 1;
    u64 q = p + h;
    u64 r2 = q <= e ? (u64)rank[q] + 1 : 0ull;
    keys[x] = (u64)rank[p] << 32 | r2;
}

// final key: (suffix rank of the row) << 32 | suffix rank of the partner strand's row at offset L - i
__global__ void __launch_bounds__(256) k_pd_key_final(const u32 *ids, const u32 *rank, const u64 *start, const u32 *row_str, u64 n, u64 *keys)
{
    u64 x = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= n) return;
    u64 p = ids[x];
    u64 j = row_str[p], sj = start[j], L = start[j + 1] - sj - 1, i = p - sj;
    u64 pp = start[j ^ 1] + (L - i);
    keys[x] = (u64)rank[p] << 32 | (u64)rank[pp];
}

struct PdMax { __device__ __forceinline__ u32 operator()(u32 a, u32 b) const { return a > b ? a : b; } };

struct Cnt4 { u32 c[4]; };
struct Cnt4Add { __device__ __forceinline__ Cnt4 operator()(const Cnt4 &a, const Cnt4 &b) const { Cnt4 r; for (int i = 0; i < 4; ++i) r.c[i] = a.c[i] + b.c[i]; return r; } };

__global__ void __launch_bounds__(128) k_fmd_pack(const u8 *bwt8, u64 n, FmdBlock *blk, Cnt4 *cnt, u64 n_blk)
{
    u64 b = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_blk) return;
    FmdBlock B; memset(&B, 0, sizeof(B));
    Cnt4 c; c.c[0] = c.c[1] = c.c[2] = c.c[3] = 0;
    for (int j = 0; j < 128; ++j) {
        u64 x = (b << 7) + j;
        if (x >= n) break;
        int s = bwt8[x], h = j >> 6, t = j & 63;
        if (s == 0) B.p2[h] |= 1ull << t;
        else { ++c.c[s - 1]; B.p0[h] |= (u64)((s - 1) & 1) << t; B.p1[h] |= (u64)((s - 1) >> 1) << t; }
    }
    blk[b] = B;
    cnt[b] = c;
}

__global__ void __launch_bounds__(256) k_fmd_counts(const Cnt4 *excl, FmdBlock *blk, u64 n_blk)
{
    u64 b = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_blk) return;
    for (int i = 0; i < 4; ++i) blk[b].cnt[i] = excl[b].c[i];
}

void fmd_build_device(FmdDevice &F, const char *d_seq, const i64 *d_off, const i32 *d_len, i64 n_reads, cudaStream_t st_unused)
{
    (void)st_unused;        // the sorter runs on the default stream; everything here follows it
    F.release();
    g_fmd_launches = 0;
    const u64 n_str = 2 * (u64)n_reads;
    if (n_reads == 0) return;
    DevBuf rows, tmp;
    rows.reserve((n_str + 1) * 4); F.start.reserve((n_str + 1) * 8);
    CU_CHECK(cudaMemset(rows.p, 0, (n_str + 1) * 4));
    k_fmd_plan<<<nblk(n_reads, 256), 256>>>(d_seq, d_off, d_len, n_reads, rows.as<u32>());
    {
        size_t tb = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tb, rows.as<u32>(), F.start.as<u64>(), (int)(n_str + 1));
        tmp.reserve(tb);
        CU_CHECK(cub::DeviceScan::ExclusiveSum(tmp.p, tb, rows.as<u32>(), F.start.as<u64>(), (int)(n_str + 1)));
    }
    u64 n = 0;
    CU_CHECK(cudaMemcpy(&n, F.start.as<u64>() + n_str, 8, cudaMemcpyDeviceToHost));
    if (n == 0) return;
    if (n >= (1ull << 31) - 1024) throw std::length_error("FMD-index of more than 2^31 symbols");
    F.text.reserve(n + 16);
    k_fmd_text<<<nblk(n_reads, 256), 256>>>(d_seq, d_off, n_reads, F.start.as<u64>(), F.text.as<u8>());
    // sort all rows (prefix doubling, see above)
    DevBuf keys[2], ids[2], row_str, rank, hidx, ngrp;
    keys[0].reserve(n * 8); keys[1].reserve(n * 8); ids[0].reserve(n * 4); ids[1].reserve(n * 4);
    row_str.reserve(n * 4); rank.reserve(n * 4); hidx.reserve(n * 4); ngrp.reserve(16);
    k_fmd_row_str<<<nblk(n_str, 256), 256>>>(F.start.as<u64>(), n_str, row_str.as<u32>());
    u32 max_rows = 0;
    {
        DevBuf mx; mx.reserve(16);
        size_t tb = 0;
        cub::DeviceReduce::Max(nullptr, tb, rows.as<u32>(), mx.as<u32>(), (int)n_str);
        tmp.reserve(tb);
        CU_CHECK(cub::DeviceReduce::Max(tmp.p, tb, rows.as<u32>(), mx.as<u32>(), (int)n_str));
        CU_CHECK(cudaMemcpy(&max_rows, mx.p, 4, cudaMemcpyDeviceToHost));
    }
    int cur = 0;
    auto sort_rows = [&](int bits) {
        cub::DoubleBuffer<u64> dk(keys[cur].as<u64>(), keys[cur ^ 1].as<u64>());
        cub::DoubleBuffer<u32> dv(ids[cur].as<u32>(), ids[cur ^ 1].as<u32>());
        size_t tb = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tb, dk, dv, (int)n, 0, bits);
        tmp.reserve(tb);
        CU_CHECK(cub::DeviceRadixSort::SortPairs(tmp.p, tb, dk, dv, (int)n, 0, bits));
        if (dk.Current() != keys[cur].as<u64>()) { std::swap(keys[0].p, keys[1].p); std::swap(keys[0].cap, keys[1].cap); }
        if (dv.Current() != ids[cur].as<u32>()) { std::swap(ids[0].p, ids[1].p); std::swap(ids[0].cap, ids[1].cap); }
        g_fmd_launches += (bits + 7) / 8 + 2;          // onesweep: histogram + scan + one pass per 8 bits
    };
    auto rerank = [&]() -> u32 {            // rank[] from the sorted keys; returns the number of distinct keys
        CU_CHECK(cudaMemset(ngrp.p, 0, 4));
        k_pd_heads<<<nblk(n, 256), 256>>>(keys[cur].as<u64>(), n, hidx.as<u32>(), ngrp.as<u32>());
        size_t tb = 0;
        cub::DeviceScan::InclusiveScan(nullptr, tb, hidx.as<u32>(), hidx.as<u32>(), PdMax(), (int)n);
        tmp.reserve(tb);
        CU_CHECK(cub::DeviceScan::InclusiveScan(tmp.p, tb, hidx.as<u32>(), hidx.as<u32>(), PdMax(), (int)n));
        k_pd_scatter_rank<<<nblk(n, 256), 256>>>(ids[cur].as<u32>(), hidx.as<u32>(), n, rank.as<u32>());
        u32 g = 0;
        CU_CHECK(cudaMemcpy(&g, ngrp.p, 4, cudaMemcpyDeviceToHost));
        g_fmd_launches += 4;
        return g;
    };
    k_pd_key0<<<nblk(n, 256), 256>>>(F.text.as<u8>(), F.start.as<u64>(), row_str.as<u32>(), n, keys[cur].as<u64>(), ids[cur].as<u32>());
    sort_rows(63);
    u32 groups = rerank();
    for (u32 h = FMD_KEY_SYMS; h < max_rows && groups < n; h <<= 1) {
        k_pd_key_double<<<nblk(n, 256), 256>>>(ids[cur].as<u32>(), rank.as<u32>(), F.start.as<u64>(), row_str.as<u32>(), n, h, keys[cur].as<u64>());
        sort_rows(64);
        groups = rerank();
    }
    k_pd_key_final<<<nblk(n, 256), 256>>>(ids[cur].as<u32>(), rank.as<u32>(), F.start.as<u64>(), row_str.as<u32>(), n, keys[cur].as<u64>());
    sort_rows(64);
    u32 *ip = ids[cur].as<u32>();
    // BWT symbols, rank blocks
    DevBuf bwt8; bwt8.reserve(n + 16);
    k_fmd_emit<<<nblk(n, 256), 256>>>(F.text.as<u8>(), F.start.as<u64>(), row_str.as<u32>(), ip, n, bwt8.as<u8>());
    const u64 n_blk = (n + 127) / 128 + 1;          // one spare block: rank(n) with n % 128 == 0 stays in range
    F.blocks.reserve(n_blk * sizeof(FmdBlock));
    DevBuf cnt; cnt.reserve(n_blk * sizeof(Cnt4));
    k_fmd_pack<<<nblk(n_blk, 128), 128>>>(bwt8.as<u8>(), n, F.blocks.as<FmdBlock>(), cnt.as<Cnt4>(), n_blk);
    {
        size_t tb = 0; Cnt4 zero; zero.c[0] = zero.c[1] = zero.c[2] = zero.c[3] = 0;
        cub::DeviceScan::ExclusiveScan(nullptr, tb, cnt.as<Cnt4>(), cnt.as<Cnt4>(), Cnt4Add(), zero, (int)n_blk);
        tmp.reserve(tb);
        CU_CHECK(cub::DeviceScan::ExclusiveScan(tmp.p, tb, cnt.as<Cnt4>(), cnt.as<Cnt4>(), Cnt4Add(), zero, (int)n_blk));
    }
    k_fmd_counts<<<nblk(n_blk, 256), 256>>>(cnt.as<Cnt4>(), F.blocks.as<FmdBlock>(), n_blk);
    Cnt4 last;
    CU_CHECK(cudaMemcpy(&last, cnt.as<Cnt4>() + (n_blk - 1), sizeof(Cnt4), cudaMemcpyDeviceToHost));
    CU_CHECK(cudaStreamSynchronize(0));      // this file is compiled with --default-stream per-thread: 0 = the calling thread's own stream
    CU_CHECK(cudaGetLastError());
    // the spare block is empty, so its running counts are the totals
    u64 tot[6] = {0, last.c[0], last.c[1], last.c[2], last.c[3], 0};
    tot[0] = n - (tot[1] + tot[2] + tot[3] + tot[4]);
    F.idx.blk = F.blocks.as<FmdBlock>(); F.idx.n = n; F.idx.n_str = tot[0];
    F.idx.cnt[0] = 0;
    for (int c = 0; c < 6; ++c) F.idx.cnt[c + 1] = F.idx.cnt[c] + tot[c];
    g_fmd_launches += 14;      // plan, scans, text, row map, keys, emit, pack, counts
    F.n_blk = n_blk;
    F.bwt8.p = bwt8.p; F.bwt8.cap = bwt8.cap; bwt8.p = nullptr; bwt8.cap = 0;     // kept for b200_fmd_bwt (debug / parity)
}

} // namespace b200
