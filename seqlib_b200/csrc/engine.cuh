// engine.cuh -- internal declarations shared by engine.cu, index.cu and index_build.cu
#pragma once
#include <cstdio>
#include <cuda_runtime.h>
#include <string>
#include <vector>
#include <stdexcept>
#include <mutex>
#include <cstdlib>
#include "pipeline.cuh"
#include "hostutil.h"

namespace b200 {

void set_error(const std::string &msg);
int fail(int code, const std::string &msg);

struct CudaError : std::runtime_error { using std::runtime_error::runtime_error; };

#define CU_CHECK(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) \
    throw b200::CudaError(std::string(#expr) + ": " + cudaGetErrorString(e__) + " (" __FILE__ ":" + std::to_string(__LINE__) + ")"); } while (0)

// Device memory comes from a process-wide pool: cudaMalloc / cudaFree of the 10^8..10^9-byte buffers a batch needs cost
// 10-500 ms per call (cudaFree also synchronises the device), which was most of the end-to-end time of
// b200_mem_align_batch.  Blocks are handed back on release() and reused best-fit; the pool keeps at most
// B200_DEVPOOL_GB (default 48) GB and frees the rest.  A block is only released after the work that used it was
// synchronised (every caller syncs its stream before its buffers go out of scope).
struct DevPool {
    struct Blk { void *p; size_t cap; int dev; };
    std::vector<Blk> free_;
    std::mutex mu;
    size_t pooled = 0, limit = 0;
    size_t n_malloc = 0, n_free = 0, n_hit = 0, slots = 0;      // B200_POOL_DEBUG=1 prints them at exit
    ~DevPool() {}
    void report() const { fprintf(stderr, "[devpool] cudaMalloc %zu, cudaFree %zu, reused %zu, pooled %.1f MB in %zu blocks\n", n_malloc, n_free, n_hit, pooled / 1048576.0, free_.size()); }
    void *get(size_t want, size_t &cap)
    {
        int dev = 0; cudaGetDevice(&dev);
        {
            std::lock_guard<std::mutex> g(mu);
            int best = -1;
            for (size_t i = 0; i < free_.size(); ++i)
                if (free_[i].dev == dev && free_[i].cap >= want && free_[i].cap <= 2 * want + (1u << 20) && (best < 0 || free_[i].cap < free_[best].cap)) best = (int)i;
            if (best >= 0) { Blk b = free_[best]; free_.erase(free_.begin() + best); pooled -= b.cap; cap = b.cap; ++n_hit; return b.p; }
            ++n_malloc;
        }
        void *p = nullptr;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {              // out of memory: drop everything pooled and try once more
            cudaGetLastError();
            trim(0);
            e = cudaMalloc(&p, want);
        }
        if (e != cudaSuccess) throw CudaError(std::string("cudaMalloc(") + std::to_string(want) + "): " + cudaGetErrorString(e));
        cap = want;
        return p;
    }
    void put(void *p, size_t cap)
    {
        if (!p) return;
        int dev = 0; cudaGetDevice(&dev);
        std::lock_guard<std::mutex> g(mu);
        if (!limit) { const char *e = getenv("B200_DEVPOOL_GB"); limit = (size_t)((e ? atof(e) : 48.0) * (1ull << 30)); if (!limit) limit = 1; }
        if (!slots) { const char *e = getenv("B200_POOL_SLOTS"); slots = e ? (size_t)atol(e) : 4096; if (!slots) slots = 1; }
        if (cap > limit / 2 || pooled + cap > limit || free_.size() >= slots) { ++n_free; cudaFree(p); return; }
        free_.push_back(Blk{p, cap, dev}); pooled += cap;
    }
    void trim(size_t keep)
    {
        std::lock_guard<std::mutex> g(mu);
        while (!free_.empty() && pooled > keep) { cudaFree(free_.back().p); pooled -= free_.back().cap; free_.pop_back(); }
    }
};
inline DevPool &dev_pool()
{
    static DevPool *p = [] { DevPool *q = new DevPool(); if (getenv("B200_POOL_DEBUG")) atexit([] { dev_pool().report(); }); return q; }();
    return *p;
}     // never destroyed: outlives static DevBufs

// A device buffer that grows on demand (never shrinks), backed by the pool.
struct DevBuf {
    void *p = nullptr; size_t cap = 0;
    void reserve(size_t bytes) {
        if (bytes <= cap) return;
        release();
        size_t want = bytes + (bytes >> 3) + 256;
        p = dev_pool().get(want, cap);
    }
    void release() { if (p) dev_pool().put(p, cap); p = nullptr; cap = 0; }
    template <class T> T *as() const { return (T *)p; }
    ~DevBuf() { release(); }
};

struct BlobHeader {          // first bytes of the device image
    u64 magic, total_bytes;
    u64 primary, L2[5], seq_len;
    i64 l_pac;
    u64 n_occ, n_sa, n_text;
    i32 sa_shift, n_seqs;
    u64 off_occ, off_sa, off_text, off_coff, off_calt, off_names, names_bytes;
};
static const u64 BLOB_MAGIC = 0x3242303032424c42ull; // "BLB200B2" (B2: Occ symbols stored as two bit planes)

struct ContigMeta { std::string name, anno; i64 offset; i32 len, n_ambs; u32 gi; i32 is_alt; };
struct HoleMeta { i64 offset; i32 len; char amb; };

} // namespace b200

struct b200_index {
    // meta
    b200::u64 primary = 0, L2[5] = {0, 0, 0, 0, 0}, seq_len = 0;
    b200::i64 l_pac = 0;
    std::vector<b200::ContigMeta> contigs;
    std::vector<b200::HoleMeta> holes;
    unsigned seed = 11;
    // host copy in bwa layout (optional)
    bool has_host = false;
    std::vector<uint32_t> h_bwt;
    int sa_intv = 32;
    std::vector<uint64_t> h_sa;
    std::vector<uint8_t> h_pac;
    std::vector<b200_contig_t> view_contigs;
    // device image
    void *d_blob = nullptr; b200::i64 blob_bytes = 0; bool owns_blob = false;
    b200::DevIndex dev;
    // prefix-chain table of the seeding kernel (seed2.cuh) and the verdict of the text proof: derived from the image, built on first use, never part of the blob
    mutable void *d_seedtab = nullptr; mutable int seedtab_K = -1, seed_text_ok = 0; mutable std::mutex seedtab_mu;
    ~b200_index();
};

namespace b200 {

// index.cu
BlobHeader plan_blob(u64 seq_len, i64 l_pac, int sa_shift, const std::vector<ContigMeta> &contigs);
void bind_blob(b200_index *idx, void *d_blob, const BlobHeader &h);     // fills idx->dev from the header
void upload_blob_meta(void *d_blob, const BlobHeader &h, const std::vector<ContigMeta> &contigs, cudaStream_t st);
int pick_sa_shift(u64 seq_len, int max_shift);
void image_from_bwa_arrays(b200_index *idx);            // host bwa-layout arrays -> device image (conversion kernels)
void host_copy_from_image(b200_index *idx);             // device image -> host bwa-layout arrays (bwt, sa at intv 32)

// index_build.cu: suffix sort of the forward+reverse text already in the blob; fills occ, sa, primary, L2.
void build_fm_index_device(b200_index *idx, const BlobHeader &h);

} // namespace b200
