// engine.cuh -- internal declarations shared by engine.cu, index.cu and index_build.cu
#pragma once
#include <cuda_runtime.h>
#include <string>
#include <vector>
#include <stdexcept>
#include "pipeline.cuh"
#include "hostutil.h"

namespace b200 {

void set_error(const std::string &msg);
int fail(int code, const std::string &msg);

struct CudaError : std::runtime_error { using std::runtime_error::runtime_error; };

#define CU_CHECK(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) \
    throw b200::CudaError(std::string(#expr) + ": " + cudaGetErrorString(e__) + " (" __FILE__ ":" + std::to_string(__LINE__) + ")"); } while (0)

// A cudaMalloc'd buffer that grows on demand (never shrinks).
struct DevBuf {
    void *p = nullptr; size_t cap = 0;
    void reserve(size_t bytes) {
        if (bytes <= cap) return;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        size_t want = bytes + (bytes >> 3) + 256;
        CU_CHECK(cudaMalloc(&p, want)); cap = want;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
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
