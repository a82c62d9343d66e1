// dist.cu -- multi-GPU plumbing of the aligner: one process per GPU, the FM-index image replicated by ONE
// NCCL broadcast (the reference's analogue of shipping bwa_idx2mem's single block, bwa/bwa.c:362-401), the read batch
// sharded by contiguous read index and scattered over NVLink.  There is no collective in the data path: after these two
// calls every rank runs b200_mem_align_batch on its own shard.
//
// NCCL is bound at run time (dlopen of libnccl.so.2): the library carries no link-time dependency on it, and a process
// that already loaded a NCCL (torch.distributed) shares that copy.
#include <dlfcn.h>
#include <cstring>
#include <string>
#include <vector>
#include "engine.cuh"

using namespace b200;

namespace {

typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclChar = 0, ncclUint8 = 1 };

struct Nccl {
    void *h = nullptr;
    int (*GetUniqueId)(ncclUniqueId *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*Broadcast)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    std::string err;
    bool load()
    {
        if (h) return true;
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *n : names) { h = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (h) break; }
        if (!h) { err = std::string("cannot load NCCL: ") + dlerror(); return false; }
#define B200_SYM(field, name) field = (decltype(field))dlsym(h, name); if (!field) { err = std::string("NCCL symbol missing: ") + name; h = nullptr; return false; }
        B200_SYM(GetUniqueId, "ncclGetUniqueId") B200_SYM(CommInitRank, "ncclCommInitRank") B200_SYM(CommDestroy, "ncclCommDestroy")
        B200_SYM(Broadcast, "ncclBroadcast") B200_SYM(Send, "ncclSend") B200_SYM(Recv, "ncclRecv")
        B200_SYM(GroupStart, "ncclGroupStart") B200_SYM(GroupEnd, "ncclGroupEnd") B200_SYM(GetErrorString, "ncclGetErrorString")
#undef B200_SYM
        return true;
    }
};
Nccl &nccl() { static Nccl n; return n; }

#define NCCL_CHECK(expr) do { int r__ = (expr); if (r__ != ncclSuccess) \
    throw std::runtime_error(std::string(#expr) + ": " + nccl().GetErrorString(r__)); } while (0)

} // namespace

struct b200_comm {
    ncclComm_t comm = nullptr; int rank = 0, world = 1; cudaStream_t st = nullptr;
    float last_bcast_ms = 0.f;
};

extern "C" {

void b200_shard_bounds(int64_t n_total, int world, int rank, int64_t *beg, int64_t *end)
{
    const int64_t q = n_total / world, r = n_total % world;
    const int64_t b = rank * q + (rank < r ? rank : r);
    if (beg) *beg = b;
    if (end) *end = b + q + (rank < r ? 1 : 0);
}

int b200_comm_unique_id(char id[128])
{
    if (!id) return fail(B200_ERR_ARG, "bad argument");
    if (!nccl().load()) return fail(B200_ERR_CUDA, nccl().err);
    ncclUniqueId u;
    int r = nccl().GetUniqueId(&u);
    if (r != ncclSuccess) return fail(B200_ERR_CUDA, nccl().GetErrorString(r));
    memcpy(id, u.internal, 128);
    return B200_OK;
}

int b200_comm_init(const char id[128], int rank, int world, b200_comm_t **out)
{
    if (!id || !out || world < 1 || rank < 0 || rank >= world) return fail(B200_ERR_ARG, "bad argument");
    *out = nullptr;
    if (!nccl().load()) return fail(B200_ERR_CUDA, nccl().err);
    b200_comm *c = new b200_comm;
    try {
        ncclUniqueId u; memcpy(u.internal, id, 128);
        c->rank = rank; c->world = world;
        CU_CHECK(cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking));
        NCCL_CHECK(nccl().CommInitRank(&c->comm, world, u, rank));
    } catch (const std::exception &e) { delete c; return fail(B200_ERR_CUDA, e.what()); }
    *out = c;
    return B200_OK;
}

void b200_comm_destroy(b200_comm_t *c)
{
    if (!c) return;
    if (c->comm) nccl().CommDestroy(c->comm);
    if (c->st) cudaStreamDestroy(c->st);
    delete c;
}

float b200_comm_last_bcast_ms(const b200_comm_t *c) { return c ? c->last_bcast_ms : 0.f; }

// One NCCL broadcast of the index image.  root: `root_idx` is its index, returned as *out unchanged.  Other ranks: a
// device buffer of the image's size is allocated, filled by the broadcast and attached; *out owns it.
int b200_index_bcast(b200_comm_t *c, const b200_index_t *root_idx, int root, b200_index_t **out)
{
    if (!c || !out || root < 0 || root >= c->world || (c->rank == root && !root_idx)) return fail(B200_ERR_ARG, "bad argument");
    *out = nullptr;
    try {
        // the image size travels first (8 bytes, same communicator)
        DevBuf dn; dn.reserve(64);
        long long nb = c->rank == root ? (long long)root_idx->blob_bytes : 0;
        if (c->rank == root) CU_CHECK(cudaMemcpyAsync(dn.p, &nb, 8, cudaMemcpyHostToDevice, c->st));
        NCCL_CHECK(nccl().Broadcast(dn.p, dn.p, 8, ncclUint8, root, c->comm, c->st));
        CU_CHECK(cudaMemcpyAsync(&nb, dn.p, 8, cudaMemcpyDeviceToHost, c->st));
        CU_CHECK(cudaStreamSynchronize(c->st));
        if (nb <= 0) return fail(B200_ERR_ARG, "empty index image");
        void *blob = nullptr;
        if (c->rank == root) blob = root_idx->d_blob;
        else if (cudaMalloc(&blob, (size_t)nb) != cudaSuccess) { cudaGetLastError(); dev_pool().trim(0); CU_CHECK(cudaMalloc(&blob, (size_t)nb)); }
        cudaEvent_t e0, e1; CU_CHECK(cudaEventCreate(&e0)); CU_CHECK(cudaEventCreate(&e1));
        CU_CHECK(cudaEventRecord(e0, c->st));
        NCCL_CHECK(nccl().Broadcast(blob, blob, (size_t)nb, ncclUint8, root, c->comm, c->st));
        CU_CHECK(cudaEventRecord(e1, c->st));
        CU_CHECK(cudaStreamSynchronize(c->st));
        cudaEventElapsedTime(&c->last_bcast_ms, e0, e1);
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        if (c->rank == root) { *out = const_cast<b200_index_t *>(root_idx); return B200_OK; }
        b200_index_t *idx = nullptr;
        int rc = b200_index_attach_blob(blob, nb, &idx);
        if (rc != B200_OK) { cudaFree(blob); return rc; }
        idx->owns_blob = true;
        *out = idx;
    } catch (const std::exception &e) { return fail(B200_ERR_CUDA, e.what()); }
    return B200_OK;
}

// Scatter of a batch of fixed-length reads held by `root` (host memory, n_total * read_len bytes): rank r receives the bases
// of reads [b_r, e_r) = b200_shard_bounds(n_total, world, r) into `shard` (host memory, (e_r - b_r) * read_len bytes).
// The bytes travel host -> root GPU -> NVLink -> peer GPU -> host, the path a device-resident consumer would use.
int b200_reads_scatter(b200_comm_t *c, int root, int64_t n_total, int read_len, const char *seqs, char *shard)
{
    if (!c || root < 0 || root >= c->world || n_total < 0 || read_len <= 0 || (c->rank == root && n_total && !seqs) || !shard) return fail(B200_ERR_ARG, "bad argument");
    try {
        int64_t b, e;
        b200_shard_bounds(n_total, c->world, c->rank, &b, &e);
        const size_t mine = (size_t)(e - b) * read_len;
        if (c->rank == root) {
            DevBuf all; all.reserve((size_t)n_total * read_len + 64);
            CU_CHECK(cudaMemcpyAsync(all.p, seqs, (size_t)n_total * read_len, cudaMemcpyHostToDevice, c->st));
            NCCL_CHECK(nccl().GroupStart());
            for (int r = 0; r < c->world; ++r) {
                if (r == root) continue;
                int64_t rb, re; b200_shard_bounds(n_total, c->world, r, &rb, &re);
                if (re > rb) NCCL_CHECK(nccl().Send(all.as<u8>() + (size_t)rb * read_len, (size_t)(re - rb) * read_len, ncclUint8, r, c->comm, c->st));
            }
            NCCL_CHECK(nccl().GroupEnd());
            CU_CHECK(cudaStreamSynchronize(c->st));
            if (mine) memcpy(shard, seqs + (size_t)b * read_len, mine);
        } else {
            DevBuf part; part.reserve(mine + 64);
            if (mine) NCCL_CHECK(nccl().Recv(part.p, mine, ncclUint8, root, c->comm, c->st));
            if (mine) CU_CHECK(cudaMemcpyAsync(shard, part.p, mine, cudaMemcpyDeviceToHost, c->st));
            CU_CHECK(cudaStreamSynchronize(c->st));
        }
    } catch (const std::exception &e) { return fail(B200_ERR_CUDA, e.what()); }
    return B200_OK;
}

} // extern "C"
