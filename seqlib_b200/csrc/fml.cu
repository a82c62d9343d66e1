// fml.cu -- fermi-lite's BFC stage as read-batched kernels: k-mer counting, error correction, unique-k-mer filter.
//
//   fml_count / bfc_ch_insert     (fermi-lite/bfc.c:66-99, htab.c:60-83)   -> k_count_emit + radix sort + k_runs/k_starts/k_table_build
//   bfc_ch_hist                   (fermi-lite/htab.c:104-127)              -> histogram accumulated by k_table_build
//   fml_correct_core              (fermi-lite/bfc.c:513-553)               -> correct_flat()
//   worker_ec -> bfc_ec1          (fermi-lite/bfc.c:401-511)               -> k_ec / k_flt (per-read code in bfc.cuh)
//
// The reference inserts k-mers one by one into 2^l_pre khash tables with saturating 8-bit/6-bit counters.  Saturating
// adds commute, so the device counts by sorting the (sub-table, stored-key) pairs and measuring run lengths, then
// builds ONE open-addressing table of 16-byte slots for the lookups of the correction kernel (each probe = one 16-byte
// load = one 32-byte HBM sector).
#include <cstring>
#include <cstdlib>
#include <ctime>
#include <thread>
#include <atomic>
#include <mutex>
#include <string>
#include <vector>
#include <cub/cub.cuh>
#include "engine.cuh"
#include "bfc.cuh"
#include "fml_host.h"
#include "fmd_dev.h"
#include "unitig.cuh"
#include "utg_walk.h"
#include "mag_host.h"

using namespace b200;

struct b200_kmer_table {
    int k = 0, l_pre = 0, q = 20;
    DevBuf slots; u64 cap = 0;
    u64 n_kmers = 0, n_distinct = 0;
    uint64_t hist[256], hist_high[64];
    CountTable view() const { CountTable t; t.slots = slots.as<CountSlot>(); t.mask = cap - 1; t.k = k; t.l_pre = l_pre; return t; }
};

struct b200_fmd {             // device-resident FMD-index (rld_t)
    b200::FmdDevice F;
};

struct b200_utgs {            // fml_utg_t[] as one handle
    std::vector<b200::MagUtg> utg;
    std::vector<b200_utg_t> view;
    std::vector<std::vector<b200_utg_ovlp_t>> ovlp;
};

namespace b200 {

static thread_local b200_fml_stats_t g_fml_stats;

struct ReadPool {            // device view of a flat read batch
    char *seq, *qual; const i64 *off; i64 n;
};

// ------------------------------------------------------------------------------------------------ counting
// pass 1: k-mers per read (exclusive-summed into the record offsets)
__global__ void __launch_bounds__(256) k_count_n(ReadPool R, int k, u32 *n_out)
{
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R.n) return;
    i64 b = R.off[i];
    n_out[i] = (u32)count_read_kmers(k, 0, 0, R.seq + b, nullptr, (int)(R.off[i + 1] - b), nullptr, nullptr);
}

// pass 2: one (lo, hi) record per k-mer
__global__ void __launch_bounds__(256) k_count_emit(ReadPool R, int k, int l_pre, int q, const u64 *rec_off, u64 *lo, u32 *hi)
{
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R.n) return;
    i64 b = R.off[i];
    u64 o = rec_off[i];
    count_read_kmers(k, l_pre, q, R.seq + b, R.qual ? R.qual + b : nullptr, (int)(R.off[i + 1] - b), lo + o, hi + o);
}

// run heads of the sorted records: low word = 1 at the first record of a key, high word = the record's is_high flag
__global__ void __launch_bounds__(256) k_runs(const u64 *lo, const u32 *hi, u64 n, u64 *flags)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u32 h = hi[i];
    bool head = i == 0 || lo[i] != lo[i - 1] || ((h ^ hi[i - 1]) & 0x7fffffffu) != 0;
    flags[i] = (u64)head | (u64)(h >> 31) << 32;
}

// scan[i] = inclusive sums of flags; a head writes its index to starts[rank]
__global__ void __launch_bounds__(256) k_starts(const u64 *lo, const u32 *hi, const u64 *scan, u64 n, u32 *starts)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool head = i == 0 || lo[i] != lo[i - 1] || ((hi[i] ^ hi[i - 1]) & 0x7fffffffu) != 0;
    if (head) starts[(u32)scan[i] - 1] = (u32)i;
}

// one thread per distinct key: saturated counts (bfc_ch_insert), slot claim, histogram (bfc_ch_hist)
__global__ void __launch_bounds__(256) k_table_build(const u64 *lo, const u32 *hi, const u64 *scan, const u32 *starts, u64 n, u64 n_distinct,
                                                     CountSlot *slots, u64 mask, unsigned long long *hist /* 256 + 64 */)
{
    __shared__ u32 sh[320];
    for (int t = threadIdx.x; t < 320; t += blockDim.x) sh[t] = 0;
    __syncthreads();
    u64 u = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (u < n_distinct) {
        u64 s = starts[u], e = u + 1 < n_distinct ? starts[u + 1] : n;
        u64 tot = e - s;
        u64 high = (scan[e - 1] >> 32) - (s ? scan[s - 1] >> 32 : 0);
        // first insert: count 1 (+ high 1); later inserts saturate at 255 / 63 (fermi-lite/htab.c:74-80)
        u32 cnt = tot > 255 ? 255u : (u32)tot;
        u32 hc = high > 63 ? 63u : (u32)high;
        // the reference only bumps the high counter of an existing key while... (no coupling to the total counter)
        KmerKey key; key.lo = lo[s]; key.hi = hi[s] & 0x7fffffffu;
        u64 w1 = key.hi << 16 | 1ull << 15 | (u64)(hc << 8 | cnt);
        u64 j = count_slot_hash(key) & mask;
        for (;;) {
            unsigned long long old = atomicCAS((unsigned long long *)&slots[j].w1, 0ull, (unsigned long long)w1);
            if (old == 0) { slots[j].w0 = key.lo; break; }
            j = (j + 1) & mask;
        }
        atomicAdd(&sh[cnt], 1u);
        atomicAdd(&sh[256 + hc], 1u);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < 320; t += blockDim.x) if (sh[t]) atomicAdd(&hist[t], (unsigned long long)sh[t]);
}

// ------------------------------------------------------------------------------------------------ correction
struct EcArgs {
    ReadPool R;
    CountTable tab;
    BfcOpt opt;
    int mode;
    u8 *scratch; size_t scratch_stride; int maxlen, heap_cap, stack_cap;
    const u32 *todo; i64 n_todo;       // null: all reads
    u8 *codes;                         // per read: ec_code
    unsigned long long *work, *lookups;
};

struct CountingTable {       // CountTable + a probe counter (roofline accounting)
    CountTable t; mutable unsigned long long n;
    __device__ int kmer_occ(const Kmer4 &z) const { ++n; return t.kmer_occ(z); }
};

template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_ec(const __grid_constant__ EcArgs A)
{
    const int lane = threadIdx.x & 31;
    EcScratch e;
    u8 *mine = A.scratch + (size_t)(blockIdx.x * blockDim.x + threadIdx.x) * A.scratch_stride;
    CountingTable tab; tab.t = A.tab; tab.n = 0;
    const i64 n = A.todo ? A.n_todo : A.R.n;
    for (;;) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(A.work, 32ull);
        base = __shfl_sync(0xffffffffu, base, 0);
        if ((i64)base >= n) break;
        i64 t = (i64)base + lane;
        if (t < n) {
            i64 i = A.todo ? (i64)A.todo[t] : t;
            i64 b = A.R.off[i];
            int len = (int)(A.R.off[i + 1] - b);
            int code = ECCODE_SCRATCH;
            if (len <= A.maxlen) {
                ec_scratch_bind(e, mine, A.maxlen, A.heap_cap, A.stack_cap);
                code = ec1(A.opt, tab, A.mode, A.R.seq + b, A.R.qual ? A.R.qual + b : nullptr, len, e);
            }
            A.codes[i] = (u8)code;
        }
        __syncwarp();
    }
    unsigned long long x = tab.n;
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) atomicAdd(A.lookups, x);
}

// Scheduling key of a read for k_ec: reads whose k-mers are all solid walk the "fixed" fast path (one probe per base), every
// non-solid k-mer opens a search.  Warps that get reads of similar difficulty diverge less, and the hardest reads go first so
// they do not form the tail of the launch.  (Purely an ordering: k_ec recomputes everything it needs.)
__global__ void __launch_bounds__(256) k_ec_difficulty(ReadPool R, CountTable tab, int min_cov, u32 *key, u32 *idx)
{
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R.n) return;
    i64 b = R.off[i];
    int len = (int)(R.off[i + 1] - b), l = 0, bad = 0;
    const char *s = R.seq + b;
    Kmer4 x; kmer_clear(x);
    for (int j = 0; j < len; ++j) {
        int c = nt6m1((unsigned char)s[j]);
        if (c < 4) {
            kmer_append(tab.k, x.x, c);
            if (++l >= tab.k) { int r = tab.kmer_occ(x); if (r < 0 || (r & 0xff) < min_cov) ++bad; }
        } else { l = 0; kmer_clear(x); bad += tab.k; }
    }
    key[i] = ~(u32)bad;          // ascending sort -> most non-solid k-mers first
    idx[i] = (u32)i;
}

__global__ void __launch_bounds__(256) k_flt(ReadPool R, CountTable tab, BfcOpt opt, i32 *len_out, unsigned long long *lookups)
{
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long nl = 0;
    if (i < R.n) {
        i64 b = R.off[i];
        int len = (int)(R.off[i + 1] - b);
        len_out[i] = len > 0 ? fltuniq1(opt, tab, R.seq + b, R.qual ? R.qual + b : nullptr, len) : 0;
        nl = len >= opt.k ? len - opt.k + 1 : 0;
    }
    for (int o = 16; o > 0; o >>= 1) nl += __shfl_xor_sync(0xffffffffu, nl, o);
    if ((threadIdx.x & 31) == 0 && nl) atomicAdd(lookups, nl);
}

__global__ void __launch_bounds__(256) k_lookup(CountTable tab, i64 n, const char *kmers, i32 *occ)
{
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Kmer4 x; kmer_clear(x);
    int ok = 1;
    for (int j = 0; j < tab.k; ++j) {
        int c = nt6m1((unsigned char)kmers[i * tab.k + j]);
        if (c > 3) { ok = 0; break; }
        kmer_append(tab.k, x.x, c);
    }
    occ[i] = ok ? tab.kmer_occ(x) : -1;
}

// ------------------------------------------------------------------------------------------------ unitig records
struct UtgArgs {
    FmdIndex e; int min_match;
    UtgNode *node;
    u8 *seq; UtgNei *nei; UtgMark *mark; u64 seq_cap, nei_cap, mark_cap;
    unsigned long long *ctr;           // [0] work, [1] seq bytes, [2] nei entries, [3] mark entries, [4] pool overflow, [5] rank queries
    const u32 *todo; u64 n_todo;
    u8 *scratch; size_t stride; int cap, s_cap, tmark_cap;
};

// copies what utg_node left in the scratch slot into the pools (bump allocation) and stores the record
__device__ __forceinline__ void utg_emit_node(const UtgArgs &A, u64 x, const UtgScratch &S, UtgNode &N)
{
    N.seq_off = N.nei_off = N.mark_off = 0;
    if (!(N.flags & UTG_OVERFLOW)) {
        u64 ns = (u64)(N.len + N.ext_len), nn = (u64)N.n_nei, nm = (u64)(N.n_mark_r + N.n_mark_c);
        u64 so = ns ? atomicAdd(A.ctr + 1, (unsigned long long)ns) : 0;
        u64 no = nn ? atomicAdd(A.ctr + 2, (unsigned long long)nn) : 0;
        u64 mo = nm ? atomicAdd(A.ctr + 3, (unsigned long long)nm) : 0;
        if (so + ns > A.seq_cap || no + nn > A.nei_cap || mo + nm > A.mark_cap) atomicAdd(A.ctr + 4, 1ull);
        else {
            N.seq_off = so; N.nei_off = no; N.mark_off = mo;
            for (u64 i = 0; i < ns; ++i) A.seq[so + i] = S.s[i];
            for (u64 i = 0; i < nn; ++i) { UtgNei o; o.x0 = S.nei[i].x[0]; o.x1 = S.nei[i].x[1]; o.x2 = S.nei[i].x[2]; o.ovlp = (i64)S.nei[i].info; A.nei[no + i] = o; }
            for (u64 i = 0; i < nm; ++i) A.mark[mo + i] = S.mark[i];
        }
    }
    A.node[x] = N;
}

// one string per thread: fm6_retrieve + fm6_get_nei + check_left (unitig.cuh), results bump-allocated into the pools
template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_utg_nodes(const __grid_constant__ UtgArgs A)
{
    const int lane = threadIdx.x & 31;
    u8 *mine = A.scratch + (size_t)(blockIdx.x * blockDim.x + threadIdx.x) * A.stride;
    UtgScratch S;
    utg_scratch_bind(S, mine, A.cap, A.s_cap, A.tmark_cap);
    const u64 n = A.todo ? A.n_todo : A.e.n_str;
    for (;;) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(A.ctr, 32ull);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= n) break;
        u64 t = base + lane;
        if (t < n) {
            u64 x = A.todo ? (u64)A.todo[t] : t;
            UtgNode N;
            utg_node(A.e, A.min_match, x, S, N, UtgNoCoop());
            utg_emit_node(A, x, S, N);
        }
        __syncwarp();
    }
}

// ---- G lanes per string, for inputs too small to fill the machine with one string per thread (window-sized assemblies):
// the time of k_utg_nodes is then the latency of ONE string (~9 ms).  Lane 0 of a group (the master) runs the reference-
// shaped code; the extensions of all overlap intervals of a round -- independent of each other, ~120 per round at 150x
// coverage -- are shared out over the G lanes through a shared-memory mailbox (two-phase loops of unitig.cuh) and consumed
// in interval order by the master.  Per-string latency drops ~4x; with 10^5+ strings the one-thread-per-string kernel
// keeps 10x more strings in flight and wins (measured 496 ms vs 1 391 ms at 2*10^6 strings).
struct UtgMailbox { int cmd, pn, arg; const FmdIntv *prev; const i32 *cat; UtgExt *R; };
enum { UTG_CMD_FWD = 1, UTG_CMD_BWD = 2, UTG_CMD_EXIT = 3 };

template <int G>
struct UtgGroupCoop {
    static constexpr bool kTwoPhase = true;
    volatile UtgMailbox *mb; unsigned mask; int gl;
    __device__ __forceinline__ static void share_fwd(const FmdIndex &e, const FmdIntv *prev, int pn, const i32 *cat, bool later, UtgExt *R, int gl)
    {
        for (int j = gl; j < pn; j += G) if (cat[j] >= 0) utg_ext_forward(e, prev[j], later, R[j]);
    }
    __device__ __forceinline__ static void share_bwd(const FmdIndex &e, const FmdIntv *prev, int pn, int sym, UtgExt *R, int gl)
    {
        for (int j = gl; j < pn; j += G) utg_ext_backward(e, prev[j], sym, R[j]);
    }
    // master side
    __device__ __forceinline__ void forward(const FmdIndex &e, const FmdIntv *prev, int pn, const i32 *cat, bool later, UtgExt *R) const
    {
        if (pn <= 2) { for (int j = 0; j < pn; ++j) if (cat[j] >= 0) utg_ext_forward(e, prev[j], later, R[j]); return; }
        mb->cmd = UTG_CMD_FWD; mb->pn = pn; mb->arg = later; mb->prev = prev; mb->cat = cat; mb->R = R;
        __syncwarp(mask);
        share_fwd(e, prev, pn, cat, later, R, 0);
        __syncwarp(mask);
    }
    __device__ __forceinline__ void backward(const FmdIndex &e, const FmdIntv *prev, int pn, int sym, UtgExt *R) const
    {
        if (pn <= 2) { for (int j = 0; j < pn; ++j) utg_ext_backward(e, prev[j], sym, R[j]); return; }
        mb->cmd = UTG_CMD_BWD; mb->pn = pn; mb->arg = sym; mb->prev = prev; mb->cat = nullptr; mb->R = R;
        __syncwarp(mask);
        share_bwd(e, prev, pn, sym, R, 0);
        __syncwarp(mask);
    }
    // helper side: serve commands until the master posts EXIT
    __device__ __forceinline__ void serve(const FmdIndex &e) const
    {
        for (;;) {
            __syncwarp(mask);
            int cmd = mb->cmd;
            if (cmd == UTG_CMD_EXIT) break;
            const FmdIntv *prev = mb->prev; int pn = mb->pn, arg = mb->arg; UtgExt *R = mb->R;
            if (cmd == UTG_CMD_FWD) share_fwd(e, prev, pn, mb->cat, arg != 0, R, gl);
            else share_bwd(e, prev, pn, arg, R, gl);
            __syncwarp(mask);
        }
    }
};

template <int G>
__global__ void __launch_bounds__(128) k_utg_nodes_group(const __grid_constant__ UtgArgs A)
{
    __shared__ UtgMailbox mbox[128 / G];
    const int lane = threadIdx.x & 31, gl = lane % G, gib = threadIdx.x / G;
    UtgGroupCoop<G> coop;
    coop.mb = &mbox[gib]; coop.gl = gl;
    coop.mask = G == 32 ? 0xffffffffu : (((1u << G) - 1u) << (lane - gl));
    if (gl != 0) { coop.serve(A.e); return; }
    const u64 group_id = ((u64)blockIdx.x * blockDim.x + threadIdx.x) / G;
    UtgScratch S;
    utg_scratch_bind(S, A.scratch + group_id * A.stride, A.cap, A.s_cap, A.tmark_cap, true);
    const u64 n = A.todo ? A.n_todo : A.e.n_str;
    for (;;) {
        u64 t = atomicAdd(A.ctr, 1ull);
        if (t >= n) break;
        u64 x = A.todo ? (u64)A.todo[t] : t;
        UtgNode N;
        utg_node(A.e, A.min_match, x, S, N, coop);
        utg_emit_node(A, x, S, N);
    }
    mbox[gib].cmd = UTG_CMD_EXIT;
    __syncwarp(coop.mask);
}

__global__ void __launch_bounds__(256) k_fmd_rank(FmdIndex e, i64 n, const u64 *q, u64 *ranks, i32 *sym)
{
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u64 ok[6];
    sym[i] = fmd_rank1a(e, q[i], ok);
    for (int c = 0; c < 6; ++c) ranks[6 * i + c] = ok[c];
}

// ------------------------------------------------------------------------------------------------ host side
struct FmlEngine {
    cudaStream_t st = nullptr;
    cudaEvent_t ev[6];
    int sm_count = 0;
    DevBuf d_seq, d_qual, d_off, d_cnt, d_recoff, d_lo[2], d_hi[2], d_flags, d_starts, d_tmp, d_hist, d_len, d_codes, d_scratch, d_ctr, d_todo, d_order[2];
    bool ready = false;
    ~FmlEngine()        // worker threads of b200_fml_assemble_windows come and go: give the stream and events back
    {
        if (!ready) return;
        cudaStreamDestroy(st);
        for (auto &e : ev) cudaEventDestroy(e);
        cudaGetLastError();
    }
    void init()
    {
        if (ready) return;
        int dev = 0;
        CU_CHECK(cudaGetDevice(&dev));
        cudaDeviceProp p;
        CU_CHECK(cudaGetDeviceProperties(&p, dev));
        if (p.major < 10) throw CudaError("libseqlib_b200 needs an sm_100 device (found sm_" + std::to_string(p.major * 10 + p.minor) + ")");
        sm_count = p.multiProcessorCount;
        CU_CHECK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        for (auto &e : ev) CU_CHECK(cudaEventCreate(&e));
        ready = true;
    }
};
static FmlEngine &fml_engine() { static thread_local FmlEngine e; e.init(); return e; }

static int clamp_l_pre(int k, int l_pre) { return bfc_l_pre(k, l_pre); }

// device reads already uploaded in E.d_seq/d_qual/d_off
static void count_on_device(FmlEngine &E, ReadPool R, i64 tot_len, int k, int q, int l_pre_in, b200_kmer_table *tab)
{
    if (k < 1 || k > 63) throw std::invalid_argument("k-mer length must be in [1, 63]");
    int l_pre = clamp_l_pre(k, l_pre_in);
    if (l_pre < 0) l_pre = 0;
    if (k <= 32 ? 2 * k < l_pre : k - l_pre >= 50) throw std::invalid_argument("unsupported (k, l_pre) pair");
    tab->k = k; tab->l_pre = l_pre; tab->q = q;
    cudaStream_t st = E.st;
    const i64 n = R.n;
    int &nl = g_fml_stats.n_launches;
    // records per read
    E.d_cnt.reserve((size_t)(n + 1) * 4); E.d_recoff.reserve((size_t)(n + 1) * 8);
    u64 n_rec = 0;
    if (n > 0) {
        k_count_n<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(R, k, E.d_cnt.as<u32>()); ++nl;
        size_t tb = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tb, E.d_cnt.as<u32>(), E.d_recoff.as<u64>(), (int)n + 1, st);
        E.d_tmp.reserve(tb);
        CU_CHECK(cudaMemsetAsync(E.d_cnt.as<u32>() + n, 0, 4, st));
        cub::DeviceScan::ExclusiveSum(E.d_tmp.p, tb, E.d_cnt.as<u32>(), E.d_recoff.as<u64>(), (int)n + 1, st); ++nl;
        CU_CHECK(cudaMemcpyAsync(&n_rec, E.d_recoff.as<u64>() + n, 8, cudaMemcpyDeviceToHost, st));
        CU_CHECK(cudaStreamSynchronize(st));
    }
    if (n_rec >= (1ull << 31)) throw std::length_error("more than 2^31 k-mers in one batch");
    tab->n_kmers = n_rec;
    memset(tab->hist, 0, sizeof(tab->hist)); memset(tab->hist_high, 0, sizeof(tab->hist_high));
    u64 n_distinct = 0;
    if (n_rec > 0) {
        for (int b = 0; b < 2; ++b) { E.d_lo[b].reserve(n_rec * 8); E.d_hi[b].reserve(n_rec * 4); }
        k_count_emit<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(R, k, l_pre, q, E.d_recoff.as<u64>(), E.d_lo[0].as<u64>(), E.d_hi[0].as<u32>()); ++nl;
        // sort by (hi, lo): LSD, stable -- lo first, then the sub-table bits of hi (only k > 32 has any)
        int lo_bits = k <= 32 ? 2 * k : 50;
        int hi_bits = k <= 32 ? 0 : l_pre;
        size_t tb = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tb, E.d_lo[0].as<u64>(), E.d_lo[1].as<u64>(), E.d_hi[0].as<u32>(), E.d_hi[1].as<u32>(), (int)n_rec, 0, lo_bits, st);
        size_t tb2 = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tb2, E.d_hi[1].as<u32>(), E.d_hi[0].as<u32>(), E.d_lo[1].as<u64>(), E.d_lo[0].as<u64>(), (int)n_rec, 0, hi_bits > 0 ? hi_bits : 1, st);
        size_t tb3 = 0;
        E.d_flags.reserve(n_rec * 8);
        cub::DeviceScan::InclusiveSum(nullptr, tb3, E.d_flags.as<u64>(), E.d_flags.as<u64>(), (int)n_rec, st);
        E.d_tmp.reserve(std::max(tb, std::max(tb2, tb3)));
        cub::DeviceRadixSort::SortPairs(E.d_tmp.p, tb, E.d_lo[0].as<u64>(), E.d_lo[1].as<u64>(), E.d_hi[0].as<u32>(), E.d_hi[1].as<u32>(), (int)n_rec, 0, lo_bits, st); ++nl;
        int cur = 1;
        if (hi_bits > 0) {
            cub::DeviceRadixSort::SortPairs(E.d_tmp.p, tb2, E.d_hi[1].as<u32>(), E.d_hi[0].as<u32>(), E.d_lo[1].as<u64>(), E.d_lo[0].as<u64>(), (int)n_rec, 0, hi_bits, st); ++nl;
            cur = 0;
        }
        const u64 *lo = E.d_lo[cur].as<u64>(); const u32 *hi = E.d_hi[cur].as<u32>();
        unsigned gb = (unsigned)((n_rec + 255) / 256);
        k_runs<<<gb, 256, 0, st>>>(lo, hi, n_rec, E.d_flags.as<u64>()); ++nl;
        cub::DeviceScan::InclusiveSum(E.d_tmp.p, tb3, E.d_flags.as<u64>(), E.d_flags.as<u64>(), (int)n_rec, st); ++nl;
        u64 last = 0;
        CU_CHECK(cudaMemcpyAsync(&last, E.d_flags.as<u64>() + (n_rec - 1), 8, cudaMemcpyDeviceToHost, st));
        CU_CHECK(cudaStreamSynchronize(st));
        n_distinct = (u32)last;
        E.d_starts.reserve(n_distinct * 4);
        k_starts<<<gb, 256, 0, st>>>(lo, hi, E.d_flags.as<u64>(), n_rec, E.d_starts.as<u32>()); ++nl;
        u64 cap = 1024;
        while (cap < 2 * n_distinct) cap <<= 1;
        tab->cap = cap;
        tab->slots.reserve(cap * sizeof(CountSlot));
        CU_CHECK(cudaMemsetAsync(tab->slots.p, 0, cap * sizeof(CountSlot), st));
        E.d_hist.reserve(320 * 8);
        CU_CHECK(cudaMemsetAsync(E.d_hist.p, 0, 320 * 8, st));
        k_table_build<<<(unsigned)((n_distinct + 255) / 256), 256, 0, st>>>(lo, hi, E.d_flags.as<u64>(), E.d_starts.as<u32>(), n_rec, n_distinct,
                                                                           tab->slots.as<CountSlot>(), cap - 1, E.d_hist.as<unsigned long long>()); ++nl;
        uint64_t h[320];
        CU_CHECK(cudaMemcpyAsync(h, E.d_hist.p, sizeof(h), cudaMemcpyDeviceToHost, st));
        CU_CHECK(cudaStreamSynchronize(st));
        CU_CHECK(cudaGetLastError());
        memcpy(tab->hist, h, 256 * 8); memcpy(tab->hist_high, h + 256, 64 * 8);
    } else {
        tab->cap = 1024;
        tab->slots.reserve(tab->cap * sizeof(CountSlot));
        CU_CHECK(cudaMemsetAsync(tab->slots.p, 0, tab->cap * sizeof(CountSlot), st));
        CU_CHECK(cudaStreamSynchronize(st));
    }
    tab->n_distinct = n_distinct;
    g_fml_stats.n_kmers = n_rec; g_fml_stats.n_distinct = n_distinct; g_fml_stats.table_bytes = tab->cap * sizeof(CountSlot);
    (void)tot_len;
}

static ReadPool upload_reads(FmlEngine &E, i64 n, const char *seqs, const char *quals, const i64 *off)
{
    i64 tot = n > 0 ? off[n] : 0;
    E.d_seq.reserve((size_t)tot + 16); E.d_off.reserve((size_t)(n + 1) * 8);
    if (tot) CU_CHECK(cudaMemcpyAsync(E.d_seq.p, seqs, (size_t)tot, cudaMemcpyHostToDevice, E.st));
    if (quals) { E.d_qual.reserve((size_t)tot + 16); if (tot) CU_CHECK(cudaMemcpyAsync(E.d_qual.p, quals, (size_t)tot, cudaMemcpyHostToDevice, E.st)); }
    CU_CHECK(cudaMemcpyAsync(E.d_off.p, off, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, E.st));
    ReadPool R; R.seq = E.d_seq.as<char>(); R.qual = quals ? E.d_qual.as<char>() : nullptr; R.off = E.d_off.as<i64>(); R.n = n;
    return R;
}

// the correction / filter kernels over uploaded reads; results copied back into the host pools
static void correct_on_device(FmlEngine &E, ReadPool R, const b200_kmer_table *tab, const BfcOpt &bo, int mode, int flt_uniq,
                              char *seqs, char *quals, const i64 *off, i32 *len_out)
{
    cudaStream_t st = E.st;
    const i64 n = R.n;
    int &nl = g_fml_stats.n_launches;
    if (n == 0) return;
    i64 tot = off[n];
    i64 maxlen = 0;
    for (i64 i = 0; i < n; ++i) maxlen = std::max(maxlen, off[i + 1] - off[i]);
    E.d_ctr.reserve(64);
    CU_CHECK(cudaMemsetAsync(E.d_ctr.p, 0, 64, st));
    unsigned long long *ctr = E.d_ctr.as<unsigned long long>();      // [0] work, [1] lookups, [2] spill work
    if (flt_uniq) {
        E.d_len.reserve((size_t)n * 4);
        k_flt<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(R, tab->view(), bo, E.d_len.as<i32>(), ctr + 1); ++nl;
        if (len_out) CU_CHECK(cudaMemcpyAsync(len_out, E.d_len.p, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    } else {
        E.d_codes.reserve((size_t)n);
        EcArgs A;
        A.R = R; A.tab = tab->view(); A.opt = bo; A.mode = mode;
        A.maxlen = (int)std::min<i64>(maxlen, 512); A.heap_cap = bo.max_heap + 8; A.stack_cap = 1536;
        A.scratch_stride = (ec_scratch_bytes(A.maxlen, A.heap_cap, A.stack_cap) + 15) & ~(size_t)15;
        const int threads = 128;
        int blocks = (int)std::min<i64>((n + threads - 1) / threads, (i64)E.sm_count * 8);       // >= the resident blocks of either variant
        E.d_scratch.reserve((size_t)blocks * threads * A.scratch_stride);
        A.scratch = E.d_scratch.as<u8>();
        A.todo = nullptr; A.n_todo = 0; A.codes = E.d_codes.as<u8>(); A.work = ctr; A.lookups = ctr + 1;
        if (n >= 4096 && n < (1ll << 31)) {       // order the reads by difficulty (see k_ec_difficulty)
            E.d_cnt.reserve((size_t)n * 4); E.d_starts.reserve((size_t)n * 4); E.d_order[0].reserve((size_t)n * 4); E.d_order[1].reserve((size_t)n * 4);
            k_ec_difficulty<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(R, A.tab, bo.min_cov, E.d_cnt.as<u32>(), E.d_order[0].as<u32>()); ++nl;
            size_t tb = 0;
            cub::DeviceRadixSort::SortPairs(nullptr, tb, E.d_cnt.as<u32>(), E.d_starts.as<u32>(), E.d_order[0].as<u32>(), E.d_order[1].as<u32>(), (int)n, 0, 32, st);
            E.d_tmp.reserve(tb);
            cub::DeviceRadixSort::SortPairs(E.d_tmp.p, tb, E.d_cnt.as<u32>(), E.d_starts.as<u32>(), E.d_order[0].as<u32>(), E.d_order[1].as<u32>(), (int)n, 0, 32, st); ++nl;
            A.todo = E.d_order[1].as<u32>(); A.n_todo = n;
        }
        static const int ec_minb = getenv("B200_EC_MINB") ? atoi(getenv("B200_EC_MINB")) : 5;
        if (ec_minb >= 8) k_ec<8><<<blocks, threads, 0, st>>>(A); else k_ec<5><<<blocks, threads, 0, st>>>(A);
        ++nl;
        std::vector<u8> codes((size_t)n);
        CU_CHECK(cudaMemcpyAsync(codes.data(), E.d_codes.p, (size_t)n, cudaMemcpyDeviceToHost, st));
        CU_CHECK(cudaStreamSynchronize(st));
        CU_CHECK(cudaGetLastError());
        // reads whose search outgrew the per-thread scratch: same kernel, few threads, large scratch
        std::vector<u32> todo;
        for (i64 i = 0; i < n; ++i) if (codes[i] == ECCODE_SCRATCH) todo.push_back((u32)i);
        g_fml_stats.n_spill = todo.size();
        for (int round = 0; !todo.empty(); ++round) {
            if (round > 4) throw std::length_error("error-correction search exceeds the largest scratch");
            EcArgs B = A;
            B.maxlen = (int)maxlen; B.stack_cap = 32768 << (3 * round);
            B.scratch_stride = (ec_scratch_bytes(B.maxlen, B.heap_cap, B.stack_cap) + 15) & ~(size_t)15;
            int sb = (int)std::min<size_t>((todo.size() + 31) / 32, 64);
            E.d_scratch.reserve((size_t)sb * 32 * B.scratch_stride);
            B.scratch = E.d_scratch.as<u8>();
            E.d_todo.reserve(todo.size() * 4);
            CU_CHECK(cudaMemcpyAsync(E.d_todo.p, todo.data(), todo.size() * 4, cudaMemcpyHostToDevice, st));
            CU_CHECK(cudaMemsetAsync(ctr + 2, 0, 8, st));
            B.todo = E.d_todo.as<u32>(); B.n_todo = (i64)todo.size(); B.work = ctr + 2;
            k_ec<5><<<sb, 32, 0, st>>>(B); ++nl;
            CU_CHECK(cudaMemcpyAsync(codes.data(), E.d_codes.p, (size_t)n, cudaMemcpyDeviceToHost, st));
            CU_CHECK(cudaStreamSynchronize(st));
            CU_CHECK(cudaGetLastError());
            std::vector<u32> next;
            for (u32 i : todo) if (codes[i] == ECCODE_SCRATCH) next.push_back(i);
            todo.swap(next);
        }
        for (i64 i = 0; i < n; ++i) ++g_fml_stats.ec_codes[codes[i] & 7];
        if (len_out) for (i64 i = 0; i < n; ++i) len_out[i] = (i32)(off[i + 1] - off[i]);
    }
    if (seqs) CU_CHECK(cudaMemcpyAsync(seqs, R.seq, (size_t)tot, cudaMemcpyDeviceToHost, st));      // null: results stay on the device
    if (seqs && quals) CU_CHECK(cudaMemcpyAsync(quals, R.qual, (size_t)tot, cudaMemcpyDeviceToHost, st));
    unsigned long long lk = 0;
    CU_CHECK(cudaMemcpyAsync(&lk, ctr + 1, 8, cudaMemcpyDeviceToHost, st));
    CU_CHECK(cudaStreamSynchronize(st));
    CU_CHECK(cudaGetLastError());
    g_fml_stats.n_lookups = lk;
}

static float ms_between(cudaEvent_t a, cudaEvent_t b) { float ms = 0; cudaEventElapsedTime(&ms, a, b); return ms; }

template <class F> static int guarded(F &&f)
{
    try { return f(); }
    catch (const CudaError &e) { return fail(B200_ERR_CUDA, e.what()); }
    catch (const std::bad_alloc &) { return fail(B200_ERR_NOMEM, "out of host memory"); }
    catch (const std::length_error &e) { return fail(B200_ERR_LIMIT, e.what()); }
    catch (const std::invalid_argument &e) { return fail(B200_ERR_ARG, e.what()); }
    catch (const std::exception &e) { return fail(B200_ERR_CUDA, e.what()); }
}


// ------------------------------------------------------------------------------------------------ assembly (host side)
static double now_ms()
{
    struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

// per-string overlap records of an index -> host pools
struct NodeHost {
    std::vector<UtgNode> node; std::vector<u8> seq; std::vector<UtgNei> nei; std::vector<UtgMark> mark;
};

static void utg_nodes_on_device(FmlEngine &E, const FmdDevice &F, int min_match, int maxlen, NodeHost &H)
{
    cudaStream_t st = E.st;
    int &nl = g_fml_stats.n_launches;
    const u64 n_str = F.idx.n_str;
    H.node.assign(n_str, UtgNode());
    if (n_str == 0) return;
    DevBuf d_node, d_seq, d_nei, d_mark, d_todo;
    d_node.reserve(n_str * sizeof(UtgNode));
    u64 seq_cap = 2 * F.idx.n + 1024, nei_cap = 4 * n_str + 4096, mark_cap = 4 * n_str + 4096;
    E.d_ctr.reserve(64);
    unsigned long long *ctr = E.d_ctr.as<unsigned long long>();
    for (int attempt = 0;; ++attempt) {
        if (attempt > 6) throw std::length_error("unitig record pools keep overflowing");
        d_seq.reserve(seq_cap); d_nei.reserve(nei_cap * sizeof(UtgNei)); d_mark.reserve(mark_cap * sizeof(UtgMark));
        UtgArgs A;
        A.e = F.idx; A.min_match = min_match; A.node = d_node.as<UtgNode>();
        A.seq = d_seq.as<u8>(); A.nei = d_nei.as<UtgNei>(); A.mark = d_mark.as<UtgMark>();
        A.seq_cap = seq_cap; A.nei_cap = nei_cap; A.mark_cap = mark_cap; A.ctr = ctr;
        A.todo = nullptr; A.n_todo = 0;
        A.cap = 2 * maxlen + 64; A.s_cap = 2 * maxlen + 32; A.tmark_cap = 64;
        const int threads = 128;
        // few strings (window-sized assemblies): 8 lanes per string cut the latency of a string ~4x; many strings: one thread
        // per string keeps 10x more of them in flight
        static const int utg_group = getenv("B200_UTG_GROUP") ? atoi(getenv("B200_UTG_GROUP")) : 0;     // 0: by size
        const bool group = utg_group ? utg_group == 8 : n_str * 8 <= (u64)E.sm_count * 4 * threads * 2;
        A.stride = (utg_scratch_bytes(A.cap, A.s_cap, A.tmark_cap, group) + 15) & ~(size_t)15;
        const int gsz = group ? 8 : 1;
        int blocks = (int)std::min<u64>((n_str * gsz + threads - 1) / threads, (u64)E.sm_count * 8);
        E.d_scratch.reserve((size_t)blocks * (threads / gsz) * A.stride);          // one slot per thread / per group
        A.scratch = E.d_scratch.as<u8>();
        CU_CHECK(cudaMemsetAsync(ctr, 0, 64, st));
        if (group) k_utg_nodes_group<8><<<blocks, threads, 0, st>>>(A);
        else k_utg_nodes<5><<<blocks, threads, 0, st>>>(A);
        ++nl;
        unsigned long long c[8];
        CU_CHECK(cudaMemcpyAsync(c, ctr, 64, cudaMemcpyDeviceToHost, st));
        CU_CHECK(cudaMemcpyAsync(H.node.data(), d_node.p, n_str * sizeof(UtgNode), cudaMemcpyDeviceToHost, st));
        CU_CHECK(cudaStreamSynchronize(st));
        CU_CHECK(cudaGetLastError());
        // strings whose scratch overflowed: same kernel, few threads, large scratch
        std::vector<u32> todo;
        for (u64 x = 0; x < n_str; ++x) if (H.node[x].flags & UTG_OVERFLOW) todo.push_back((u32)x);
        for (int round = 0; !todo.empty(); ++round) {
            if (round > 3) throw std::length_error("unitig record scratch keeps overflowing");
            UtgArgs B = A;
            B.cap = (16 * maxlen + 4096) << (2 * round); B.s_cap = 2 * maxlen + 32; B.tmark_cap = 4096 << (2 * round);
            B.stride = (utg_scratch_bytes(B.cap, B.s_cap, B.tmark_cap) + 15) & ~(size_t)15;
            int sb = (int)std::min<size_t>((todo.size() + 31) / 32, 32);
            E.d_scratch.reserve((size_t)sb * 32 * B.stride);
            B.scratch = E.d_scratch.as<u8>();
            d_todo.reserve(todo.size() * 4);
            CU_CHECK(cudaMemcpyAsync(d_todo.p, todo.data(), todo.size() * 4, cudaMemcpyHostToDevice, st));
            CU_CHECK(cudaMemsetAsync(ctr, 0, 8, st));
            B.todo = d_todo.as<u32>(); B.n_todo = todo.size();
            k_utg_nodes<5><<<sb, 32, 0, st>>>(B); ++nl;
            CU_CHECK(cudaMemcpyAsync(c, ctr, 64, cudaMemcpyDeviceToHost, st));
            CU_CHECK(cudaMemcpyAsync(H.node.data(), d_node.p, n_str * sizeof(UtgNode), cudaMemcpyDeviceToHost, st));
            CU_CHECK(cudaStreamSynchronize(st));
            CU_CHECK(cudaGetLastError());
            std::vector<u32> next;
            for (u32 x : todo) if (H.node[x].flags & UTG_OVERFLOW) next.push_back(x);
            todo.swap(next);
            g_fml_stats.n_spill += B.n_todo;
        }
        if (c[4]) {          // a pool was too small: grow to what was asked for and run again
            seq_cap = std::max<u64>(seq_cap, c[1] + 1024); nei_cap = std::max<u64>(nei_cap, 2 * c[2] + 1024); mark_cap = std::max<u64>(mark_cap, 2 * c[3] + 1024);
            continue;
        }
        H.seq.resize(c[1]); H.nei.resize(c[2]); H.mark.resize(c[3]);
        if (c[1]) CU_CHECK(cudaMemcpyAsync(H.seq.data(), d_seq.p, c[1], cudaMemcpyDeviceToHost, st));
        if (c[2]) CU_CHECK(cudaMemcpyAsync(H.nei.data(), d_nei.p, c[2] * sizeof(UtgNei), cudaMemcpyDeviceToHost, st));
        if (c[3]) CU_CHECK(cudaMemcpyAsync(H.mark.data(), d_mark.p, c[3] * sizeof(UtgMark), cudaMemcpyDeviceToHost, st));
        CU_CHECK(cudaStreamSynchronize(st));
        break;
    }
}

// fml_seq2fmi + fml_fmi2mag over reads resident on the device; g receives the graph as fml_fmi2mag returns it
static bool mag_from_device_reads(FmlEngine &E, ReadPool R, const i32 *d_len, int maxlen, const b200_fml_opt_t &opt, FmdDevice &F, Mag &g)
{
    CU_CHECK(cudaStreamSynchronize(E.st));
    CU_CHECK(cudaEventRecord(E.ev[4], 0));
    fmd_build_device(F, R.seq, R.off, d_len, R.n, E.st);
    CU_CHECK(cudaEventRecord(E.ev[5], 0));
    CU_CHECK(cudaEventSynchronize(E.ev[5]));
    g_fml_stats.ms_fmd = ms_between(E.ev[4], E.ev[5]);
    g_fml_stats.n_launches += g_fmd_launches;
    g_fml_stats.fmd_symbols = F.idx.n; g_fml_stats.n_strings = F.idx.n_str;
    g.v.clear();
    if (F.idx.n == 0) return false;          // fml_seq2fmi returned NULL
    NodeHost H;
    CU_CHECK(cudaEventRecord(E.ev[4], E.st));
    utg_nodes_on_device(E, F, opt.min_asm_ovlp, maxlen, H);
    CU_CHECK(cudaEventRecord(E.ev[5], E.st));
    CU_CHECK(cudaEventSynchronize(E.ev[5]));
    g_fml_stats.ms_nodes = ms_between(E.ev[4], E.ev[5]);
    double t0 = now_ms();
    UtgPools P; P.node = H.node.data(); P.n_str = F.idx.n_str; P.seq = H.seq.data(); P.nei = H.nei.data(); P.mark = H.mark.data();
    utg_walk_all(P, opt.min_asm_ovlp, opt.min_merge_len, g);
    g_fml_stats.ms_walk_host = (float)(now_ms() - t0);
    g_fml_stats.n_vertices = g.v.size();
    return true;
}

static b200_utgs *utgs_from_mag(const Mag &g)
{
    b200_utgs *U = new b200_utgs();
    U->utg = g.to_utg();
    U->view.resize(U->utg.size()); U->ovlp.resize(U->utg.size());
    for (size_t i = 0; i < U->utg.size(); ++i) {
        MagUtg &m = U->utg[i];
        b200_utg_t &v = U->view[i];
        v.len = (int32_t)m.seq.size(); v.nsr = m.nsr;
        v.seq = &m.seq[0]; v.cov = &m.cov[0];          // std::string keeps the terminating NUL
        v.n_ovlp[0] = m.n_ovlp[0]; v.n_ovlp[1] = m.n_ovlp[1];
        for (const MagUtgOvlp &o : m.ovlp) { b200_utg_ovlp_t t; t.len = o.len; t.from = (uint32_t)o.from; t.id = o.id; t.to = (uint32_t)o.to; U->ovlp[i].push_back(t); }
        v.ovlp = U->ovlp[i].empty() ? nullptr : U->ovlp[i].data();
    }
    return U;
}

static int host_maxlen(i64 n, const i64 *off)
{
    i64 m = 1;
    for (i64 i = 0; i < n; ++i) m = std::max(m, off[i + 1] - off[i]);
    if (m >= (1 << 24)) throw std::length_error("read longer than 2^24 bases");
    return (int)m;
}

} // namespace b200

extern "C" {

void b200_fml_opt_init(b200_fml_opt_t *opt)
{
    memset(opt, 0, sizeof(*opt));
    opt->n_threads = 1; opt->ec_k = 0; opt->min_cnt = 4; opt->max_cnt = 8; opt->min_asm_ovlp = 33; opt->min_merge_len = 0;
    b200_magopt_t &o = opt->mag_opt;
    o.trim_len = 0; o.trim_depth = 6; o.min_elen = 300; o.min_ovlp = 0; o.min_merge_len = 0; o.min_ensr = 4; o.min_insr = 3;
    o.min_dratio1 = 0.7f; o.max_bcov = 10.f; o.max_bfrac = 0.15f; o.max_bvtx = 64; o.max_bdist = 512; o.max_bdiff = 50;
    o.flag = 0x80 | 0x40;      // MAG_F_NO_SIMPL | MAG_F_POPOPEN
}

void b200_fml_opt_adjust_lens(b200_fml_opt_t *opt, int64_t n_seqs, int64_t tot_len_)
{
    uint64_t tot_len = (uint64_t)tot_len_;
    int log_len;
    if (opt->n_threads < 1) opt->n_threads = 1;
    for (log_len = 10; log_len < 32; ++log_len)
        if (1ULL << log_len > tot_len) break;
    if (opt->ec_k == 0) opt->ec_k = (log_len + 12) / 2;
    if (opt->ec_k % 2 == 0) ++opt->ec_k;
    opt->mag_opt.min_elen = (int)((double)tot_len / n_seqs * 2.5 + .499);
}

void b200_fml_opt_adjust(b200_fml_opt_t *opt, int n_seqs, const b200_fseq1_t *seqs)
{
    uint64_t tot_len = 0;
    for (int i = 0; i < n_seqs; ++i) tot_len += seqs[i].l_seq;
    b200_fml_opt_adjust_lens(opt, n_seqs, (int64_t)tot_len);
}

int b200_fml_last_stats(b200_fml_stats_t *out)
{
    if (!out) return fail(B200_ERR_ARG, "null stats");
    *out = g_fml_stats;
    return B200_OK;
}

int b200_fml_count(int64_t n, const char *seqs, const char *quals, const int64_t *off, int k, int q, int l_pre, b200_kmer_table_t **out)
{
    if (!out || n < 0 || (n > 0 && (!seqs || !off))) return fail(B200_ERR_ARG, "b200_fml_count: bad arguments");
    *out = nullptr;
    return guarded([&]() {
        FmlEngine &E = fml_engine();
        memset(&g_fml_stats, 0, sizeof(g_fml_stats));
        int64_t zero[1] = {0};
        ReadPool R = upload_reads(E, n, seqs, quals, n > 0 ? off : zero);
        b200_kmer_table *tab = new b200_kmer_table();
        try {
            CU_CHECK(cudaEventRecord(E.ev[0], E.st));
            count_on_device(E, R, n > 0 ? off[n] : 0, k, q, l_pre, tab);
            CU_CHECK(cudaEventRecord(E.ev[1], E.st));
            CU_CHECK(cudaStreamSynchronize(E.st));
            g_fml_stats.ms_count = g_fml_stats.ms_total = ms_between(E.ev[0], E.ev[1]);
        } catch (...) { delete tab; throw; }
        *out = tab;
        return (int)B200_OK;
    });
}

int b200_kmer_table_hist(const b200_kmer_table_t *tab, uint64_t cnt[256], uint64_t high[64], int *mode)
{
    if (!tab) return fail(B200_ERR_ARG, "null table");
    if (cnt) memcpy(cnt, tab->hist, 256 * 8);
    if (high) memcpy(high, tab->hist_high, 64 * 8);
    if (mode) *mode = fml_hist_mode(tab->hist);
    return B200_OK;
}

int64_t b200_kmer_table_size(const b200_kmer_table_t *tab) { return tab ? (int64_t)tab->n_distinct : 0; }

int b200_kmer_table_lookup(const b200_kmer_table_t *tab, int64_t n, const char *kmers, int32_t *occ)
{
    if (!tab || n < 0 || (n > 0 && (!kmers || !occ))) return fail(B200_ERR_ARG, "b200_kmer_table_lookup: bad arguments");
    if (n == 0) return B200_OK;
    return guarded([&]() {
        FmlEngine &E = fml_engine();
        E.d_seq.reserve((size_t)n * tab->k); E.d_len.reserve((size_t)n * 4);
        CU_CHECK(cudaMemcpyAsync(E.d_seq.p, kmers, (size_t)n * tab->k, cudaMemcpyHostToDevice, E.st));
        k_lookup<<<(unsigned)((n + 255) / 256), 256, 0, E.st>>>(tab->view(), n, E.d_seq.as<char>(), E.d_len.as<i32>());
        CU_CHECK(cudaMemcpyAsync(occ, E.d_len.p, (size_t)n * 4, cudaMemcpyDeviceToHost, E.st));
        CU_CHECK(cudaStreamSynchronize(E.st));
        CU_CHECK(cudaGetLastError());
        return (int)B200_OK;
    });
}

void b200_kmer_table_destroy(b200_kmer_table_t *tab) { delete tab; }

int b200_kmer_correct_flat(const b200_kmer_table_t *tab, int min_cov, int mode, int flt_uniq, int64_t n,
                           char *seqs, char *quals, const int64_t *off, int32_t *len_out)
{
    if (!tab || n < 0 || (n > 0 && (!seqs || !off)) || (flt_uniq && !len_out)) return fail(B200_ERR_ARG, "b200_kmer_correct_flat: bad arguments");
    if (n == 0) return B200_OK;
    return guarded([&]() {
        FmlEngine &E = fml_engine();
        memset(&g_fml_stats, 0, sizeof(g_fml_stats));
        ReadPool R = upload_reads(E, n, seqs, quals, off);
        BfcOpt bo; bfc_opt_defaults(bo);
        bo.k = tab->k; bo.l_pre = tab->l_pre; bo.q = tab->q; bo.min_cov = min_cov;
        CU_CHECK(cudaEventRecord(E.ev[0], E.st));
        correct_on_device(E, R, tab, bo, mode, flt_uniq, seqs, quals, off, len_out);
        CU_CHECK(cudaEventRecord(E.ev[1], E.st));
        CU_CHECK(cudaStreamSynchronize(E.st));
        (flt_uniq ? g_fml_stats.ms_flt : g_fml_stats.ms_ec) = g_fml_stats.ms_total = ms_between(E.ev[0], E.ev[1]);
        return (int)B200_OK;
    });
}

int b200_fml_correct_flat(const b200_fml_opt_t *opt, int flt_uniq, int64_t n, char *seqs, char *quals,
                          const int64_t *off, int32_t *len_out, float *kcov_out)
{
    if (!opt || n < 0 || (n > 0 && (!seqs || !off)) || (flt_uniq && !len_out)) return fail(B200_ERR_ARG, "b200_fml_correct_flat: bad arguments");
    // fml_correct_core (fermi-lite/bfc.c:513-553)
    BfcOpt bo; bfc_opt_defaults(bo);
    bo.k = flt_uniq ? opt->min_asm_ovlp : opt->ec_k;
    if (bo.k <= 0) {         // the reference's k = 0 run changes nothing and reports 255
        if (kcov_out) *kcov_out = 255.0f;
        if (len_out) for (int64_t i = 0; i < n; ++i) len_out[i] = (int32_t)(off[i + 1] - off[i]);
        return B200_OK;
    }
    if (bo.k > 63) return fail(B200_ERR_LIMIT, "k-mer length above 63 (BFC_MAX_KMER)");
    uint64_t tot_len = n > 0 ? (uint64_t)off[n] : 0;
    bo.l_pre = fml_initial_l_pre(tot_len);
    return guarded([&]() {
        FmlEngine &E = fml_engine();
        memset(&g_fml_stats, 0, sizeof(g_fml_stats));
        int64_t zero[1] = {0};
        CU_CHECK(cudaEventRecord(E.ev[0], E.st));
        ReadPool R = upload_reads(E, n, seqs, quals, n > 0 ? off : zero);
        b200_kmer_table tab;
        CU_CHECK(cudaEventRecord(E.ev[1], E.st));
        count_on_device(E, R, (i64)tot_len, bo.k, bo.q, bo.l_pre, &tab);
        CU_CHECK(cudaEventRecord(E.ev[2], E.st));
        bo.l_pre = tab.l_pre;
        int mode = fml_hist_mode(tab.hist);
        float kcov; int min_cov;
        fml_kcov_min_cov(tab.hist, opt->min_cnt, opt->max_cnt, kcov, min_cov);
        bo.min_cov = min_cov;
        correct_on_device(E, R, &tab, bo, mode, flt_uniq, seqs, quals, off, len_out);
        CU_CHECK(cudaEventRecord(E.ev[3], E.st));
        CU_CHECK(cudaStreamSynchronize(E.st));
        g_fml_stats.ms_count = ms_between(E.ev[1], E.ev[2]);
        (flt_uniq ? g_fml_stats.ms_flt : g_fml_stats.ms_ec) = ms_between(E.ev[2], E.ev[3]);
        g_fml_stats.ms_total = ms_between(E.ev[0], E.ev[3]);
        if (kcov_out) *kcov_out = kcov;
        return (int)B200_OK;
    });
}

static int fseq_call(const b200_fml_opt_t *opt, int flt_uniq, int n, b200_fseq1_t *s, float *kcov)
{
    if (!opt || n < 0 || (n > 0 && !s)) return fail(B200_ERR_ARG, "bad arguments");
    std::vector<int64_t> off((size_t)n + 1, 0);
    bool has_qual = false;
    for (int i = 0; i < n; ++i) {
        off[i + 1] = off[i] + (s[i].l_seq > 0 ? s[i].l_seq : 0);
        if (s[i].l_seq > 0 && s[i].qual) has_qual = true;
    }
    // the reference decides per read whether qualities exist (qual == NULL: every base counts as high quality and no
    // quality is written back): such reads ride along with a filler above any threshold
    std::vector<char> seqs((size_t)off[n] + 1), quals(has_qual ? (size_t)off[n] + 1 : 0, '~');
    for (int i = 0; i < n; ++i) if (s[i].l_seq > 0) {
        memcpy(seqs.data() + off[i], s[i].seq, s[i].l_seq);
        if (s[i].qual) memcpy(quals.data() + off[i], s[i].qual, s[i].l_seq);
    }
    std::vector<int32_t> len((size_t)n + 1);
    int rc = b200_fml_correct_flat(opt, flt_uniq, n, seqs.data(), has_qual ? quals.data() : nullptr, off.data(), len.data(), kcov);
    if (rc != B200_OK) return rc;
    for (int i = 0; i < n; ++i) {
        if (s[i].l_seq <= 0) {
            if (flt_uniq) { free(s[i].seq); free(s[i].qual); s[i].l_seq = 0; s[i].seq = s[i].qual = nullptr; }   // (0 + k - 1) / 0 is never > .8
            continue;
        }
        if (flt_uniq && len[i] == 0) { free(s[i].seq); free(s[i].qual); s[i].l_seq = 0; s[i].seq = s[i].qual = nullptr; continue; }
        memcpy(s[i].seq, seqs.data() + off[i], len[i]); s[i].seq[len[i]] = 0;
        if (s[i].qual) { memcpy(s[i].qual, quals.data() + off[i], len[i]); s[i].qual[len[i]] = 0; }
        s[i].l_seq = len[i];
    }
    return B200_OK;
}

int b200_fml_correct(const b200_fml_opt_t *opt, int n, b200_fseq1_t *seqs, float *kcov) { return fseq_call(opt, 0, n, seqs, kcov); }
int b200_fml_fltuniq(const b200_fml_opt_t *opt, int n, b200_fseq1_t *seqs, float *kcov) { return fseq_call(opt, 1, n, seqs, kcov); }


// ------------------------------------------------------------------------------------------------ assembly entry points
int b200_fmd_build(int64_t n, const char *seqs, const int64_t *off, b200_fmd_t **out)
{
    if (!out || n < 0 || (n > 0 && (!seqs || !off))) return fail(B200_ERR_ARG, "b200_fmd_build: bad arguments");
    *out = nullptr;
    return guarded([&]() {
        FmlEngine &E = fml_engine();
        int64_t zero[1] = {0};
        ReadPool R = upload_reads(E, n, seqs, nullptr, n > 0 ? off : zero);
        host_maxlen(n, n > 0 ? off : zero);
        CU_CHECK(cudaStreamSynchronize(E.st));
        b200_fmd *f = new b200_fmd();
        try { fmd_build_device(f->F, R.seq, R.off, nullptr, n, E.st); } catch (...) { delete f; throw; }
        *out = f;
        return (int)B200_OK;
    });
}

int64_t b200_fmd_len(const b200_fmd_t *f) { return f ? (int64_t)f->F.idx.n : 0; }

int b200_fmd_info(const b200_fmd_t *f, uint64_t cnt[7], uint64_t mcnt[7])
{
    if (!f) return fail(B200_ERR_ARG, "null index");
    for (int i = 0; i < 7; ++i) cnt[i] = f->F.idx.cnt[i];
    mcnt[0] = f->F.idx.n;
    for (int i = 1; i < 7; ++i) mcnt[i] = f->F.idx.cnt[i] - f->F.idx.cnt[i - 1];
    return B200_OK;
}

int b200_fmd_bwt(const b200_fmd_t *f, uint8_t *bwt)
{
    if (!f || !bwt) return fail(B200_ERR_ARG, "b200_fmd_bwt: bad arguments");
    if (f->F.idx.n == 0) return B200_OK;
    return guarded([&]() {
        CU_CHECK(cudaMemcpy(bwt, f->F.bwt8.p, f->F.idx.n, cudaMemcpyDeviceToHost));
        return (int)B200_OK;
    });
}

int b200_fmd_rank1a(const b200_fmd_t *f, int64_t n_q, const uint64_t *q, uint64_t *ranks, int32_t *sym)
{
    if (!f || n_q < 0 || (n_q > 0 && (!q || !ranks || !sym))) return fail(B200_ERR_ARG, "b200_fmd_rank1a: bad arguments");
    if (n_q == 0) return B200_OK;
    if (f->F.idx.n == 0) return fail(B200_ERR_ARG, "empty index");
    return guarded([&]() {
        FmlEngine &E = fml_engine();
        DevBuf dq, dr, ds;
        dq.reserve(n_q * 8); dr.reserve(n_q * 48); ds.reserve(n_q * 4);
        CU_CHECK(cudaMemcpyAsync(dq.p, q, n_q * 8, cudaMemcpyHostToDevice, E.st));
        k_fmd_rank<<<(unsigned)((n_q + 255) / 256), 256, 0, E.st>>>(f->F.idx, n_q, dq.as<u64>(), dr.as<u64>(), ds.as<i32>());
        CU_CHECK(cudaMemcpyAsync(ranks, dr.p, n_q * 48, cudaMemcpyDeviceToHost, E.st));
        CU_CHECK(cudaMemcpyAsync(sym, ds.p, n_q * 4, cudaMemcpyDeviceToHost, E.st));
        CU_CHECK(cudaStreamSynchronize(E.st));
        CU_CHECK(cudaGetLastError());
        return (int)B200_OK;
    });
}

void b200_fmd_destroy(b200_fmd_t *f) { delete f; }

int b200_fml_mag_text(const b200_fml_opt_t *opt, int stage, int64_t n, const char *seqs, const int64_t *off,
                      char **text, int64_t *text_len, float *rdist)
{
    if (!opt || !text || !text_len || n < 0 || (n > 0 && (!seqs || !off))) return fail(B200_ERR_ARG, "b200_fml_mag_text: bad arguments");
    *text = nullptr; *text_len = 0;
    return guarded([&]() {
        FmlEngine &E = fml_engine();
        memset(&g_fml_stats, 0, sizeof(g_fml_stats));
        int64_t zero[1] = {0};
        ReadPool R = upload_reads(E, n, seqs, nullptr, n > 0 ? off : zero);
        int maxlen = host_maxlen(n, n > 0 ? off : zero);
        FmdDevice F; Mag g;
        mag_from_device_reads(E, R, nullptr, maxlen, *opt, F, g);
        if (rdist) *rdist = g.rdist;
        if (stage >= 1) g.fml_clean(*opt);
        std::string t = g.text();
        char *o = (char *)malloc(t.size() + 1);
        if (!o) throw std::bad_alloc();
        memcpy(o, t.data(), t.size()); o[t.size()] = 0;
        *text = o; *text_len = (int64_t)t.size();
        return (int)B200_OK;
    });
}

int b200_fml_seqs2utg_flat(const b200_fml_opt_t *opt, int64_t n, const char *seqs, const int64_t *off, b200_utgs_t **out)
{
    if (!opt || !out || n < 0 || (n > 0 && (!seqs || !off))) return fail(B200_ERR_ARG, "b200_fml_seqs2utg_flat: bad arguments");
    *out = nullptr;
    return guarded([&]() {
        FmlEngine &E = fml_engine();
        memset(&g_fml_stats, 0, sizeof(g_fml_stats));
        int64_t zero[1] = {0};
        ReadPool R = upload_reads(E, n, seqs, nullptr, n > 0 ? off : zero);
        int maxlen = host_maxlen(n, n > 0 ? off : zero);
        FmdDevice F; Mag g;
        mag_from_device_reads(E, R, nullptr, maxlen, *opt, F, g);
        double t0 = now_ms();
        g.fml_clean(*opt);
        g_fml_stats.ms_clean_host = (float)(now_ms() - t0);
        *out = utgs_from_mag(g);
        g_fml_stats.n_utg = (*out)->utg.size();
        return (int)B200_OK;
    });
}

int b200_fml_assemble_flat(const b200_fml_opt_t *opt0, int64_t n, const char *seqs, const char *quals, const int64_t *off, b200_utgs_t **out)
{
    if (!opt0 || !out || n < 0 || (n > 0 && (!seqs || !off))) return fail(B200_ERR_ARG, "b200_fml_assemble_flat: bad arguments");
    *out = nullptr;
    if (n == 0) { *out = new b200_utgs(); return B200_OK; }
    return guarded([&]() {
        FmlEngine &E = fml_engine();
        memset(&g_fml_stats, 0, sizeof(g_fml_stats));
        b200_fml_opt_t opt = *opt0;
        uint64_t tot_len = (uint64_t)off[n];
        b200_fml_opt_adjust_lens(&opt, n, (int64_t)tot_len);
        if (opt.ec_k > 63 || opt.min_asm_ovlp > 63 || opt.min_asm_ovlp < 1) throw std::length_error("k-mer length outside [1, 63]");
        int maxlen = host_maxlen(n, off);
        CU_CHECK(cudaEventRecord(E.ev[0], E.st));
        ReadPool R = upload_reads(E, n, seqs, quals, off);
        float kcov = 0; int min_cov = 0;
        // fml_correct
        if (opt.ec_k >= 0) {
            if (opt.ec_k > 0) {
                BfcOpt bo; bfc_opt_defaults(bo);
                bo.k = opt.ec_k; bo.l_pre = fml_initial_l_pre(tot_len);
                b200_kmer_table tab;
                CU_CHECK(cudaEventRecord(E.ev[1], E.st));
                count_on_device(E, R, (i64)tot_len, bo.k, bo.q, bo.l_pre, &tab);
                CU_CHECK(cudaEventRecord(E.ev[2], E.st));
                bo.l_pre = tab.l_pre;
                fml_kcov_min_cov(tab.hist, opt.min_cnt, opt.max_cnt, kcov, min_cov);
                bo.min_cov = min_cov;
                correct_on_device(E, R, &tab, bo, fml_hist_mode(tab.hist), 0, nullptr, nullptr, off, nullptr);
                CU_CHECK(cudaEventRecord(E.ev[3], E.st));
                CU_CHECK(cudaStreamSynchronize(E.st));
                g_fml_stats.ms_count = ms_between(E.ev[1], E.ev[2]);
                g_fml_stats.ms_ec = ms_between(E.ev[2], E.ev[3]);
            }
        }
        // fml_fltuniq
        {
            BfcOpt bo; bfc_opt_defaults(bo);
            bo.k = opt.min_asm_ovlp; bo.l_pre = fml_initial_l_pre(tot_len);
            b200_kmer_table tab;
            CU_CHECK(cudaEventRecord(E.ev[1], E.st));
            count_on_device(E, R, (i64)tot_len, bo.k, bo.q, bo.l_pre, &tab);
            bo.l_pre = tab.l_pre;
            fml_kcov_min_cov(tab.hist, opt.min_cnt, opt.max_cnt, kcov, min_cov);
            bo.min_cov = min_cov;
            correct_on_device(E, R, &tab, bo, fml_hist_mode(tab.hist), 1, nullptr, nullptr, off, nullptr);
            CU_CHECK(cudaEventRecord(E.ev[2], E.st));
            CU_CHECK(cudaStreamSynchronize(E.st));
            g_fml_stats.ms_flt = ms_between(E.ev[1], E.ev[2]);
        }
        // fml_seq2fmi + fml_fmi2mag
        FmdDevice F; Mag g;
        if (!mag_from_device_reads(E, R, E.d_len.as<i32>(), maxlen, opt, F, g)) { *out = new b200_utgs(); return (int)B200_OK; }
        // min_ensr / min_insr from kcov (fermi-lite/misc.c:295-298), fml_mag_clean, fml_mag2utg
        opt.mag_opt.min_ensr = opt.mag_opt.min_ensr > kcov * .1 ? opt.mag_opt.min_ensr : (int)(kcov * .1 + .499);
        opt.mag_opt.min_ensr = opt.mag_opt.min_ensr < opt0->max_cnt ? opt.mag_opt.min_ensr : opt0->max_cnt;
        opt.mag_opt.min_ensr = opt.mag_opt.min_ensr > opt0->min_cnt ? opt.mag_opt.min_ensr : opt0->min_cnt;
        opt.mag_opt.min_insr = opt.mag_opt.min_ensr - 1;
        double t0 = now_ms();
        g.fml_clean(opt);
        g_fml_stats.ms_clean_host = (float)(now_ms() - t0);
        *out = utgs_from_mag(g);
        g_fml_stats.n_utg = (*out)->utg.size();
        CU_CHECK(cudaEventRecord(E.ev[3], E.st));
        CU_CHECK(cudaStreamSynchronize(E.st));
        g_fml_stats.ms_total = ms_between(E.ev[0], E.ev[3]);
        return (int)B200_OK;
    });
}

int b200_fml_assemble_windows(const b200_fml_opt_t *opt, int64_t n_windows, const int64_t *win_off,
                              const char *seqs, const char *quals, const int64_t *off, int n_threads, b200_utgs_t **out)
{
    if (!opt || !out || n_windows < 0 || (n_windows > 0 && (!win_off || !off || !seqs))) return fail(B200_ERR_ARG, "b200_fml_assemble_windows: bad arguments");
    for (int64_t w = 0; w < n_windows; ++w) {
        out[w] = nullptr;
        if (win_off[w + 1] < win_off[w]) return fail(B200_ERR_ARG, "b200_fml_assemble_windows: win_off must not decrease");
    }
    if (n_windows == 0) return B200_OK;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return fail(B200_ERR_CUDA, "b200_fml_assemble_windows: no CUDA device"); }
    unsigned nt = n_threads > 0 ? (unsigned)n_threads : 4;      // measured best on a B200 box (scripts/win_probe.py); see DESIGN.md section 6
    if (nt > 32) nt = 32;
    if ((int64_t)nt > n_windows) nt = (unsigned)n_windows;
    std::atomic<int64_t> next(0);
    std::mutex mu; int first_rc = B200_OK; std::string first_msg;
    b200_fml_stats_t total; memset(&total, 0, sizeof(total));
    auto work = [&]() {
        cudaSetDevice(dev);
        std::vector<int64_t> loff;
        for (;;) {
            int64_t w = next.fetch_add(1);
            if (w >= n_windows) break;
            const int64_t r0 = win_off[w], n = win_off[w + 1] - r0;
            loff.resize((size_t)n + 1);
            for (int64_t i = 0; i <= n; ++i) loff[(size_t)i] = off[r0 + i] - off[r0];
            int rc = b200_fml_assemble_flat(opt, n, seqs + off[r0], quals ? quals + off[r0] : nullptr, loff.data(), &out[w]);
            std::lock_guard<std::mutex> g(mu);
            if (rc != B200_OK && first_rc == B200_OK) { first_rc = rc; first_msg = b200_last_error(); }
            total.n_launches += g_fml_stats.n_launches; total.n_utg += g_fml_stats.n_utg; total.n_strings += g_fml_stats.n_strings;
            total.fmd_symbols += g_fml_stats.fmd_symbols; total.n_lookups += g_fml_stats.n_lookups;
            // summed over the windows: with concurrent windows the stage times add up to more than the wall time
            total.ms_count += g_fml_stats.ms_count; total.ms_ec += g_fml_stats.ms_ec; total.ms_flt += g_fml_stats.ms_flt; total.ms_fmd += g_fml_stats.ms_fmd;
            total.ms_nodes += g_fml_stats.ms_nodes; total.ms_walk_host += g_fml_stats.ms_walk_host; total.ms_clean_host += g_fml_stats.ms_clean_host;
            total.ms_total += g_fml_stats.ms_total;
        }
    };
    if (nt == 1) work();
    else {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nt; ++t) th.emplace_back(work);
        for (auto &x : th) x.join();
    }
    g_fml_stats = total;
    if (first_rc != B200_OK) return fail(first_rc, first_msg.c_str());
    return B200_OK;
}

int b200_utgs_view(const b200_utgs_t *u, int *n_utg, const b200_utg_t **utg)
{
    if (!u || !n_utg || !utg) return fail(B200_ERR_ARG, "b200_utgs_view: bad arguments");
    *n_utg = (int)u->view.size();
    *utg = u->view.empty() ? nullptr : u->view.data();
    return B200_OK;
}

void b200_utgs_free(b200_utgs_t *u) { delete u; }

int b200_fml_assemble(const b200_fml_opt_t *opt, int n_seqs, const b200_fseq1_t *s, int *n_utg, b200_utg_t **utg)
{
    if (!opt || !n_utg || !utg || n_seqs < 0 || (n_seqs > 0 && !s)) return fail(B200_ERR_ARG, "b200_fml_assemble: bad arguments");
    *n_utg = 0; *utg = nullptr;
    std::vector<int64_t> off((size_t)n_seqs + 1, 0);
    bool has_qual = false;
    for (int i = 0; i < n_seqs; ++i) {
        off[i + 1] = off[i] + (s[i].l_seq > 0 ? s[i].l_seq : 0);
        if (s[i].l_seq > 0 && s[i].qual) has_qual = true;
    }
    std::vector<char> seqs((size_t)off[n_seqs] + 1), quals(has_qual ? (size_t)off[n_seqs] + 1 : 0, '~');
    for (int i = 0; i < n_seqs; ++i) if (s[i].l_seq > 0) {
        memcpy(seqs.data() + off[i], s[i].seq, s[i].l_seq);
        if (s[i].qual) memcpy(quals.data() + off[i], s[i].qual, s[i].l_seq);
    }
    b200_utgs_t *U = nullptr;
    int rc = b200_fml_assemble_flat(opt, n_seqs, seqs.data(), has_qual ? quals.data() : nullptr, off.data(), &U);
    if (rc != B200_OK) return rc;
    int n = (int)U->view.size();
    if (n == 0) { delete U; return B200_OK; }         // the reference returns NULL with *n_utg untouched by the graph code
    b200_utg_t *a = (b200_utg_t *)calloc(n, sizeof(b200_utg_t));
    if (!a) { delete U; return fail(B200_ERR_NOMEM, "out of host memory"); }
    for (int i = 0; i < n; ++i) {
        const b200_utg_t &v = U->view[i];
        a[i] = v;
        a[i].seq = (char *)malloc(v.len + 1); memcpy(a[i].seq, v.seq, v.len); a[i].seq[v.len] = 0;
        a[i].cov = (char *)malloc(v.len + 1); memcpy(a[i].cov, v.cov, v.len); a[i].cov[v.len] = 0;
        int no = v.n_ovlp[0] + v.n_ovlp[1];
        a[i].ovlp = (b200_utg_ovlp_t *)calloc(no ? no : 1, sizeof(b200_utg_ovlp_t));
        if (no) memcpy(a[i].ovlp, v.ovlp, no * sizeof(b200_utg_ovlp_t));
    }
    delete U;
    *n_utg = n; *utg = a;
    return B200_OK;
}

void b200_fml_utg_destroy(int n_utg, b200_utg_t *utg)
{
    if (!utg) return;
    for (int i = 0; i < n_utg; ++i) { free(utg[i].seq); free(utg[i].cov); free(utg[i].ovlp); }
    free(utg);
}

} // extern "C"
