// engine.cu -- read-batched seed-and-extend on the GPU: kernels, chunk driver, C ABI.
//
//   b200_mem_align_batch  <- the compute of BWAAligner::alignSequence (src/BWAAligner.cpp:104-128):
//                            mem_align1 (bwa/bwamem_extra.c:103-115) + mem_reg2aln (bwa/bwamem.c:1119-1189) per region
//   b200_ksw_extend2_batch<- ksw_extend2 (bwa/ksw.c:416-515)
//
// Per chunk of reads four kernels run back to back on one stream, each thread
// taking reads from a shared work counter (a warp claims 32 reads at a time):
//   k_seed -> k_chain -> k_extend -> k_finalize, then k_gather compacts the hits
// of the chunk in read order (cub scan of the per-read counts).  There is no
// CPU path: without a usable CUDA device every entry point returns B200_ERR_CUDA.
#include <cub/cub.cuh>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <string>
#include <algorithm>
#include <mutex>
#include <unordered_map>
#include <time.h>
#include "engine.cuh"
#include "extend_group.cuh"
#include "finalize_group.cuh"

using namespace b200;

namespace b200 {

static double wall_now();
struct DevCounters { unsigned long long v[8]; };   // occ_blocks, sa_reads, ref_bytes, sw_cells, n_ext, n_global, tab_lo, tab_hi

__device__ __forceinline__ void flush_counters(const CtrLocal &c, DevCounters *g)
{
    // warp-aggregate, one atomic per warp and counter
    unsigned long long v[8] = {c.occ_blocks, c.sa_reads, c.ref_bytes, c.sw_cells, c.n_ext, c.n_global, c.tab_lo, c.tab_hi};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        unsigned long long x = v[k];
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if ((threadIdx.x & 31) == 0 && x) atomicAdd(&g->v[k], x);
    }
}

// a warp claims 32 consecutive work items
// How many work items a warp takes at a time: all `full` slots of the warp when there is enough work, fewer for small batches, so
// that the items spread over the warps of the grid instead of filling a few warps whose lanes then serialise on each other's
// divergent paths (a 1 024-read call is bound by the latency of ONE read's dependent chain, not by throughput).
__device__ __forceinline__ int warp_share(i64 n_work, int full)
{
    const i64 warps = (i64)gridDim.x * (blockDim.x >> 5);
    const i64 w = (n_work + warps - 1) / warps;
    return (int)(w < 1 ? 1 : w > full ? full : w);
}
// a warp claims `width` consecutive work items; lanes >= width get an index beyond n_work semantics via the caller's check
__device__ __forceinline__ i64 claim(unsigned long long *counter, int width)
{
    unsigned long long base = 0;
    if ((threadIdx.x & 31) == 0) base = atomicAdd(counter, (unsigned long long)width);
    base = __shfl_sync(0xffffffffu, base, 0);
    return (i64)base + (threadIdx.x & 31);
}

struct KArgs {
    DevIndex ix; Opt opt; Caps caps; Batch B;
    const i32 *order; i64 n_work;           // work item i processes read order[i] (NULL: i)
    u8 *scratch; size_t scratch_stride;
    unsigned long long *work_ctr; DevCounters *ctrs;
    const double *log_tab; int n_log;
    SeedTab tab;                            // prefix-chain table of the index + text-path permission (seed2.cuh); K == 0: no table
    const unsigned long long *n_work_dev;   // when set: the number of work items lives on the device (retry lists)
};

template <int STAGE>
__device__ __forceinline__ void stage_body(const KArgs &A)
{
    u8 *scr = A.scratch + (size_t)(blockIdx.x * blockDim.x + threadIdx.x) * A.scratch_stride;
    const i64 n_work = A.n_work_dev ? (i64)*A.n_work_dev : A.n_work;
    CtrLocal ctr;
    const int width = warp_share(n_work, 32);
    for (;;) {
        i64 w = claim(A.work_ctr, width);
        if (w - (threadIdx.x & 31) >= n_work) break;
        if ((int)(threadIdx.x & 31) < width && w < n_work) {
            i64 rid = A.order ? A.order[w] : w;
            if (STAGE == 0) stage_seed_t<true>(A.ix, A.opt, A.caps, A.B, rid, scr, ctr);
            else if (STAGE == 1) stage_chain(A.ix, A.opt, A.caps, A.B, rid, scr, ctr, A.log_tab, A.n_log);
            else if (STAGE == 2) stage_extend(A.ix, A.opt, A.caps, A.B, rid, scr, ctr);
            else stage_finalize(A.ix, A.opt, A.caps, A.B, rid, scr, A.log_tab, A.n_log, ctr);
        }
    }
    flush_counters(ctr, A.ctrs);
}
// (stages 1 and 3 are latency bound on their scratch slots: five blocks per SM is the measured optimum, more thrash the caches, fewer hide less)
template <int STAGE>
__global__ void __launch_bounds__(128, (STAGE == 1 || STAGE == 3) ? 5 : 4) k_stage(const __grid_constant__ KArgs A) { stage_body<STAGE>(A); }
// the same with a register cap that lets MINB blocks of 128 threads share an SM: the thread-per-read chain / finalize stages
// wait on dependent scratch accesses (issue slots 13-15 % busy), more resident warps hide more of that latency
template <int STAGE, int MINB>
__global__ void __launch_bounds__(128, MINB) k_stage_occ(const __grid_constant__ KArgs A) { stage_body<STAGE>(A); }

// ---------------------------------------------------------------------------------------
// Seeding, main pass: the single-gather-site machine of seed2.cuh.
//   shared memory per block: a ring of CAP packed long entries per thread (12 B each, see SmemList) + the read as 2-bit words (word w of
//   thread t at [w * 128 + t]); the short entries of a sweep are a bit mask in a register.
//   A lane whose read is finished records it and claims the next one by itself, so the 32 lanes stay inside the one
//   loop and reach the Occ gathers together.  Intervals go straight to the read's fixed slot of the interval pool
//   (read r owns [r * stride, (r+1) * stride)), in production order; k_sort_intv orders them afterwards.
//   Reads the machine does not take (N bases, list or slot overflow) get OVF_INTV and go through the spill pass,
//   i.e. the reference-shaped k_stage<0>.
// The ring of long work-list entries of k_seed2 in shared memory: 12 bytes per entry (word k of entry e of thread t at
// [(3 e + k) * 128 + t], conflict free): x0, x1 low words; then x0 / x1 high bits | size | end (8 bits).  The 24 bits in front of the
// end hold `hb` high bits of each coordinate and 24 - 2 hb bits of size: hb = 1 for texts below 2^33 symbols (a human-sized
// reference: sizes up to 4.19 M, i.e. every repeat family keeps its entries in the ring), hb = 4 up to 2^36 (sizes up to 65 534).
// An entry whose size does not fit is not kept: take() says so and the machine sends the read to the reference-shaped kernel.
template <int hb>
struct SmemList {
    u32 *p;
    __device__ __forceinline__ void put(int e, u64 x0, u64 x1, u64 x2, u32 end)
    {
        u32 *q = p + e * 384;
        const u32 smax = (1u << (24 - 2 * hb)) - 1u;
        q[0] = (u32)x0; q[128] = (u32)x1;
        q[256] = (u32)(x0 >> 32) | (u32)(x1 >> 32) << hb | (x2 < (u64)smax ? (u32)x2 : smax) << (2 * hb) | end << 24;
    }
    __device__ __forceinline__ bool take(int e, u64 &x0, u64 &x1, u64 &x2, u32 &end) const
    {
        const u32 *q = p + e * 384;
        const u32 w = q[256], hm = (1u << hb) - 1u, smax = (1u << (24 - 2 * hb)) - 1u;
        x0 = (u64)q[0] | (u64)(w & hm) << 32; x1 = (u64)q[128] | (u64)((w >> hb) & hm) << 32;
        x2 = (w >> (2 * hb)) & smax; end = w >> 24;
        return x2 != (u64)smax;
    }
    __device__ __forceinline__ u32 end(int e) const { return p[e * 384 + 256] >> 24; }
};
struct SmemQuery {
    const u32 *p;
    __device__ __forceinline__ int operator[](int i) const { return (int)((p[(i >> 4) * 128] >> ((i & 15) * 2)) & 3u); }
    // bases [st, st + ln), ln <= 16, base t in bits 2t.. (the word after the read's last one is zero padding)
    __device__ __forceinline__ u32 key(int st, int ln) const
    {
        const u32 lo = p[(st >> 4) * 128], hi = p[((st >> 4) + 1) * 128];
        const u32 w = __funnelshift_r(lo, hi, (st & 15) * 2);
        return ln >= 16 ? w : w & ((1u << (2 * ln)) - 1u);
    }
};

// 2-bit words of every read (16 bases per word, base i in bits 2(i&15)..) + a flag for reads with a base > 3
__global__ void k_pack_reads(const u8 *__restrict__ seq, const i64 *__restrict__ off, i64 n, int qw, u32 *__restrict__ packed, u32 *__restrict__ bad)
{
    i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * qw) return;
    i64 r = t / qw; int w = (int)(t - r * qw);
    i64 b = off[r], len = off[r + 1] - b;
    u32 v = 0; bool n_base = false;
    for (int k = 0; k < 16; ++k) {
        i64 i = (i64)w * 16 + k;
        if (i >= len) break;
        u32 c = seq[b + i];
        if (c > 3) { n_base = true; c = 0; }
        v |= c << (2 * k);
    }
    packed[t] = v;
    if (n_base || (w == 0 && len > (i64)qw * 16)) atomicOr(bad + r, 1u);
}

template <int CAP, int MINB, int PHASED, int HB>
__global__ void __launch_bounds__(128, MINB) k_seed2(const __grid_constant__ KArgs A, const u32 *__restrict__ packed, int qw, const u32 *__restrict__ bad, int stride)
{
    // The SMEM passes (bwt_smem1a from every start, re-seeding of long SMEMs).  Tried and dropped (round 2): L2 eviction-policy
    // hints / a persisting window for the low table levels (no change in hit rate or time, profiles/r02_seed_l2hint_ab.txt);
    // entry ends in shared memory + intervals in an L2-resident global slot (6 blocks per SM but a global access per step: no gain).
    extern __shared__ u32 seed_smem[];
    SmemList<HB> L; L.p = seed_smem + threadIdx.x;
    u32 *myq = seed_smem + CAP * 384 + threadIdx.x;
    SmemQuery Q; Q.p = myq;
    myq[qw * 128] = 0;
    SeedMachine<SmemList<HB>, SmemQuery> m;
    m.mode = 0; m.ovf = 0;
    CtrLocal ctr;
    i64 rid = -1;
    bool done = (int)(threadIdx.x & 31) >= warp_share(A.n_work, 32);
    int phase = 0;
    for (;;) {
        if (m.mode == 0 && !done) {
            if (rid >= 0) {                                     // record the read that just finished
                ReadRec &R = A.B.rec[rid];
                if (m.ovf) { A.B.ovf[rid] |= OVF_INTV; R.n_intv = 0; R.intv_off = 0; }
                else { R.n_intv = m.out.n; R.intv_off = rid * stride; }
                rid = -1;
            }
            i64 w = (i64)atomicAdd(A.work_ctr, 1ull);
            if (w >= A.n_work) done = true;
            else {
                rid = A.order ? A.order[w] : w;
                ReadRec &R = A.B.rec[rid];
                int len = (int)(A.B.seq_off[rid + 1] - A.B.seq_off[rid]);
                if (A.B.ovf[rid] || bad[rid]) {
                    if (!A.B.ovf[rid]) A.B.ovf[rid] = OVF_INTV;
                    R.n_intv = 0; R.intv_off = 0; rid = -1;
                } else {
                    const u32 *src = packed + rid * qw;
                    for (int k = 0; k < qw; ++k) myq[k * 128] = src[k];
                    IntvSink out; out.a = A.B.pool.intv + rid * stride; out.n = 0; out.cap = stride; out.overflow = false;
                    m.init(A.opt, len, CAP, L, Q, out, A.tab.K, 2, A.tab.text);
                    m.start(A.ix);
                }
            }
        }
        if (__all_sync(0xffffffffu, done)) break;
        bool go = m.mode != 0;
        if (PHASED) {
            // the lanes of a warp take turns by class: forward sweeps (chain entry, a few Occ steps, text path) or backward sweeps
            // (rows).  A lane of the other class waits, so the lanes that run are in the same few code paths at the same time
            // (107 -> 95 ms per 10 M reads).  Finer turns -- always the kind of step most lanes wait for, six kinds -- are slower
            // (112 ms): every turn costs a gather latency, and fewer lanes share it.
            const int cls = m.mode == 0 ? -1 : (m.mode == m.M_BWD || m.mode == m.M_BTX) ? 1 : 0;
            if (__ballot_sync(0xffffffffu, cls == phase) == 0) phase ^= 1;
            go = cls == phase;
        }
        if (go) {
            u64 a, o, s, na = 0, no = 0, ns = 0; int c; u32 key; const void *ga, *gb; ChainView cv;
            const int kind = m.request(A.ix, a, o, s, c, key, ga, gb);
            if (kind >= 0) {
                gather(A.ix, A.tab, kind, key, ga, gb, a, o, s, c, na, no, ns, cv, ctr);
                if (kind == 0) m.consume(A.ix, na, no, ns); else if (kind == 1) m.consume_chain(A.ix, cv); else m.consume_aux(A.ix);
            }
        }
    }
    flush_counters(ctr, A.ctrs);
}

// The third pass (bwt_seed_strategy1 from every start, bwa/bwamem.c:173-184) as its own kernel: every lane is a forward walk, no
// work list, 48 bytes of shared memory per thread -- occupancy and convergence that the mixed kernel cannot have.
struct NoList {
    __device__ __forceinline__ void put(int, u64, u64, u64, u32) {}
    __device__ __forceinline__ bool take(int, u64 &, u64 &, u64 &, u32 &) const { return false; }
    __device__ __forceinline__ u32 end(int) const { return 0; }
};
__global__ void __launch_bounds__(128, 8) k_seed3(const __grid_constant__ KArgs A, const u32 *__restrict__ packed, int qw, int stride)
{
    extern __shared__ u32 seed_smem[];
    u32 *myq = seed_smem + threadIdx.x;
    SmemQuery Q; Q.p = myq;
    myq[qw * 128] = 0;
    SeedMachine<NoList, SmemQuery> m;
    m.mode = 0; m.ovf = 0;
    CtrLocal ctr;
    i64 rid = -1;
    bool done = (int)(threadIdx.x & 31) >= warp_share(A.n_work, 32);
    for (;;) {
        if (m.mode == 0 && !done) {
            if (rid >= 0) {
                ReadRec &R = A.B.rec[rid];
                if (m.ovf) { A.B.ovf[rid] |= OVF_INTV; R.n_intv = 0; R.intv_off = 0; }
                else R.n_intv = m.out.n;
                rid = -1;
            }
            i64 w = (i64)atomicAdd(A.work_ctr, 1ull);
            if (w >= A.n_work) done = true;
            else {
                rid = A.order ? A.order[w] : w;
                if (A.B.ovf[rid]) rid = -1;                      // not taken by k_seed2: the reference-shaped kernel does all three passes
                else {
                    const int len = (int)(A.B.seq_off[rid + 1] - A.B.seq_off[rid]);
                    const u32 *src = packed + rid * qw;
                    for (int k = 0; k < qw; ++k) myq[k * 128] = src[k];
                    IntvSink out; out.a = A.B.pool.intv + rid * stride; out.n = 0; out.cap = stride; out.overflow = false;
                    NoList L;
                    m.init(A.opt, len, 0, L, Q, out, A.tab.K, 3, 0);
                    m.start3(A.ix, A.B.rec[rid].n_intv);
                }
            }
        }
        if (__all_sync(0xffffffffu, done)) break;
        if (m.mode != 0) {
            u64 a, o, s, na = 0, no = 0, ns = 0; int c; u32 key; const void *ga, *gb; ChainView cv;
            const int kind = m.request(A.ix, a, o, s, c, key, ga, gb);
            if (kind >= 0) {
                gather(A.ix, A.tab, kind, key, ga, gb, a, o, s, c, na, no, ns, cv, ctr);
                if (kind == 0) m.consume(A.ix, na, no, ns); else if (kind == 1) m.consume_chain(A.ix, cv); else m.consume_aux(A.ix);
            }
        }
    }
    flush_counters(ctr, A.ctrs);
}

// Builder scratch: level j holds the packed interval of every string of j bases (key = the string, base i in bits 2i..);
// level j from level j-1: entry[key] = forward extension of entry[key's first j-1 bases] by its last base, computed by the
// very code the seeding kernel runs (so an entry is what the iterated bwt_extend would have produced, bwa/bwt.c:262-275,
// including the coordinates of empty intervals).
__global__ void k_seedtab_level(const DevIndex ix, PIntv *__restrict__ base, int j)
{
    const u64 key = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (key >= (1ull << (2 * j))) return;
    PIntv *out = base + seedtab_level_off(j) + key;
    if (j == 1) {
        Intv t; set_intv(ix, (int)key, t);
        *out = pintv_pack(t.x0, t.x1, t.x2, 0);
        return;
    }
    const u64 parent = key & ((1ull << (2 * (j - 1))) - 1);
    const int c = (int)(key >> (2 * (j - 1)));
    u64 x0, x1, x2, na, no, ns; u32 e;
    pintv_unpack(base[seedtab_level_off(j - 1) + parent], x0, x1, x2, e);
    CtrLocal ctr;
    extend_lean(ix, x1, x0, x2, 3 - c, na, no, ns, ctr);
    *out = pintv_pack(no, na, ns, 0);
}
// The chain table (seed2.cuh) from levels 1..K-1: the K-mer's own interval is one more extension, the prefix sizes are read
// from the levels (consecutive keys share their prefixes: coalesced, cache resident for the low levels).
__global__ void k_chain_build(const DevIndex ix, const PIntv *__restrict__ lev, ChainEnt *__restrict__ out, int K)
{
    const u64 key = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (key >= (1ull << (2 * K))) return;
    u64 sz[18];
    u64 x0 = 0, x1 = 0, x2 = 0; u32 e;
    for (int m = 1; m < K; ++m) {
        pintv_unpack(lev[seedtab_level_off(m) + (key & ((1ull << (2 * m)) - 1))], x0, x1, x2, e);
        sz[m] = x2;
    }
    const int c = (int)(key >> (2 * (K - 1)));
    u64 na, no, ns;
    CtrLocal ctr;
    extend_lean(ix, x1, x0, x2, 3 - c, na, no, ns, ctr);       // x0..x2 = the (K-1)-prefix
    sz[K] = ns;
    out[key] = chain_make(K, no, na, ns, sz);
}

// The text path of the seeding machine (seed2.cuh) compares reads with ix.text: that is only the FM-index's answer when the text
// IS the text of the BWT.  BWT[k] == text[SA[k] - 1] for every rank k proves it (SA is a permutation of 0..seq_len); an index whose
// BWT was built over another randomisation of the N bases (SeqLib's ConstructIndex, src/BWAIndex.cpp:102-125) fails and seeds through the
// Occ blocks only.
__global__ void k_verify_text(const DevIndex ix, unsigned long long *__restrict__ bad)
{
    const u64 k = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (k > ix.seq_len || k == ix.primary) return;
    const u64 p = k ? ix.sa[k] : ix.seq_len;                               // rank 0 is the empty suffix (bwa keeps -1 there)
    if (p == 0 || p > ix.seq_len) { atomicAdd(bad, 1ull); return; }       // rank k != primary has a preceding text symbol
    const u64 x = k - (k > ix.primary);
    const OccBlock &b = ix.occ[x >> 6];
    if (occ_sym(b, (int)(x & 63)) != text_base(ix, (i64)(p - 1))) atomicAdd(bad, 1ull);
}

// order each read's intervals by (start, end) -- the ks_introsort of mem_collect_intv (bwa/bwamem.c:186); equal keys are identical intervals
__global__ void k_sort_intv(const ReadRec *__restrict__ rec, const u32 *__restrict__ ovf, i64 n, Intv *__restrict__ pool)
{
    i64 r = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n || ovf[r]) return;
    sort_intv_by_info(pool + rec[r].intv_off, rec[r].n_intv);
}
__global__ void k_set_u64(unsigned long long *p, unsigned long long v) { *p = v; }

// stage 2 with G lanes per read (see extend_group.cuh); a warp claims 32/G reads at a time
template <int G, bool REG>
__global__ void __launch_bounds__(128, REG ? 4 : 3) k_extend_group(const __grid_constant__ KArgs A)
{
    extern __shared__ __align__(16) u8 smem_raw[];
    __shared__ i8 smat[32];
    if (threadIdx.x < 25) smat[threadIdx.x] = A.opt.mat[threadIdx.x];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int gib = threadIdx.x / G;
    GroupCtx<G> g;
    g.gl = threadIdx.x % G;
    g.mask = G == 32 ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(G - 1)));
    u8 *smem = REG ? smem_raw : smem_raw + (size_t)gib * group_smem_bytes(A.caps.maxlen);
    u8 *scr = A.scratch + (size_t)(blockIdx.x * (128 / G) + gib) * A.scratch_stride;
    const i64 n_work = A.n_work_dev ? (i64)*A.n_work_dev : A.n_work;
    const int share = warp_share(n_work, 32 / G);
    CtrLocal ctr;
    for (;;) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(A.work_ctr, (unsigned long long)share);
        base = __shfl_sync(0xffffffffu, base, 0);
        if ((i64)base >= n_work) break;
        i64 w = (i64)base + lane / G;
        if (lane / G < share && w < n_work) {
            i64 rid = A.order ? A.order[w] : w;
            stage_extend_group<G, REG>(g, A.ix, A.opt, A.caps, A.B, rid, scr, smem, smat, ctr);
        }
        __syncwarp();
    }
    flush_counters(ctr, A.ctrs);
}

// stage 2 as a packed 16-bit anti-diagonal wavefront (ksw_wave.cuh): G lanes per read, two target rows per lane, the
// column state streamed lane to lane by warp shuffles.  Shared memory: (maxlen + 2) stream words per group.  Reads it
// cannot finish exactly (gap events, N bases, scores beyond 13 bits) are appended to B.retry_list for k_extend_group.
static size_t wave_smem_bytes(int maxlen, int G) { return ((size_t)WAVE_WORDS(maxlen, G) * 4 + 15) & ~(size_t)15; }
template <int G>
__global__ void __launch_bounds__(128, 6) k_extend_wave(const __grid_constant__ KArgs A)
{
    extern __shared__ __align__(16) u8 smem_raw[];
    const int lane = threadIdx.x & 31;
    const int gib = threadIdx.x / G;
    GroupCtx<G> g;
    g.gl = threadIdx.x % G;
    g.mask = G == 32 ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(G - 1)));
    u8 *smem = smem_raw + (size_t)gib * ((((size_t)WAVE_WORDS(A.caps.maxlen, G) * 4) + 15) & ~(size_t)15);
    u8 *scr = A.scratch + (size_t)(blockIdx.x * (128 / G) + gib) * A.scratch_stride;
    const int share = warp_share(A.n_work, 32 / G);
    CtrLocal ctr;
    for (;;) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(A.work_ctr, (unsigned long long)share);
        base = __shfl_sync(0xffffffffu, base, 0);
        if ((i64)base >= A.n_work) break;
        i64 w = (i64)base + lane / G;
        if (lane / G < share && w < A.n_work) {
            i64 rid = A.order ? A.order[w] : w;
            stage_extend_group<G, 2>(g, A.ix, A.opt, A.caps, A.B, rid, scr, smem, nullptr, ctr);
        }
        __syncwarp();
    }
    flush_counters(ctr, A.ctrs);
}

// gapped CIGARs: G lanes per queued hit (see finalize_group.cuh)
template <int G>
__global__ void __launch_bounds__(128, 6) k_finalize_dp(const __grid_constant__ KArgs A)
{
    extern __shared__ __align__(16) u8 smem_raw[];
    __shared__ i8 smat[32];
    if (threadIdx.x < 25) smat[threadIdx.x] = A.opt.mat[threadIdx.x];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int gib = threadIdx.x / G;
    GroupCtx<G> g;
    g.gl = threadIdx.x % G;
    g.mask = G == 32 ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(G - 1)));
    u8 *smem = smem_raw + (size_t)gib * findp_smem_bytes(A.caps.maxlen);
    u8 *scr = A.scratch + (size_t)(blockIdx.x * (128 / G) + gib) * A.scratch_stride;
    const i64 n_jobs = (i64)*A.B.n_dp_jobs < A.B.cap_dp_jobs ? (i64)*A.B.n_dp_jobs : A.B.cap_dp_jobs;
    const int share = warp_share(n_jobs, 32 / G);
    CtrLocal ctr;
    for (;;) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(A.work_ctr, (unsigned long long)share);
        base = __shfl_sync(0xffffffffu, base, 0);
        if ((i64)base >= n_jobs) break;
        i64 w = (i64)base + lane / G;
        if (lane / G < share && w < n_jobs) finalize_dp_job<G>(g, A.ix, A.opt, A.caps, A.B, A.B.dp_jobs[w], scr, smem, smat, ctr);
        __syncwarp();
    }
    flush_counters(ctr, A.ctrs);
}

// ASCII -> nt4 codes (nst_nt4_table, bwa/bntseq.c:46-63; mem_align1_core keeps codes < 4 as they are, bwa/bwamem.c:1087-1088)
__global__ void k_encode(const u8 *__restrict__ in, u8 *__restrict__ out, i64 n)
{
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i < n; i += (i64)gridDim.x * blockDim.x) {
        u8 c = in[i], v;
        switch (c) {
        case 'A': case 'a': v = 0; break; case 'C': case 'c': v = 1; break;
        case 'G': case 'g': v = 2; break; case 'T': case 't': v = 3; break;
        case '-': v = 5; break;
        default: v = c < 4 ? c : 4;
        }
        out[i] = v;
    }
}

__global__ void k_iota32(i32 *a, i64 n) { i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; if (i < n) a[i] = (i32)i; }
__global__ void k_clear_u32(u32 *a, i64 n) { i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; if (i < n) a[i] = 0; }
__global__ void k_clear_list(u32 *a, const i32 *list, i64 n) { i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; if (i < n) a[list[i]] = 0; }
__global__ void k_clear_list_dev(u32 *a, const i32 *list, const unsigned long long *n)
{
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < (i64)*n; i += (i64)gridDim.x * blockDim.x) a[list[i]] = 0;
}

// list the reads of [0,n) whose ovf bits intersect mask
__global__ void k_list_ovf(const u32 *__restrict__ ovf, i64 n, u32 mask, i32 *__restrict__ list, unsigned long long *cnt)
{
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (ovf[i] & mask) { unsigned long long o = atomicAdd(cnt, 1ull); list[o] = (i32)i; }
}

// reads that could not be processed: no hits
__global__ void k_drop_reads(ReadRec *__restrict__ rec, const i32 *__restrict__ list, i64 n)
{
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { ReadRec &R = rec[list[i]]; R.n_hits = R.n_cigar = R.n_md = 0; R.hit_off = 0; }
}

__global__ void k_counts(const ReadRec *__restrict__ rec, i64 n, i64 *__restrict__ nh, i64 *__restrict__ nc, i64 *__restrict__ nm)
{
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    nh[i] = rec[i].n_hits; nc[i] = rec[i].n_cigar; nm[i] = rec[i].n_md;
}

// compact the chunk's hits in read order; one warp per read
__global__ void k_gather(const ReadRec *__restrict__ rec, i64 n, Pools P, const i64 *__restrict__ oh, const i64 *__restrict__ oc,
                         const i64 *__restrict__ om, i64 base_h, i64 base_c, i64 base_m,
                         i64 *__restrict__ hit_off, b200_hit_t *__restrict__ hits, u32 *__restrict__ cigar, char *__restrict__ md)
{
    i64 r = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (r >= n) return;
    const ReadRec R = rec[r];
    if (lane == 0) hit_off[r] = base_h + oh[r];
    i64 c = oc[r], m = om[r];
    for (int i = 0; i < R.n_hits; ++i) {
        const b200_hit_t *src = P.hits + R.hit_off + i;
        b200_hit_t *dst = hits + oh[r] + i;
        const u32 *s32 = (const u32 *)src; u32 *d32 = (u32 *)dst;
        for (int k = lane; k < (int)(sizeof(b200_hit_t) / 4); k += 32) d32[k] = s32[k];
        __syncwarp();
        int ncg = src->n_cigar, nmd = src->md_len + 1;
        for (int k = lane; k < ncg; k += 32) cigar[c + k] = P.cigar[src->cigar_off + k];
        for (int k = lane; k < nmd; k += 32) md[m + k] = P.md[src->md_off + k];
        __syncwarp();
        if (lane == 0) { dst->cigar_off = base_c + c; dst->md_off = base_m + m; }
        c += ncg; m += nmd;
    }
}

// ---------------------------------------------------------------------------------------
// Pinned host buffers for results, recycled through a small free list: a results handle owns its buffers until
// b200_results_free() hands them back, so steady-state calls pay neither cudaHostAlloc nor page faults nor zero fill.
struct PinBuf {
    void *p = nullptr; size_t cap = 0;
};
struct PinPool {
    std::vector<PinBuf> free_list;
    std::mutex mu;
    PinBuf get(size_t bytes)
    {
        if (bytes == 0) bytes = 64;
        {
            std::lock_guard<std::mutex> g(mu);
            int best = -1;
            for (size_t i = 0; i < free_list.size(); ++i)
                if (free_list[i].cap >= bytes && (best < 0 || free_list[i].cap < free_list[best].cap)) best = (int)i;
            if (best >= 0) { PinBuf b = free_list[best]; free_list.erase(free_list.begin() + best); return b; }
        }
        PinBuf b; b.cap = bytes + (bytes >> 3) + 4096;
        if (cudaHostAlloc(&b.p, b.cap, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); b.p = nullptr; }
        if (!b.p) throw std::bad_alloc();
        return b;
    }
    void put(PinBuf b)
    {
        if (!b.p) return;
        std::lock_guard<std::mutex> g(mu);
        if (free_list.size() >= 16) {            // drop the smallest
            size_t k = 0;
            for (size_t i = 1; i < free_list.size(); ++i) if (free_list[i].cap < free_list[k].cap) k = i;
            if (free_list[k].cap < b.cap) { cudaFreeHost(free_list[k].p); free_list[k] = b; } else cudaFreeHost(b.p);
            return;
        }
        free_list.push_back(b);
    }
};
static PinPool &pin_pool() { static PinPool p; return p; }

struct HostResults {
    PinBuf b_off, b_hits, b_cigar, b_md;
    i64 n_reads = 0, n_hits = 0, n_cigar = 0, n_md = 0;
    i64 *hit_off() const { return (i64 *)b_off.p; }
    b200_hit_t *hits() const { return (b200_hit_t *)b_hits.p; }
    u32 *cigar() const { return (u32 *)b_cigar.p; }
    char *md() const { return (char *)b_md.p; }
    void alloc(i64 nr, i64 nh, i64 nc, i64 nm)
    {
        n_reads = nr; n_hits = nh; n_cigar = nc; n_md = nm;
        b_off = pin_pool().get((size_t)(nr + 1) * 8); b_hits = pin_pool().get((size_t)nh * sizeof(b200_hit_t));
        b_cigar = pin_pool().get((size_t)nc * 4); b_md = pin_pool().get((size_t)nm);
    }
    // streaming fill (b200_mem_align_batch): capacities are estimates, grown when a chunk does not fit
    static void grow(PinBuf &b, size_t used, size_t need, cudaStream_t copy_stream)
    {
        if (need <= b.cap) return;
        CU_CHECK(cudaStreamSynchronize(copy_stream));          // copies into the old buffer must have landed
        PinBuf nb = pin_pool().get(need + need / 2);
        if (used) memcpy(nb.p, b.p, used);
        pin_pool().put(b);
        b = nb;
    }
    ~HostResults() { pin_pool().put(b_off); pin_pool().put(b_hits); pin_pool().put(b_cigar); pin_pool().put(b_md); }
};

struct Engine {
    int device = -1, sms = 148;
    cudaStream_t st = nullptr, st_copy = nullptr, st_up = nullptr;      // kernels / result copies (D2H) / read uploads (H2D) that overlap the kernels
    cudaEvent_t ev[8], ev_copy, ev_up;
    // chunk buffers
    DevBuf wave_scratch, retry_list, packed, seedflag, seq_ascii, seq, seq_off, ids, ovf, rec, list, log_tab, scratch, spill_scratch, group_scratch, group_scratch2, dp_scratch, dp_scratch2, dp_jobs, sort_keys, sort_vals, sort_vals2, work, small;
    DevBuf p_intv, p_chain, p_seed, p_reg, p_hit, p_cigar, p_md;
    DevBuf nh, nc, nm, oh, oc, om, cubtmp, o_hit_off, o_hits, o_cigar, o_md;
    double pool_scale = 1.0;
    i64 pool_min[N_POOLS] = {0, 0, 0, 0, 0, 0, 0};       // per-pool capacity floor learnt from failed attempts (kept for later chunks)
    b200_stage_stats_t stats;

    void init()
    {
        int d = 0;
        CU_CHECK(cudaGetDevice(&d));
        if (st && d == device) return;
        device = d;
        CU_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, d));
        CU_CHECK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        CU_CHECK(cudaStreamCreateWithFlags(&st_copy, cudaStreamNonBlocking));
        CU_CHECK(cudaStreamCreateWithFlags(&st_up, cudaStreamNonBlocking));
        CU_CHECK(cudaEventCreateWithFlags(&ev_copy, cudaEventDisableTiming));
        CU_CHECK(cudaEventCreateWithFlags(&ev_up, cudaEventDisableTiming));
        if (const char *g = getenv("B200_L2_FETCH")) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(g));
        for (int i = 0; i < 8; ++i) CU_CHECK(cudaEventCreate(&ev[i]));
        memset(&stats, 0, sizeof(stats));
    }
};

static Engine &engine() { static thread_local Engine e; e.init(); return e; }

template <int STAGE>
static int stage_grid(int sms)
{
    int per = 0;
    CU_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_stage<STAGE>, 128, 0));
    if (per < 1) per = 1;
    if ((STAGE == 1 || STAGE == 3) && per > 5) per = 5;
    return sms * per;
}

struct ChunkCtx {
    Engine *E; const b200_index *idx; Opt opt; Caps caps, big; int maxlen;
    i64 n; // reads in this chunk
    Batch B; std::vector<double> logtab;
};

static size_t max4(size_t a, size_t b, size_t c, size_t d) { return std::max(std::max(a, b), std::max(c, d)); }

template <int STAGE>
static void launch_stage(Engine &E, KArgs &A, int grid, int tpb = 128)
{
    CU_CHECK(cudaMemsetAsync(A.work_ctr, 0, 8, E.st));
    k_stage<STAGE><<<grid, tpb, 0, E.st>>>(A);
    CU_CHECK(cudaGetLastError());
}
static double scratch_budget_bytes();
template <int STAGE, int MINB>
static void launch_stage_occ(Engine &E, KArgs &A, size_t stride)
{
    int per = 0;
    CU_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_stage_occ<STAGE, MINB>, 128, 0));
    if (per < 1) per = 1;
    int grid = E.sms * per;
    i64 fit = (i64)(scratch_budget_bytes() / (double)(stride * 128));
    if (fit < 1) fit = 1;
    if (grid > fit) grid = (int)fit;
    CU_CHECK(cudaMemsetAsync(A.work_ctr, 0, 8, E.st));
    k_stage_occ<STAGE, MINB><<<grid, 128, 0, E.st>>>(A);
    CU_CHECK(cudaGetLastError());
}

// Main-pass seeding with k_seed2: usable for batches of short reads on indexes below 2^36 symbols.
static const int SEED2_CAP = 16;          // long work-list entries per read in shared memory (a power of two: ring)
static const int SEED2_STRIDE = 40;       // interval slots per read in the pool
static bool seed2_enabled() { static int on = getenv("B200_SEED_V1") ? 0 : 1; return on != 0; }
static bool seed2_usable(const KArgs &A)
{
    return seed2_enabled() && !A.order && A.caps.maxlen <= 255 && A.ix.seq_len < (1ull << 36) && A.B.pool.cap[POOL_INTV] >= A.B.n_reads * (i64)SEED2_STRIDE;
}
template <int CAP, int MINB, int PHASED, int HB>
static void launch_seed2_hb(Engine &E, KArgs &A, int qw)
{
    size_t smem = (size_t)128 * (CAP * 12 + (qw + 1) * 4);
    CU_CHECK(cudaFuncSetAttribute(k_seed2<CAP, MINB, PHASED, HB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    int per = 0;
    CU_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_seed2<CAP, MINB, PHASED, HB>, 128, smem));
    if (per < 1) per = 1;
    { static const int occ = getenv("B200_OCC_SEED") ? atoi(getenv("B200_OCC_SEED")) : 0; if (occ > 0 && occ < per) per = occ; }
    k_seed2<CAP, MINB, PHASED, HB><<<E.sms * per, 128, smem, E.st>>>(A, E.packed.as<u32>(), qw, E.seedflag.as<u32>(), SEED2_STRIDE);
    CU_CHECK(cudaGetLastError());
    // then the third pass
    size_t smem3 = (size_t)128 * (qw + 1) * 4;
    int per3 = 0;
    CU_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per3, k_seed3, 128, smem3));
    if (per3 < 1) per3 = 1;
    CU_CHECK(cudaMemsetAsync(A.work_ctr, 0, 8, E.st));
    k_seed3<<<E.sms * per3, 128, smem3, E.st>>>(A, E.packed.as<u32>(), qw, SEED2_STRIDE);
}
template <int CAP, int MINB, int PHASED = 0>
static void launch_seed2_cap(Engine &E, KArgs &A, int qw)
{
    if (A.ix.seq_len < (1ull << 33)) launch_seed2_hb<CAP, MINB, PHASED, 1>(E, A, qw);
    else launch_seed2_hb<CAP, MINB, PHASED, 4>(E, A, qw);
}
static void launch_seed2(Engine &E, KArgs &A, int grid0 = 0)
{
    const i64 n = A.n_work;
    const int qw = (A.caps.maxlen + 15) / 16;
    E.packed.reserve((size_t)n * qw * 4 + 64); E.seedflag.reserve((size_t)n * 4 + 64);
    CU_CHECK(cudaMemsetAsync(E.seedflag.p, 0, (size_t)n * 4, E.st));
    k_pack_reads<<<(unsigned)((n * qw + 255) / 256), 256, 0, E.st>>>(A.B.seq, A.B.seq_off, n, qw, E.packed.as<u32>(), E.seedflag.as<u32>());
    k_set_u64<<<1, 1, 0, E.st>>>(A.B.pool.used + POOL_INTV, (unsigned long long)(n * SEED2_STRIDE));   // spill-pass allocations start after the fixed slots
    static int cap_sel = getenv("B200_SEED_CAP") ? atoi(getenv("B200_SEED_CAP")) : SEED2_CAP;
    CU_CHECK(cudaMemsetAsync(A.work_ctr, 0, 8, E.st));
    static int minb = getenv("B200_SEED_MINB") ? atoi(getenv("B200_SEED_MINB")) : 6;
    static int phased = getenv("B200_SEED_PHASED") ? atoi(getenv("B200_SEED_PHASED")) : 1;
    if (cap_sel == 8) launch_seed2_cap<8, 6, 1>(E, A, qw);
    else if (minb == 6 && phased) launch_seed2_cap<SEED2_CAP, 6, 1>(E, A, qw);
    else if (minb == 6) launch_seed2_cap<SEED2_CAP, 6>(E, A, qw);
    else launch_seed2_cap<SEED2_CAP, 6>(E, A, qw);
    CU_CHECK(cudaGetLastError());
    k_sort_intv<<<(unsigned)((n + 127) / 128), 128, 0, E.st>>>(A.B.rec, A.B.ovf, n, A.B.pool.intv);
    CU_CHECK(cudaGetLastError());
    // Reads the machine did not take (an N base, a work list or interval slot that overflowed) go through the reference-shaped
    // kernel right away, at full width and with the main-pass slots: only what overflows THOSE reaches the few-thread spill pass.
    // (A batch where 1 % of the reads hold an N would otherwise put 10^4 reads per chunk through the spill pass.)
    if (grid0 > 0 && A.scratch && E.list.p) {
        unsigned long long *cnt = A.work_ctr + 4;
        CU_CHECK(cudaMemsetAsync(cnt, 0, 8, E.st));
        k_list_ovf<<<(unsigned)((n + 255) / 256), 256, 0, E.st>>>(A.B.ovf, n, OVF_INTV, E.list.as<i32>(), cnt);
        k_clear_list_dev<<<64, 256, 0, E.st>>>(A.B.ovf, E.list.as<i32>(), cnt);
        KArgs A2 = A; A2.order = E.list.as<i32>(); A2.n_work_dev = cnt;
        CU_CHECK(cudaMemsetAsync(A2.work_ctr, 0, 8, E.st));
        k_stage<0><<<grid0, 128, 0, E.st>>>(A2);
        CU_CHECK(cudaGetLastError());
        E.stats.n_launches += 3;
    }
    E.stats.n_launches += 3;
}

// Prefix-chain table of an index (seed2.cuh): built once, on the first batch that can use it.
// K = min(B200_SEED_TAB_K (default 15), floor(log4(seq_len)), what half of the free HBM holds); 34 GB at K = 15
// (+ 5.7 GB of builder scratch, freed).  A batch whose options the table cannot serve (min_seed_len <= K, thresholds > 255)
// runs on the Occ blocks alone.
static SeedTab seed_tables(Engine &E, const b200_index *idx, const Opt &opt)
{
    SeedTab T; T.base = nullptr; T.K = 0; T.text = 0;
    std::lock_guard<std::mutex> g(idx->seedtab_mu);
    if (idx->seedtab_K < 0) {
        static const int want = getenv("B200_SEED_TAB_K") ? atoi(getenv("B200_SEED_TAB_K")) : 15;
        int K = want > 16 ? 16 : want;
        while (K > 0 && (1ull << (2 * K)) > idx->seq_len) --K;
        size_t free_b = 0, total_b = 0;
        CU_CHECK(cudaMemGetInfo(&free_b, &total_b));
        while (K > 0 && ((1ull << (2 * K)) * sizeof(ChainEnt) + seedtab_entries(K - 1) * sizeof(PIntv)) > (free_b + dev_pool().pooled) / 2) --K;
        if (idx->seq_len >= (1ull << 36) || K < 4) K = 0;
        if (K > 0) {
            void *p = nullptr, *lev = nullptr;
            const size_t bytes = (1ull << (2 * K)) * sizeof(ChainEnt), lev_bytes = seedtab_entries(K - 1) * sizeof(PIntv);
            if (cudaMalloc(&p, bytes) != cudaSuccess || cudaMalloc(&lev, lev_bytes) != cudaSuccess) {
                cudaGetLastError(); dev_pool().trim(0);
                if (!p) CU_CHECK(cudaMalloc(&p, bytes));
                CU_CHECK(cudaMalloc(&lev, lev_bytes));
            }
            for (int j = 1; j < K; ++j) {
                const u64 n = 1ull << (2 * j);
                k_seedtab_level<<<(unsigned)((n + 255) / 256), 256, 0, E.st>>>(idx->dev, (PIntv *)lev, j);
            }
            k_chain_build<<<(unsigned)(((1ull << (2 * K)) + 255) / 256), 256, 0, E.st>>>(idx->dev, (const PIntv *)lev, (ChainEnt *)p, K);
            CU_CHECK(cudaGetLastError());
            CU_CHECK(cudaStreamSynchronize(E.st));
            CU_CHECK(cudaFree(lev));
            idx->d_seedtab = p;
        }
        idx->seedtab_K = K;
        // may the machine follow size-one intervals through the text?  (full suffix array + text == the BWT's text)
        idx->seed_text_ok = 0;
        static const int text_on = getenv("B200_SEED_NO_TEXT") ? 0 : 1;
        if (text_on && idx->dev.sa_shift == 0 && idx->dev.n_sa > idx->seq_len) {
            unsigned long long *bad = E.small.as<unsigned long long>() + 500, h_bad = 1;
            CU_CHECK(cudaMemsetAsync(bad, 0, 8, E.st));
            k_verify_text<<<(unsigned)((idx->seq_len + 256) / 256), 256, 0, E.st>>>(idx->dev, bad);
            CU_CHECK(cudaMemcpyAsync(&h_bad, bad, 8, cudaMemcpyDeviceToHost, E.st));
            CU_CHECK(cudaStreamSynchronize(E.st));
            idx->seed_text_ok = h_bad == 0;
        }
    }
    if (seedtab_opt_ok(idx->seedtab_K, opt.min_seed_len, opt.split_width, opt.max_mem_intv)) { T.base = (const ChainEnt *)idx->d_seedtab; T.K = idx->seedtab_K; }
    T.text = idx->seed_text_ok;
    return T;
}

// the wavefront kernel computes scores as (match ? a : -b): the matrix must be what bwa_fill_scmat(a, b) makes (bwa/bwa.c:64-77)
static bool wave_opt_ok(const Opt &o)
{
    if (o.a <= 0 || o.b < 1 || o.e_del <= 0 || o.e_ins <= 0 || o.o_del < 0 || o.o_ins < 0) return false;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) if (o.mat[i * 5 + j] != (i == j ? o.a : -o.b)) return false;
    return o.a + o.b < 256 && o.o_del + o.e_del < 4000 && o.o_ins + o.e_ins < 4000;
}

static double scratch_budget_bytes()
{
    static const double budget_gb = getenv("B200_SCRATCH_GB") ? atof(getenv("B200_SCRATCH_GB")) : 16.0;
    return budget_gb * (double)(1ull << 30);
}

// Runs the four stages over `n_work` reads (all reads of the chunk, or the spill list).
// spill: 0 = main pass, 1 = spill pass (big slots, few blocks), 2 = last-resort pass (huge slots, one warp per block)
static void run_stages(Engine &E, KArgs A, int spill, float *ms4)
{
    int g[4];
    const int tpb = spill >= 2 ? 32 : 128;            // threads per block of the thread-per-read kernels
    if (!spill) { g[0] = stage_grid<0>(E.sms); g[1] = stage_grid<1>(E.sms); g[2] = stage_grid<2>(E.sms); g[3] = stage_grid<3>(E.sms); }
    else g[0] = g[1] = g[2] = g[3] = std::max<int>(1, (int)std::min<i64>((A.n_work + tpb - 1) / tpb, (i64)E.sms * 4));      // bounded by the scratch budget below
    Caps tc = A.caps;
    if (A.B.dp_jobs) tc.z = 64;          // gapped hits go to k_finalize_dp, the thread-per-read stage needs no direction matrix
    size_t stride = max4(seed_scratch_bytes(tc), chain_scratch_bytes(tc), extend_scratch_bytes(tc), finalize_scratch_bytes(tc));
    stride = (stride + 63) & ~(size_t)63;
    {   // long reads: the per-thread slot grows with the read length (hundreds of KB in the main pass, up to ~100 MB in the
        // spill pass), so the number of resident threads is bounded by a scratch budget instead of by occupancy alone
        i64 fit = (i64)(scratch_budget_bytes() / (double)(stride * tpb));
        if (fit < 1) fit = 1;
        for (int i = 0; i < 4; ++i) if (g[i] > fit) g[i] = (int)fit;
    }
    int gmax = std::max(std::max(g[0], g[1]), std::max(g[2], g[3]));
    static const int occ1 = getenv("B200_STAGE1_MINB") ? atoi(getenv("B200_STAGE1_MINB")) : 0, occ3 = getenv("B200_STAGE3_MINB") ? atoi(getenv("B200_STAGE3_MINB")) : 0;
    if (!spill && (occ1 || occ3)) gmax = std::max(gmax, (int)std::min<i64>((i64)E.sms * 16, (i64)(scratch_budget_bytes() / (double)(stride * tpb))));
    DevBuf &S = spill ? E.spill_scratch : E.scratch;
    S.reserve(stride * (size_t)gmax * tpb);
    A.scratch = S.as<u8>(); A.scratch_stride = stride;
    cudaEvent_t *ev = E.ev;
    CU_CHECK(cudaEventRecord(ev[0], E.st));
    if (!spill && seed2_usable(A)) launch_seed2(E, A, g[0]);
    else launch_stage<0>(E, A, g[0], tpb);
    CU_CHECK(cudaEventRecord(ev[1], E.st));
    if (!spill && occ1 == 12) launch_stage_occ<1, 12>(E, A, stride);
    else if (!spill && occ1 == 10) launch_stage_occ<1, 10>(E, A, stride);
    else if (!spill && occ1 == 16) launch_stage_occ<1, 16>(E, A, stride);
    else launch_stage<1>(E, A, g[1], tpb);
    CU_CHECK(cudaEventRecord(ev[2], E.st));
    {
        const int G = 8;
        static int reg_ok = getenv("B200_EXTEND_SMEM") ? 0 : 1;
        const bool use_reg = reg_ok && A.caps.maxlen + 1 <= G * EXT_REG_CMAX;      // DP state in registers (ksw_reg.cuh)
        size_t smem = use_reg ? 0 : (size_t)(128 / G) * group_smem_bytes(A.caps.maxlen);
        static int group_ok = getenv("B200_SCALAR_EXTEND") ? 0 : 1;
        if (group_ok && smem <= 200 * 1024) {
            // reads ordered by estimated work: heaviest first, equal work side by side in a warp
            const i32 *order = A.order;
            if (!spill && A.B.work && !A.order) {
                i64 n = A.n_work;
                E.sort_keys.reserve(n * 4 + 64); E.sort_vals.reserve(n * 4 + 64); E.sort_vals2.reserve(n * 4 + 64);
                k_iota32<<<(unsigned)((n + 255) / 256), 256, 0, E.st>>>(E.sort_vals.as<i32>(), n);
                size_t tb = 0;
                cub::DeviceRadixSort::SortPairsDescending(nullptr, tb, A.B.work, E.sort_keys.as<u32>(), E.sort_vals.as<i32>(), E.sort_vals2.as<i32>(), (int)n, 0, 24, E.st);
                E.cubtmp.reserve(tb);
                CU_CHECK(cub::DeviceRadixSort::SortPairsDescending(E.cubtmp.p, tb, A.B.work, E.sort_keys.as<u32>(), E.sort_vals.as<i32>(), E.sort_vals2.as<i32>(), (int)n, 0, 24, E.st));
                order = E.sort_vals2.as<i32>();
            }
            // main pass: the packed 16-bit wavefront kernel; what it hands back goes through the row-synchronous kernel below
            static const int wave_g = getenv("B200_WAVE_G") ? atoi(getenv("B200_WAVE_G")) : 4;
            const bool use_wave = !spill && wave_g > 0 && wave_opt_ok(A.opt) && A.caps.maxlen <= WAVE_MAXQ && A.B.retry_list &&
                                  (long)A.caps.maxlen * A.opt.a * 2 + std::max(A.opt.pen_clip5, A.opt.pen_clip3) + 16 < 8000;
            if (use_wave) {
                const int WG = wave_g == 8 ? 8 : wave_g == 2 ? 2 : 4;
                size_t wsmem = (size_t)(128 / WG) * wave_smem_bytes(A.caps.maxlen, WG);
                int per = 0;
                if (WG == 8) { CU_CHECK(cudaFuncSetAttribute(k_extend_wave<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsmem)); CU_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_extend_wave<8>, 128, wsmem)); }
                else if (WG == 2) { CU_CHECK(cudaFuncSetAttribute(k_extend_wave<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsmem)); CU_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_extend_wave<2>, 128, wsmem)); }
                else { CU_CHECK(cudaFuncSetAttribute(k_extend_wave<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsmem)); CU_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_extend_wave<4>, 128, wsmem)); }
                if (per < 1) per = 1;
                { static const int occ = getenv("B200_OCC_WAVE") ? atoi(getenv("B200_OCC_WAVE")) : 0; if (occ > 0 && occ < per) per = occ; }
                int grid = (int)std::min<i64>((i64)E.sms * per, std::max<i64>(1, (A.n_work + (128 / WG) - 1) / (128 / WG)));
                size_t gstride = (extend_group_scratch_bytes(A.caps) + 63) & ~(size_t)63;
                E.wave_scratch.reserve(gstride * (size_t)grid * (128 / WG));
                KArgs A2 = A; A2.scratch = E.wave_scratch.as<u8>(); A2.scratch_stride = gstride; A2.order = order;
                CU_CHECK(cudaMemsetAsync(A2.work_ctr, 0, 8, E.st));
                CU_CHECK(cudaMemsetAsync(A.B.n_retry, 0, 8, E.st));
                if (WG == 8) k_extend_wave<8><<<grid, 128, wsmem, E.st>>>(A2);
                else if (WG == 2) k_extend_wave<2><<<grid, 128, wsmem, E.st>>>(A2);
                else k_extend_wave<4><<<grid, 128, wsmem, E.st>>>(A2);
                CU_CHECK(cudaGetLastError());
                E.stats.n_launches += 1;
            }
            int per = 0;
            if (use_reg) CU_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_extend_group<G, true>, 128, 0));
            else {
                CU_CHECK(cudaFuncSetAttribute(k_extend_group<G, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                CU_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_extend_group<G, false>, 128, smem));
            }
            if (per < 1) per = 1;
            { static const int occ = getenv("B200_OCC_EXT") ? atoi(getenv("B200_OCC_EXT")) : 0; if (occ > 0 && occ < per) per = occ; }
            int grid = (int)std::min<i64>((i64)E.sms * per, std::max<i64>(1, (A.n_work + (128 / G) - 1) / (128 / G)));
            if (use_wave) grid = std::min(grid, E.sms);       // the retry list is short
            size_t gstride = (extend_group_scratch_bytes(A.caps) + 63) & ~(size_t)63;
            if (spill) grid = (int)std::max<i64>(1, std::min<i64>(grid, (i64)(scratch_budget_bytes() / (double)(gstride * (128 / G)))));
            (spill ? E.group_scratch2 : E.group_scratch).reserve(gstride * (size_t)grid * (128 / G));
            KArgs A2 = A; A2.scratch = (spill ? E.group_scratch2 : E.group_scratch).as<u8>(); A2.scratch_stride = gstride; A2.order = order;
            if (use_wave) { A2.order = A.B.retry_list; A2.n_work_dev = A.B.n_retry; A2.B.retry_list = nullptr; }
            CU_CHECK(cudaMemsetAsync(A2.work_ctr, 0, 8, E.st));
            if (use_reg) k_extend_group<G, true><<<grid, 128, 0, E.st>>>(A2);
            else k_extend_group<G, false><<<grid, 128, smem, E.st>>>(A2);
            CU_CHECK(cudaGetLastError());
        } else launch_stage<2>(E, A, g[2], tpb);
    }
    CU_CHECK(cudaEventRecord(ev[3], E.st));
    { KArgs At = A; At.caps = tc;
      if (!spill && occ3 == 8) launch_stage_occ<3, 8>(E, At, stride);
      else if (!spill && occ3 == 10) launch_stage_occ<3, 10>(E, At, stride);
      else if (!spill && occ3 == 12) launch_stage_occ<3, 12>(E, At, stride);
      else launch_stage<3>(E, At, g[3], tpb); }
    if (A.B.dp_jobs) {
        const int G = 8;
        size_t smem = (size_t)(128 / G) * findp_smem_bytes(A.caps.maxlen);
        int per = 0;
        CU_CHECK(cudaFuncSetAttribute(k_finalize_dp<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CU_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_finalize_dp<G>, 128, smem));
        if (per < 1) per = 1;
        int grid = spill ? std::min(E.sms * per, 64) : E.sms * per;
        size_t gstride = (findp_scratch_bytes(A.caps) + 63) & ~(size_t)63;
        if (spill) grid = (int)std::max<i64>(1, std::min<i64>(grid, (i64)(scratch_budget_bytes() / (double)(gstride * (128 / G)))));
        (spill ? E.dp_scratch2 : E.dp_scratch).reserve(gstride * (size_t)grid * (128 / G));
        KArgs A3 = A; A3.scratch = (spill ? E.dp_scratch2 : E.dp_scratch).as<u8>(); A3.scratch_stride = gstride;
        CU_CHECK(cudaMemsetAsync(A3.work_ctr, 0, 8, E.st));
        k_finalize_dp<G><<<grid, 128, smem, E.st>>>(A3);
        CU_CHECK(cudaGetLastError());
        E.stats.n_launches += 1;
    }
    CU_CHECK(cudaEventRecord(ev[4], E.st));
    E.stats.n_launches += 4;
    if (ms4) {
        CU_CHECK(cudaEventSynchronize(ev[4]));
        for (int i = 0; i < 4; ++i) { float t = 0; cudaEventElapsedTime(&t, ev[i], ev[i + 1]); ms4[i] += t; }
    }
}

// Device-side result of one chunk (compact, read order).
struct ChunkOut { i64 n_hits, n_cigar, n_md; };

// Processes reads [r0, r0+n) whose encoded bases / offsets / ids are already on the device
// (d_seq nt4, d_off relative to the batch start, d_ids).  Leaves compact results in E.o_* and returns their sizes.
static ChunkOut process_chunk(Engine &E, const b200_index *idx, const Opt &opt, int maxlen, const u8 *d_seq, const i64 *d_off,
                              const i64 *d_ids, i64 n, i64 base_h, i64 base_c, i64 base_m, const double *d_log, int n_log)
{
    Caps caps = default_caps(maxlen), big = big_caps(maxlen, opt);
    E.ovf.reserve(n * 4 + 64); E.rec.reserve(n * sizeof(ReadRec) + 64); E.list.reserve(n * 4 + 64); E.small.reserve(4096);
    unsigned long long *d_small = E.small.as<unsigned long long>();   // [0]=work ctr, [1]=list count, [8..15]=pool used, [16..]=DevCounters
    for (int attempt = 0; attempt < 8; ++attempt) {
        double sc = E.pool_scale;
        i64 cap[N_POOLS] = {(i64)(n * (seed2_enabled() ? SEED2_STRIDE + 8 : 24) * sc) + 65536, (i64)(n * 6 * sc) + 65536, (i64)(n * 24 * sc) + 65536, (i64)(n * 6 * sc) + 65536,
                            (i64)(n * 3 * sc) + 65536, (i64)(n * 12 * sc) + 65536, (i64)(n * 48 * sc) + 65536};
        for (int k = 0; k < N_POOLS; ++k) cap[k] = std::max(cap[k], E.pool_min[k]);      // demand seen by an earlier, failed attempt
        E.p_intv.reserve(cap[POOL_INTV] * sizeof(Intv)); E.p_chain.reserve(cap[POOL_CHAIN] * sizeof(Chain)); E.p_seed.reserve(cap[POOL_SEED] * sizeof(Seed));
        E.p_reg.reserve(cap[POOL_REG] * sizeof(Reg)); E.p_hit.reserve(cap[POOL_HIT] * sizeof(b200_hit_t)); E.p_cigar.reserve(cap[POOL_CIGAR] * 4);
        E.p_md.reserve(cap[POOL_MD]);
        CU_CHECK(cudaMemsetAsync(E.small.p, 0, 4096, E.st));
        k_clear_u32<<<(unsigned)((n + 255) / 256), 256, 0, E.st>>>(E.ovf.as<u32>(), n);
        KArgs A; memset(&A, 0, sizeof(A));
        A.ix = idx->dev; A.opt = opt; A.caps = caps;
        A.tab = seed_tables(E, idx, opt);
        A.B.n_reads = n; A.B.seq = d_seq; A.B.seq_off = d_off; A.B.hash_id = d_ids; A.B.ovf = E.ovf.as<u32>(); A.B.rec = E.rec.as<ReadRec>();
        Pools &P = A.B.pool;
        P.intv = E.p_intv.as<Intv>(); P.chains = E.p_chain.as<Chain>(); P.seeds = E.p_seed.as<Seed>(); P.regs = E.p_reg.as<Reg>();
        P.hits = E.p_hit.as<b200_hit_t>(); P.cigar = E.p_cigar.as<u32>(); P.md = E.p_md.as<char>();
        for (int k = 0; k < N_POOLS; ++k) P.cap[k] = cap[k];
        P.used = d_small + 8;
        { static int bal = getenv("B200_NO_BALANCE") ? 0 : 1; if (bal) { E.work.reserve(n * 4 + 64); A.B.work = E.work.as<u32>(); } }
        A.order = nullptr; A.n_work = n; A.work_ctr = d_small; A.ctrs = (DevCounters *)(d_small + 16);
        E.retry_list.reserve(n * 4 + 64);
        A.B.retry_list = E.retry_list.as<i32>(); A.B.n_retry = d_small + 3;
        {   // queue of hits that need a banded global alignment (drained by k_finalize_dp)
            size_t smem = (size_t)(128 / 8) * findp_smem_bytes(maxlen);
            static int dp_ok = getenv("B200_SCALAR_FINALIZE") ? 0 : 1;
            if (dp_ok && smem <= 200 * 1024) {
                E.dp_jobs.reserve((size_t)cap[POOL_HIT] * sizeof(DpJob));
                A.B.dp_jobs = E.dp_jobs.as<DpJob>(); A.B.n_dp_jobs = d_small + 2; A.B.cap_dp_jobs = cap[POOL_HIT];
            }
        }
        A.log_tab = d_log; A.n_log = n_log;
        float ms4[4] = {0, 0, 0, 0};
        static int trace = getenv("B200_TRACE") ? 1 : 0;
        double tw0 = wall_now();
        run_stages(E, A, 0, ms4);
        double tw1 = wall_now();
        // spill pass for reads that overflowed their scratch slot
        const u32 SCR = OVF_INTV | OVF_SEED | OVF_CHAIN | OVF_REG | OVF_OUT | OVF_SCRATCH;
        k_list_ovf<<<(unsigned)((n + 255) / 256), 256, 0, E.st>>>(E.ovf.as<u32>(), n, SCR, E.list.as<i32>(), d_small + 1);
        unsigned long long h_small[24];
        CU_CHECK(cudaMemcpyAsync(h_small, d_small, sizeof(h_small), cudaMemcpyDeviceToHost, E.st));
        CU_CHECK(cudaStreamSynchronize(E.st));
        i64 n_sp = (i64)h_small[1];
        for (int tier = 1; tier <= 2 && n_sp; ++tier) {
            // tier 1: the spill pass (big slots); tier 2: reads that overflow even those (repeat-saturated: up to max_occ seeds
            // per interval) get slots sized for the worst case the reference accepts, a warp at a time
            if (tier == 1) E.stats.n_overflow += (u64)n_sp;
            KArgs S = A; S.caps = tier == 1 ? big : big_caps(maxlen, opt, 2); S.order = E.list.as<i32>(); S.n_work = n_sp;
            CU_CHECK(cudaMemsetAsync(d_small + 2, 0, 8, E.st));       // the previous pass's DP queue has been drained
            if (trace) {
                std::vector<i32> lst(n_sp); std::vector<u32> fl(n);
                CU_CHECK(cudaMemcpy(lst.data(), E.list.p, n_sp * 4, cudaMemcpyDeviceToHost));
                CU_CHECK(cudaMemcpy(fl.data(), E.ovf.p, n * 4, cudaMemcpyDeviceToHost));
                int cnt[8] = {0};
                for (i64 k = 0; k < n_sp; ++k) for (int b = 0; b < 8; ++b) if (fl[lst[k]] >> b & 1) ++cnt[b];
                fprintf(stderr, "[b200 trace] spill tier %d reasons: intv %d seed %d chain %d reg %d out %d scratch %d pool %d\n", tier, cnt[0], cnt[1], cnt[2], cnt[3], cnt[4], cnt[5], cnt[6]);
            }
            k_clear_list<<<(unsigned)((n_sp + 255) / 256), 256, 0, E.st>>>(E.ovf.as<u32>(), E.list.as<i32>(), n_sp);
            run_stages(E, S, tier, ms4);
            E.stats.n_launches += 1;
            if (tier == 1) {        // anything left that is not a pool problem?
                CU_CHECK(cudaMemsetAsync(d_small + 1, 0, 8, E.st));
                k_list_ovf<<<(unsigned)((n + 255) / 256), 256, 0, E.st>>>(E.ovf.as<u32>(), n, SCR, E.list.as<i32>(), d_small + 1);
                CU_CHECK(cudaMemcpyAsync(h_small, d_small, sizeof(h_small), cudaMemcpyDeviceToHost, E.st));
                CU_CHECK(cudaStreamSynchronize(E.st));
                n_sp = (i64)h_small[1];
            }
        }
        double tw2 = wall_now();
        // any read still flagged?  pool exhaustion => retry the chunk with larger pools; anything else is a hard limit
        CU_CHECK(cudaMemsetAsync(d_small + 1, 0, 8, E.st));
        k_list_ovf<<<(unsigned)((n + 255) / 256), 256, 0, E.st>>>(E.ovf.as<u32>(), n, 0xffffffffu, E.list.as<i32>(), d_small + 1);
        CU_CHECK(cudaMemcpyAsync(h_small, d_small, sizeof(h_small), cudaMemcpyDeviceToHost, E.st));
        CU_CHECK(cudaStreamSynchronize(E.st));
        E.stats.n_launches += 4;
        if (h_small[1]) {
            bool pool_full = false;
            // a full pool: the bump counter kept counting, so it tells what this attempt would have needed
            for (int k = 0; k < N_POOLS; ++k)
                if ((i64)h_small[8 + k] > cap[k]) { pool_full = true; E.pool_min[k] = std::max<i64>(2 * cap[k], (i64)h_small[8 + k] + (i64)h_small[8 + k] / 4 + 65536); }
            if ((i64)h_small[2] > cap[POOL_HIT]) { pool_full = true; E.pool_min[POOL_HIT] = std::max<i64>(2 * cap[POOL_HIT], (i64)h_small[2] + 65536); }   // the DP job queue is sized like the hit pool
            if (pool_full && attempt < 7) continue;
            // A read that still does not fit (more than 2^22 seeds, or a pool that six doublings did not satisfy) is
            // reported without hits instead of failing the whole batch: the other reads of the chunk are unaffected.
            k_drop_reads<<<(unsigned)((h_small[1] + 255) / 256), 256, 0, E.st>>>(E.rec.as<ReadRec>(), E.list.as<i32>(), (i64)h_small[1]);
            E.stats.n_failed += h_small[1];
            set_error("b200_mem_align_batch: " + std::to_string((unsigned long long)h_small[1]) + " read(s) exceeded the working-set limits and are reported unaligned");
        }
        E.stats.ms_seed += ms4[0]; E.stats.ms_chain += ms4[1]; E.stats.ms_extend += ms4[2]; E.stats.ms_finalize += ms4[3];
        const unsigned long long *c = h_small + 16;
        E.stats.occ_blocks += c[0]; E.stats.sa_reads += c[1]; E.stats.ref_bytes += c[2]; E.stats.sw_cells += c[3]; E.stats.n_ext += c[4]; E.stats.n_global += c[5];
        E.stats.tab_lookups_lo += c[6]; E.stats.tab_lookups_hi += c[7];
        E.stats.ext_fallback += h_small[3];
        // compact in read order
        E.nh.reserve(n * 8 + 8); E.nc.reserve(n * 8 + 8); E.nm.reserve(n * 8 + 8); E.oh.reserve(n * 8 + 16); E.oc.reserve(n * 8 + 16); E.om.reserve(n * 8 + 16);
        k_counts<<<(unsigned)((n + 255) / 256), 256, 0, E.st>>>(E.rec.as<ReadRec>(), n, E.nh.as<i64>(), E.nc.as<i64>(), E.nm.as<i64>());
        size_t tb = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tb, E.nh.as<i64>(), E.oh.as<i64>(), (int)n, E.st);
        E.cubtmp.reserve(tb);
        CU_CHECK(cub::DeviceScan::ExclusiveSum(E.cubtmp.p, tb, E.nh.as<i64>(), E.oh.as<i64>(), (int)n, E.st));
        CU_CHECK(cub::DeviceScan::ExclusiveSum(E.cubtmp.p, tb, E.nc.as<i64>(), E.oc.as<i64>(), (int)n, E.st));
        CU_CHECK(cub::DeviceScan::ExclusiveSum(E.cubtmp.p, tb, E.nm.as<i64>(), E.om.as<i64>(), (int)n, E.st));
        ChunkOut out;
        out.n_hits = (i64)h_small[8 + POOL_HIT]; out.n_cigar = (i64)h_small[8 + POOL_CIGAR]; out.n_md = (i64)h_small[8 + POOL_MD];
        // pools may hold abandoned allocations of spilled reads: exact totals come from the scans
        i64 last[3], lastn[3];
        CU_CHECK(cudaMemcpyAsync(&last[0], E.oh.as<i64>() + (n - 1), 8, cudaMemcpyDeviceToHost, E.st));
        CU_CHECK(cudaMemcpyAsync(&last[1], E.oc.as<i64>() + (n - 1), 8, cudaMemcpyDeviceToHost, E.st));
        CU_CHECK(cudaMemcpyAsync(&last[2], E.om.as<i64>() + (n - 1), 8, cudaMemcpyDeviceToHost, E.st));
        CU_CHECK(cudaMemcpyAsync(&lastn[0], E.nh.as<i64>() + (n - 1), 8, cudaMemcpyDeviceToHost, E.st));
        CU_CHECK(cudaMemcpyAsync(&lastn[1], E.nc.as<i64>() + (n - 1), 8, cudaMemcpyDeviceToHost, E.st));
        CU_CHECK(cudaMemcpyAsync(&lastn[2], E.nm.as<i64>() + (n - 1), 8, cudaMemcpyDeviceToHost, E.st));
        CU_CHECK(cudaStreamSynchronize(E.st));
        out.n_hits = last[0] + lastn[0]; out.n_cigar = last[1] + lastn[1]; out.n_md = last[2] + lastn[2];
        E.o_hit_off.reserve(n * 8 + 8); E.o_hits.reserve(out.n_hits * sizeof(b200_hit_t) + 64); E.o_cigar.reserve(out.n_cigar * 4 + 64); E.o_md.reserve(out.n_md + 64);
        k_gather<<<(unsigned)((n * 32 + 255) / 256), 256, 0, E.st>>>(E.rec.as<ReadRec>(), n, P, E.oh.as<i64>(), E.oc.as<i64>(), E.om.as<i64>(), base_h, base_c, base_m,
                                                                    E.o_hit_off.as<i64>(), E.o_hits.as<b200_hit_t>(), E.o_cigar.as<u32>(), E.o_md.as<char>());
        CU_CHECK(cudaGetLastError());
        E.stats.n_launches += 5;
        if (trace) { CU_CHECK(cudaStreamSynchronize(E.st)); double tw3 = wall_now();
            fprintf(stderr, "[b200 trace] chunk n=%lld stages %.1f ms, spill(%lld reads) %.1f ms, compact %.1f ms; stage ms %.1f %.1f %.1f %.1f\n", (long long)n, 1e3 * (tw1 - tw0), (long long)n_sp, 1e3 * (tw2 - tw1), 1e3 * (tw3 - tw2), ms4[0], ms4[1], ms4[2], ms4[3]); }
        return out;
    }
    throw std::runtime_error("pool growth did not converge");
}

static double wall_now() { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

static i64 chunk_reads()
{
    const char *e = getenv("B200_CHUNK");
    i64 v = e ? atoll(e) : (1 << 21);       // 2 M reads: ~3 ms of kernel tails per chunk argue for few, large chunks
    return v < 1024 ? 1024 : v;
}

} // namespace b200

struct b200_results { HostResults r; };

struct b200_batch {
    HostResults *sink = nullptr;      // when set, b200_batch_run copies every chunk's result to the host while the next chunk runs
    const b200_index *idx; Opt opt; i64 n; int maxlen;
    const i64 *h_off = nullptr;       // the caller's offsets; only dereferenced inside b200_mem_align_batch (lazy upload)
    DevBuf d_seq, d_off, d_ids, d_log; int n_log;
    // lazy upload (b200_mem_align_batch): the bases of chunk i are copied on the copy stream and encoded right before chunk i
    // runs, so chunk 0 starts after its own 1/k of the transfer; `uploaded` = bytes already issued
    const char *lazy_src = nullptr; DevBuf lazy_ascii; i64 uploaded = 0, total_bytes = 0; bool lazy = false;
    // device-resident compact results of the last run, per chunk
    struct Chunk { i64 r0, n; ChunkOut out; DevBuf hit_off, hits, cigar, md; };
    std::vector<Chunk *> chunks;
    ~b200_batch() { for (auto c : chunks) delete c; }
};

static thread_local bool g_lazy_upload = false;

extern "C" {

static int check_reads(const b200_mem_opt_t *opt, i64 n, const int64_t *off, int *maxlen_out)
{
    // branch-free min / max so that the compiler vectorises the scan (10^7 offsets per call on the end-to-end path)
    i64 mn = 0, mx = 1;
    for (i64 i = 0; i < n; ++i) {
        const i64 l = off[i + 1] - off[i];
        mn = l < mn ? l : mn;
        mx = l > mx ? l : mx;
    }
    if (mn < 0 || mx > (1 << 20)) return fail(B200_ERR_ARG, "bad read length");
    const int maxlen = (int)mx;
    // reads long enough for mem_flt_chained_seeds (bwa/bwamem.c:624-641, > ~730 bp) are handled in stage_chain (seedsw.cuh)
    (void)opt;
    *maxlen_out = maxlen;
    return B200_OK;
}

int b200_batch_create(const b200_index_t *idx, const b200_mem_opt_t *opt, int64_t n, const char *seqs, const int64_t *off,
                      const int64_t *ids, b200_batch_t **out)
{
    if (!idx || !opt || !out || n < 0 || (n && (!seqs || !off))) return fail(B200_ERR_ARG, "bad argument");
    *out = nullptr;
    static const int64_t zero_off[1] = {0};
    if (n == 0 && !off) off = zero_off;
    int maxlen = 1, rc;
    if ((rc = check_reads(opt, n, off, &maxlen)) != B200_OK) return rc;
    b200_batch *b = new b200_batch;
    try {
        Engine &E = engine();
        b->idx = idx; b->opt = opt_from_abi(*opt); b->n = n; b->maxlen = maxlen;
        b->h_off = off;
        i64 base = off[0], total = off[n] - base;
        std::vector<i64> rel;
        const i64 *rel_src = off;                 // offsets relative to the first base: the caller's array when it starts at 0
        if (base != 0) { rel.resize(n + 1); for (i64 i = 0; i <= n; ++i) rel[i] = off[i] - base; rel_src = rel.data(); }
        std::vector<i64> myids;
        if (!ids) { myids.resize(n); for (i64 i = 0; i < n; ++i) myids[i] = lrand48(); ids = myids.data(); }
        b->d_seq.reserve(total + 64); b->d_off.reserve((n + 1) * 8); b->d_ids.reserve(n * 8 + 8);
        b->total_bytes = total;
        if (g_lazy_upload) {
            b->lazy = true; b->lazy_src = seqs + base; b->lazy_ascii.reserve(total + 64);
            // the first chunk's bases start travelling now, behind the offsets / ids / log table set-up below
            const i64 first = off[std::min<i64>(n, chunk_reads())] - base;
            if (first > 0) { CU_CHECK(cudaMemcpyAsync(b->lazy_ascii.p, b->lazy_src, first, cudaMemcpyHostToDevice, E.st_up)); b->uploaded = first; }
        } else {
            E.seq_ascii.reserve(total + 64);
            CU_CHECK(cudaMemcpyAsync(E.seq_ascii.p, seqs + base, total, cudaMemcpyHostToDevice, E.st));
        }
        CU_CHECK(cudaMemcpyAsync(b->d_off.p, rel_src, (n + 1) * 8, cudaMemcpyHostToDevice, E.st));
        CU_CHECK(cudaMemcpyAsync(b->d_ids.p, ids, n * 8, cudaMemcpyHostToDevice, E.st));
        if (total && !b->lazy) k_encode<<<E.sms * 8, 256, 0, E.st>>>(E.seq_ascii.as<u8>(), b->d_seq.as<u8>(), total);
        std::vector<double> lt = make_log_table(maxlen, b->opt);
        b->n_log = (int)lt.size();
        b->d_log.reserve(lt.size() * 8);
        CU_CHECK(cudaMemcpyAsync(b->d_log.p, lt.data(), lt.size() * 8, cudaMemcpyHostToDevice, E.st));
        CU_CHECK(cudaStreamSynchronize(E.st));
    } catch (const std::exception &e) { delete b; return fail(B200_ERR_CUDA, e.what()); }
    *out = b;
    return B200_OK;
}

int b200_batch_run(b200_batch_t *b, int *n_launches)
{
    if (!b) return fail(B200_ERR_ARG, "bad argument");
    try {
        Engine &E = engine();
        memset(&E.stats, 0, sizeof(E.stats));
        CU_CHECK(cudaEventRecord(E.ev[6], E.st));
        i64 CH = chunk_reads(), bh = 0, bc = 0, bm = 0;
        // Chunk plan: a quarter-size first chunk (the kernels start after a quarter of the first transfer) and a tapering tail
        // (the result copy of the last chunk is the only one no kernel overlaps), full chunks in between.
        std::vector<i64> plan;
        {
            i64 rem = b->n;
            const i64 q = std::max<i64>(CH / 4, 1024);
            static const int taper = getenv("B200_CHUNK_TAPER") ? atoi(getenv("B200_CHUNK_TAPER")) : 0;   // measured: every extra chunk costs ~3 ms of kernel tails, more than the shorter head / tail copies save
            if (taper == 2 && rem > 3 * CH) {
                // same number of chunks, but a 3/4 first chunk (its upload is the only one no kernel hides) and a half-size last one
                // (its result copy is the only one no kernel hides); the middle chunks share the rest
                const i64 first = (CH * 3 / 4) & ~(i64)1023, last = (CH / 2) & ~(i64)1023, mid_total = rem - first - last;
                const i64 k = std::max<i64>(1, (mid_total * 4 + CH * 5 - 1) / (CH * 5));
                const i64 mid = ((mid_total + k - 1) / k + 1023) & ~(i64)1023;
                plan.push_back(first); rem -= first;
                while (rem > last) { i64 t = std::min(mid, rem - last); plan.push_back(t); rem -= t; }
                plan.push_back(rem); rem = 0;
            }
            if (!taper || taper == 2) { while (rem > 0) { i64 t = std::min(CH, rem); plan.push_back(t); rem -= t; } }
            if (rem > CH) { plan.push_back(q); rem -= q; }
            while (rem > CH + CH / 2) { plan.push_back(CH); rem -= CH; }
            while (rem > q) { i64 t = std::max(q, (rem / 2 + 1023) & ~(i64)1023); if (t > rem) t = rem; plan.push_back(t); rem -= t; }
            if (rem > 0) plan.push_back(rem);
        }
        size_t ci = 0;
        for (i64 r0 = 0; r0 < b->n; r0 += plan[ci], ++ci) {
            i64 n = plan[ci];
            if (b->lazy) {
                // this chunk's bases must have been issued (chunk 0: now; later chunks: while the previous one ran); the kernels
                // wait for exactly those copies -- the event is recorded BEFORE the next chunk's transfer is queued behind it
                const i64 lo = b->h_off[r0] - b->h_off[0], hi = b->h_off[r0 + n] - b->h_off[0];
                if (hi > b->uploaded) {
                    CU_CHECK(cudaMemcpyAsync(b->lazy_ascii.as<u8>() + b->uploaded, b->lazy_src + b->uploaded, hi - b->uploaded, cudaMemcpyHostToDevice, E.st_up));
                    b->uploaded = hi;
                }
                CU_CHECK(cudaEventRecord(E.ev_up, E.st_up));
                CU_CHECK(cudaStreamWaitEvent(E.st, E.ev_up, 0));
                // the next chunk's bases travel while this chunk's kernels run
                const i64 upto = b->h_off[std::min(b->n, r0 + n + (ci + 1 < plan.size() ? plan[ci + 1] : 0))] - b->h_off[0];
                if (upto > b->uploaded) {
                    CU_CHECK(cudaMemcpyAsync(b->lazy_ascii.as<u8>() + b->uploaded, b->lazy_src + b->uploaded, upto - b->uploaded, cudaMemcpyHostToDevice, E.st_up));
                    b->uploaded = upto;
                }
                if (hi > lo) k_encode<<<E.sms * 4, 256, 0, E.st>>>(b->lazy_ascii.as<u8>() + lo, b->d_seq.as<u8>() + lo, hi - lo);
            }
            b200_batch::Chunk *c = ci < b->chunks.size() ? b->chunks[ci] : nullptr;
            if (c) {    // hand the previous run's result buffers back to the engine so they are reused, not re-allocated
                std::swap(c->hit_off.p, E.o_hit_off.p); std::swap(c->hit_off.cap, E.o_hit_off.cap);
                std::swap(c->hits.p, E.o_hits.p); std::swap(c->hits.cap, E.o_hits.cap);
                std::swap(c->cigar.p, E.o_cigar.p); std::swap(c->cigar.cap, E.o_cigar.cap);
                std::swap(c->md.p, E.o_md.p); std::swap(c->md.cap, E.o_md.cap);
            } else { c = new b200_batch::Chunk; b->chunks.push_back(c); }
            ChunkOut o = process_chunk(E, b->idx, b->opt, b->maxlen, b->d_seq.as<u8>(), b->d_off.as<i64>() + r0, b->d_ids.as<i64>() + r0, n, bh, bc, bm,
                                       b->d_log.as<double>(), b->n_log);
            c->r0 = r0; c->n = n; c->out = o;
            // keep the compact result on the device (swap buffers with the engine's output buffers)
            std::swap(c->hit_off.p, E.o_hit_off.p); std::swap(c->hit_off.cap, E.o_hit_off.cap);
            std::swap(c->hits.p, E.o_hits.p); std::swap(c->hits.cap, E.o_hits.cap);
            std::swap(c->cigar.p, E.o_cigar.p); std::swap(c->cigar.cap, E.o_cigar.cap);
            std::swap(c->md.p, E.o_md.p); std::swap(c->md.cap, E.o_md.cap);
            if (b->sink) {
                HostResults &H = *b->sink;
                HostResults::grow(H.b_hits, (size_t)bh * sizeof(b200_hit_t), (size_t)(bh + o.n_hits) * sizeof(b200_hit_t), E.st_copy);
                HostResults::grow(H.b_cigar, (size_t)bc * 4, (size_t)(bc + o.n_cigar) * 4, E.st_copy);
                HostResults::grow(H.b_md, (size_t)bm, (size_t)(bm + o.n_md), E.st_copy);
                CU_CHECK(cudaEventRecord(E.ev_copy, E.st));
                CU_CHECK(cudaStreamWaitEvent(E.st_copy, E.ev_copy, 0));
                CU_CHECK(cudaMemcpyAsync(H.hit_off() + r0, c->hit_off.p, n * 8, cudaMemcpyDeviceToHost, E.st_copy));
                if (o.n_hits) CU_CHECK(cudaMemcpyAsync(H.hits() + bh, c->hits.p, o.n_hits * sizeof(b200_hit_t), cudaMemcpyDeviceToHost, E.st_copy));
                if (o.n_cigar) CU_CHECK(cudaMemcpyAsync(H.cigar() + bc, c->cigar.p, o.n_cigar * 4, cudaMemcpyDeviceToHost, E.st_copy));
                if (o.n_md) CU_CHECK(cudaMemcpyAsync(H.md() + bm, c->md.p, o.n_md, cudaMemcpyDeviceToHost, E.st_copy));
            }
            bh += o.n_hits; bc += o.n_cigar; bm += o.n_md;
        }
        if (b->sink) {
            CU_CHECK(cudaStreamSynchronize(E.st_copy));
            HostResults &H = *b->sink;
            H.n_reads = b->n; H.n_hits = bh; H.n_cigar = bc; H.n_md = bm;
            H.hit_off()[b->n] = bh;
        }
        CU_CHECK(cudaEventRecord(E.ev[7], E.st));
        CU_CHECK(cudaEventSynchronize(E.ev[7]));
        cudaEventElapsedTime(&E.stats.ms_total, E.ev[6], E.ev[7]);
        if (n_launches) *n_launches = E.stats.n_launches;
    } catch (const std::exception &e) { return fail(B200_ERR_CUDA, e.what()); }
    return B200_OK;
}

int b200_batch_fetch(b200_batch_t *b, b200_results_t **out)
{
    if (!b || !out) return fail(B200_ERR_ARG, "bad argument");
    *out = nullptr;
    b200_results *R = new b200_results;
    try {
        Engine &E = engine();
        i64 th = 0, tc = 0, tm = 0;
        for (auto c : b->chunks) { th += c->out.n_hits; tc += c->out.n_cigar; tm += c->out.n_md; }
        HostResults &H = R->r;
        H.alloc(b->n, th, tc, tm);
        i64 ph = 0, pc = 0, pm = 0;
        for (auto c : b->chunks) {
            CU_CHECK(cudaMemcpyAsync(H.hit_off() + c->r0, c->hit_off.p, c->n * 8, cudaMemcpyDeviceToHost, E.st));
            if (c->out.n_hits) CU_CHECK(cudaMemcpyAsync(H.hits() + ph, c->hits.p, c->out.n_hits * sizeof(b200_hit_t), cudaMemcpyDeviceToHost, E.st));
            if (c->out.n_cigar) CU_CHECK(cudaMemcpyAsync(H.cigar() + pc, c->cigar.p, c->out.n_cigar * 4, cudaMemcpyDeviceToHost, E.st));
            if (c->out.n_md) CU_CHECK(cudaMemcpyAsync(H.md() + pm, c->md.p, c->out.n_md, cudaMemcpyDeviceToHost, E.st));
            ph += c->out.n_hits; pc += c->out.n_cigar; pm += c->out.n_md;
        }
        CU_CHECK(cudaStreamSynchronize(E.st));
        H.hit_off()[b->n] = th;
    } catch (const std::exception &e) { delete R; return fail(B200_ERR_CUDA, e.what()); }
    *out = R;
    return B200_OK;
}

void b200_batch_destroy(b200_batch_t *b) { delete b; }

int b200_mem_align_batch(const b200_index_t *idx, const b200_mem_opt_t *opt, int64_t n, const char *seqs, const int64_t *off,
                         const int64_t *ids, b200_results_t **out)
{
    if (!out) return fail(B200_ERR_ARG, "out is NULL");
    *out = nullptr;
    b200_batch_t *b = nullptr;
    static int trace = getenv("B200_TRACE") ? 1 : 0;
    double t0 = wall_now();
    g_lazy_upload = getenv("B200_EAGER_UPLOAD") ? false : true;
    int rc = b200_batch_create(idx, opt, n, seqs, off, ids, &b);
    g_lazy_upload = false;
    if (rc != B200_OK) return rc;
    double t1 = wall_now();
    b200_results *R = new b200_results;
    try {
        // results stream to the host chunk by chunk (copy stream) while the next chunk's kernels run
        HostResults &H = R->r;
        H.b_off = pin_pool().get((size_t)(n + 1) * 8);
        H.b_hits = pin_pool().get((size_t)(2 * n + 1024) * sizeof(b200_hit_t));
        H.b_cigar = pin_pool().get((size_t)(4 * n + 1024) * 4);
        H.b_md = pin_pool().get((size_t)(16 * n + 1024));
        H.hit_off()[0] = 0;
        b->sink = &H;
    } catch (const std::exception &e) { delete R; b200_batch_destroy(b); return fail(B200_ERR_NOMEM, e.what()); }
    rc = b200_batch_run(b, nullptr);
    if (rc != B200_OK) {
        // a failure in the middle of the stream: copies into R's pinned buffers / the batch's device buffers may still be
        // in flight, and both go back to their pools below
        try { Engine &E = engine(); cudaStreamSynchronize(E.st_up); cudaStreamSynchronize(E.st_copy); cudaStreamSynchronize(E.st); } catch (...) {}
    }
    double t2 = wall_now();
    b->sink = nullptr;
    b200_batch_destroy(b);
    if (rc == B200_OK) *out = R; else delete R;
    if (trace) fprintf(stderr, "[b200 trace] align_batch n=%lld: create %.1f ms, run+copy %.1f ms, destroy %.1f ms\n", (long long)n,
                       1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (wall_now() - t2));
    return rc;
}

int b200_results_view(const b200_results_t *res, b200_results_view_t *v)
{
    if (!res || !v) return fail(B200_ERR_ARG, "bad argument");
    const HostResults &H = res->r;
    v->n_reads = H.n_reads; v->hit_off = H.hit_off(); v->hits = H.hits(); v->cigar = H.cigar(); v->md = H.md();
    v->n_hits = H.n_hits; v->n_cigar = H.n_cigar; v->n_md = H.n_md;
    return B200_OK;
}

void b200_results_free(b200_results_t *res) { delete res; }

namespace { std::mutex g_host_mu; std::unordered_map<void *, PinBuf> g_host_bufs; }
int b200_host_alloc(size_t bytes, void **out)
{
    if (!out) return fail(B200_ERR_ARG, "out is NULL");
    *out = nullptr;
    try {
        engine();                                   // a device context must exist before cudaHostAlloc
        PinBuf b = pin_pool().get(bytes);
        std::lock_guard<std::mutex> g(g_host_mu);
        g_host_bufs[b.p] = b;
        *out = b.p;
    } catch (const std::bad_alloc &) { return fail(B200_ERR_NOMEM, "pinned host allocation failed"); }
    catch (const std::exception &e) { return fail(B200_ERR_CUDA, e.what()); }
    return B200_OK;
}
void b200_host_free(void *p)
{
    if (!p) return;
    PinBuf b;
    {
        std::lock_guard<std::mutex> g(g_host_mu);
        auto it = g_host_bufs.find(p);
        if (it == g_host_bufs.end()) return;
        b = it->second; g_host_bufs.erase(it);
    }
    pin_pool().put(b);
}

int b200_last_stats(b200_stage_stats_t *out)
{
    if (!out) return fail(B200_ERR_ARG, "bad argument");
    try { *out = engine().stats; } catch (const std::exception &e) { return fail(B200_ERR_CUDA, e.what()); }
    return B200_OK;
}

// ---- stage dump for tests: intervals of every read (mem_collect_intv) -------------------
int b200_debug_collect_intv(const b200_index_t *idx, const b200_mem_opt_t *opt, int64_t n, const char *seqs, const int64_t *off,
                            int64_t **intv_off, b200_intv_t **intv)
{
    if (!idx || !opt || !intv_off || !intv) return fail(B200_ERR_ARG, "bad argument");
    b200_batch_t *b = nullptr;
    int rc = b200_batch_create(idx, opt, n, seqs, off, nullptr, &b);
    if (rc != B200_OK) return rc;
    try {
        Engine &E = engine();
        Caps big = big_caps(b->maxlen, b->opt);
        E.ovf.reserve(n * 4 + 64); E.rec.reserve(n * sizeof(ReadRec) + 64); E.small.reserve(4096);
        i64 cap = n * (i64)(big.intv + SEED2_STRIDE) + 1024;
        E.p_intv.reserve(cap * sizeof(Intv));
        CU_CHECK(cudaMemsetAsync(E.small.p, 0, 4096, E.st));
        k_clear_u32<<<(unsigned)((n + 255) / 256), 256, 0, E.st>>>(E.ovf.as<u32>(), n);
        KArgs A; memset(&A, 0, sizeof(A));
        unsigned long long *d_small = E.small.as<unsigned long long>();
        A.ix = idx->dev; A.opt = b->opt; A.caps = big;
        A.tab = seed_tables(E, idx, b->opt);
        A.tab.text = 0;                                    // the reference's coordinates (x0, x1) for every interval
        A.B.n_reads = n; A.B.seq = b->d_seq.as<u8>(); A.B.seq_off = b->d_off.as<i64>(); A.B.hash_id = b->d_ids.as<i64>();
        A.B.ovf = E.ovf.as<u32>(); A.B.rec = E.rec.as<ReadRec>();
        A.B.pool.intv = E.p_intv.as<Intv>(); A.B.pool.cap[POOL_INTV] = cap; A.B.pool.used = d_small + 8;
        A.n_work = n; A.work_ctr = d_small; A.ctrs = (DevCounters *)(d_small + 16);
        size_t stride = (seed_scratch_bytes(big) + 63) & ~(size_t)63;
        int grid = 8;
        E.spill_scratch.reserve(stride * grid * 128);
        A.scratch = E.spill_scratch.as<u8>(); A.scratch_stride = stride;
        if (seed2_usable(A) && !getenv("B200_DEBUG_SEED_V1")) {
            // the production path: k_seed2 for every read it takes, the reference-shaped kernel for the rest
            launch_seed2(E, A);
            E.list.reserve(n * 4 + 64);
            k_list_ovf<<<(unsigned)((n + 255) / 256), 256, 0, E.st>>>(E.ovf.as<u32>(), n, 0xffffffffu, E.list.as<i32>(), d_small + 1);
            unsigned long long n_sp = 0;
            CU_CHECK(cudaMemcpyAsync(&n_sp, d_small + 1, 8, cudaMemcpyDeviceToHost, E.st));
            CU_CHECK(cudaStreamSynchronize(E.st));
            if (n_sp) {
                k_clear_list<<<(unsigned)((n_sp + 255) / 256), 256, 0, E.st>>>(E.ovf.as<u32>(), E.list.as<i32>(), (i64)n_sp);
                KArgs S = A; S.order = E.list.as<i32>(); S.n_work = (i64)n_sp;
                launch_stage<0>(E, S, grid);
            }
        } else launch_stage<0>(E, A, grid);
        std::vector<ReadRec> rec(n);
        CU_CHECK(cudaMemcpyAsync(rec.data(), E.rec.p, n * sizeof(ReadRec), cudaMemcpyDeviceToHost, E.st));
        CU_CHECK(cudaStreamSynchronize(E.st));
        i64 tot = 0;
        for (i64 i = 0; i < n; ++i) tot += rec[i].n_intv;
        std::vector<Intv> pool(cap);
        CU_CHECK(cudaMemcpy(pool.data(), E.p_intv.p, cap * sizeof(Intv), cudaMemcpyDeviceToHost));
        int64_t *o = (int64_t *)malloc((n + 1) * 8);
        b200_intv_t *iv = (b200_intv_t *)malloc((tot + 1) * sizeof(b200_intv_t));
        i64 k = 0;
        for (i64 i = 0; i < n; ++i) {
            o[i] = k;
            for (int j = 0; j < rec[i].n_intv; ++j, ++k) {
                const Intv &p = pool[rec[i].intv_off + j];
                iv[k].x0 = p.x0; iv[k].x1 = p.x1; iv[k].x2 = p.x2; iv[k].info = p.info;
            }
        }
        o[n] = k;
        *intv_off = o; *intv = iv;
    } catch (const std::exception &e) { b200_batch_destroy(b); return fail(B200_ERR_CUDA, e.what()); }
    b200_batch_destroy(b);
    return B200_OK;
}

} // extern "C"
