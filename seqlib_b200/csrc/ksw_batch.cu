// ksw_batch.cu -- batched ksw_extend2 (bwa/ksw.c:416-515): config-3 microbenchmark and unit parity entry point.
#include <vector>
#include <cstring>
#include "engine.cuh"
#include "extend_group.cuh"

using namespace b200;

namespace b200 {

struct ExtArgs {
    const b200_ext_job_t *jobs; i64 n; const u8 *qp, *tp; i8 mat[25]; int o_del, e_del, o_ins, e_ins;
    b200_ext_out_t *out; EH *eh; int eh_stride; unsigned long long *cells; unsigned long long *work;
    u8 *redo;          // wavefront kernel: jobs it hands back (1); group kernel: when set, only those jobs are run
    int a, b;          // match / mismatch score when the matrix is bwa_fill_scmat(a, b), else a == 0
};

struct CellCtr { unsigned long long sw_cells, n_ext; };

// v1: one thread per job, persistent threads, the scalar recurrence of ksw.cuh
__global__ void __launch_bounds__(128) k_ext_scalar(const __grid_constant__ ExtArgs A)
{
    EH *eh = A.eh + (size_t)(blockIdx.x * blockDim.x + threadIdx.x) * A.eh_stride;
    CellCtr c; c.sw_cells = 0; c.n_ext = 0;
    for (;;) {
        unsigned long long base = 0;
        if ((threadIdx.x & 31) == 0) base = atomicAdd(A.work, 32ull);
        base = __shfl_sync(0xffffffffu, base, 0);
        if ((i64)base >= A.n) break;
        i64 i = (i64)base + (threadIdx.x & 31);
        if (i < A.n) {
            const b200_ext_job_t j = A.jobs[i];
            BytesSeq q; q.p = A.qp + j.q_off; q.step = 1;
            BytesSeq t; t.p = A.tp + j.t_off; t.step = 1;
            ExtResult r = extend2(j.qlen, q, j.tlen, t, A.mat, A.o_del, A.e_del, A.o_ins, A.e_ins, j.w, j.end_bonus, j.zdrop, j.h0, eh, c);
            b200_ext_out_t o; o.score = r.score; o.qle = r.qle; o.tle = r.tle; o.gtle = r.gtle; o.gscore = r.gscore; o.max_off = r.max_off;
            A.out[i] = o;
        }
    }
    unsigned long long x = c.sw_cells;
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(A.cells, x);
}


// v2: G lanes per job (extend2_group), per-group state in shared memory
template <int G>
__global__ void __launch_bounds__(128) k_ext_group(const __grid_constant__ ExtArgs A, int maxq)
{
    extern __shared__ __align__(16) u8 smem_raw[];
    __shared__ i8 smat[32];
    if (threadIdx.x < 25) smat[threadIdx.x] = A.mat[threadIdx.x];
    __syncthreads();
    const int lane = threadIdx.x & 31, gib = threadIdx.x / G;
    GroupCtx<G> g; g.gl = threadIdx.x % G;
    g.mask = G == 32 ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(G - 1)));
    u8 *smem = smem_raw + (size_t)gib * group_smem_bytes(maxq);
    int *H = (int *)smem, *E = H + (maxq + 2);
    u8 *q = (u8 *)(E + (maxq + 2));
    CellCtr c; c.sw_cells = 0; c.n_ext = 0;
    for (;;) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(A.work, (unsigned long long)(32 / G));
        base = __shfl_sync(0xffffffffu, base, 0);
        if ((i64)base >= A.n) break;
        i64 i = (i64)base + lane / G;
        if (i < A.n && (!A.redo || A.redo[i])) {
            const b200_ext_job_t j = A.jobs[i];
            for (int k = g.gl; k < j.qlen; k += G) q[k] = A.qp[j.q_off + k];
            g.sync();
            BytesSeq qs; qs.p = q; qs.step = 1;
            BytesSeq ts; ts.p = A.tp + j.t_off; ts.step = 1;
            ExtResult r = extend2_group(g, j.qlen, qs, j.tlen, ts, smat, A.o_del, A.e_del, A.o_ins, A.e_ins, j.w, j.end_bonus, j.zdrop, j.h0, H, E, c);
            if (g.gl == 0) { b200_ext_out_t o; o.score = r.score; o.qle = r.qle; o.tle = r.tle; o.gtle = r.gtle; o.gscore = r.gscore; o.max_off = r.max_off; A.out[i] = o; }
        }
        __syncwarp();
    }
    unsigned long long x = c.sw_cells;
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) atomicAdd(A.cells, x);
}

// v3: the packed 16-bit anti-diagonal wavefront (ksw_wave.cuh), G lanes per job; jobs it cannot finish exactly are flagged in A.redo
template <int G>
__global__ void __launch_bounds__(128, 6) k_ext_wave(const __grid_constant__ ExtArgs A, int maxq)
{
    extern __shared__ __align__(16) u8 smem_raw[];
    const int lane = threadIdx.x & 31, gib = threadIdx.x / G;
    GroupCtx<G> g; g.gl = threadIdx.x % G;
    g.mask = G == 32 ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(G - 1)));
    u32 *ehs = (u32 *)(smem_raw + (size_t)gib * (((size_t)WAVE_WORDS(maxq, G) * 4 + 15) & ~(size_t)15));
    CellCtr c; c.sw_cells = 0; c.n_ext = 0;
    for (;;) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(A.work, (unsigned long long)(32 / G));
        base = __shfl_sync(0xffffffffu, base, 0);
        if ((i64)base >= A.n) break;
        i64 i = (i64)base + lane / G;
        if (i < A.n) {
            const b200_ext_job_t j = A.jobs[i];
            BytesSeq qs; qs.p = A.qp + j.q_off; qs.step = 1;
            BytesSeq ts; ts.p = A.tp + j.t_off; ts.step = 1;
            ExtResult r;
            bool ok = A.a > 0 && wave_eligible(j.qlen, j.tlen, j.h0, A.a, j.end_bonus);
            if (ok) ok = extend2_wave<G>(g, j.qlen, qs, j.tlen, ts, A.a, A.b, A.o_del, A.e_del, A.o_ins, A.e_ins, j.w, j.end_bonus, j.zdrop, j.h0, ehs, r, c);
            if (g.gl == 0) {
                A.redo[i] = ok ? 0 : 1;
                if (ok) { b200_ext_out_t o; o.score = r.score; o.qle = r.qle; o.tle = r.tle; o.gtle = r.gtle; o.gscore = r.gscore; o.max_off = r.max_off; A.out[i] = o; }
            }
        }
        __syncwarp();
    }
    unsigned long long x = c.sw_cells;
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) atomicAdd(A.cells, x);
}

template <int G>
static void launch_ext_wave(const ExtArgs &A, int maxq, int sms)
{
    size_t smem = (size_t)(128 / G) * (((size_t)WAVE_WORDS(maxq, G) * 4 + 15) & ~(size_t)15);
    CU_CHECK(cudaFuncSetAttribute(k_ext_wave<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per = 1;
    CU_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_ext_wave<G>, 128, smem));
    k_ext_wave<G><<<sms * (per < 1 ? 1 : per), 128, smem>>>(A, maxq);
}

template <int G>
static void launch_ext_group(const ExtArgs &A, int maxq, int sms)
{
    size_t smem = (size_t)(128 / G) * group_smem_bytes(maxq);
    CU_CHECK(cudaFuncSetAttribute(k_ext_group<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per = 1;
    CU_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_ext_group<G>, 128, smem));
    k_ext_group<G><<<sms * (per < 1 ? 1 : per), 128, smem>>>(A, maxq);
}

} // namespace b200

extern "C" int b200_ksw_extend2_batch(int64_t n, const b200_ext_job_t *jobs, const uint8_t *qpool, int64_t qpool_len,
                                      const uint8_t *tpool, int64_t tpool_len, const int8_t mat[25], int o_del, int e_del, int o_ins, int e_ins,
                                      b200_ext_out_t *out, uint64_t *cells, float *kernel_ms)
{
    if (n < 0 || (n && (!jobs || !qpool || !tpool || !out))) return fail(B200_ERR_ARG, "bad argument");
    try {
        int dev = 0, sms = 148;
        CU_CHECK(cudaGetDevice(&dev));
        CU_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        int maxq = 1;
        for (i64 i = 0; i < n; ++i) { if (jobs[i].qlen > maxq) maxq = jobs[i].qlen; if (jobs[i].h0 <= 0) return fail(B200_ERR_ARG, "h0 must be positive"); }
        DevBuf dj, dq, dt, dout, deh, dsmall;
        dj.reserve(n * sizeof(b200_ext_job_t) + 64); dq.reserve(qpool_len + 64); dt.reserve(tpool_len + 64); dout.reserve(n * sizeof(b200_ext_out_t) + 64); dsmall.reserve(64);
        CU_CHECK(cudaMemcpy(dj.p, jobs, n * sizeof(b200_ext_job_t), cudaMemcpyHostToDevice));
        CU_CHECK(cudaMemcpy(dq.p, qpool, qpool_len, cudaMemcpyHostToDevice));
        CU_CHECK(cudaMemcpy(dt.p, tpool, tpool_len, cudaMemcpyHostToDevice));
        CU_CHECK(cudaMemset(dsmall.p, 0, 64));
        int per = 1;
        CU_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_ext_scalar, 128, 0));
        int grid = sms * (per < 1 ? 1 : per);
        int stride = maxq + 2;
        deh.reserve((size_t)grid * 128 * stride * sizeof(EH));
        ExtArgs A; memset(&A, 0, sizeof(A));
        A.jobs = dj.as<b200_ext_job_t>(); A.n = n; A.qp = dq.as<u8>(); A.tp = dt.as<u8>(); memcpy(A.mat, mat, 25);
        A.o_del = o_del; A.e_del = e_del; A.o_ins = o_ins; A.e_ins = e_ins; A.out = dout.as<b200_ext_out_t>();
        A.eh = deh.as<EH>(); A.eh_stride = stride; A.cells = dsmall.as<unsigned long long>(); A.work = dsmall.as<unsigned long long>() + 1;
        cudaEvent_t e0, e1; CU_CHECK(cudaEventCreate(&e0)); CU_CHECK(cudaEventCreate(&e1));
        CU_CHECK(cudaEventRecord(e0));
        int G = getenv("B200_KSW_G") ? atoi(getenv("B200_KSW_G")) : 16;
        size_t need = (size_t)(128 / (G > 0 ? G : 1)) * group_smem_bytes(maxq);
        // default: the wavefront kernel first, then the row-synchronous group kernel over what it handed back
        // (B200_KSW_WAVE=0: group kernel only -- its cell count is the reference's band-trimmed count)
        const int WG = getenv("B200_KSW_WAVE") ? atoi(getenv("B200_KSW_WAVE")) : 4;
        bool fill = mat[0] > 0 && mat[1] <= -1;
        for (int x = 0; x < 4 && fill; ++x) for (int y = 0; y < 4; ++y) if (mat[x * 5 + y] != (x == y ? mat[0] : mat[1])) fill = false;
        if (WG > 0 && fill && maxq <= WAVE_MAXQ && G != 0 && need <= 200 * 1024 && -mat[1] + mat[0] < 256) {
            DevBuf dredo; dredo.reserve(n + 64);
            A.redo = dredo.as<u8>(); A.a = mat[0]; A.b = -mat[1];
            if (WG == 8) launch_ext_wave<8>(A, maxq, sms); else if (WG == 2) launch_ext_wave<2>(A, maxq, sms); else launch_ext_wave<4>(A, maxq, sms);
            CU_CHECK(cudaMemsetAsync(A.work, 0, 8));
            launch_ext_group<16>(A, maxq, sms);
            CU_CHECK(cudaDeviceSynchronize());
        }
        else if (G == 0 || need > 200 * 1024) k_ext_scalar<<<grid, 128>>>(A);
        else if (G == 4) launch_ext_group<4>(A, maxq, sms);
        else if (G == 8) launch_ext_group<8>(A, maxq, sms);
        else if (G == 32) launch_ext_group<32>(A, maxq, sms);
        else launch_ext_group<16>(A, maxq, sms);
        CU_CHECK(cudaEventRecord(e1));
        CU_CHECK(cudaEventSynchronize(e1));
        CU_CHECK(cudaGetLastError());
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        CU_CHECK(cudaMemcpy(out, dout.p, n * sizeof(b200_ext_out_t), cudaMemcpyDeviceToHost));
        unsigned long long hc = 0;
        CU_CHECK(cudaMemcpy(&hc, dsmall.p, 8, cudaMemcpyDeviceToHost));
        if (cells) *cells = hc;
        if (kernel_ms) *kernel_ms = ms;
    } catch (const std::exception &e) { return fail(B200_ERR_CUDA, e.what()); }
    return B200_OK;
}
