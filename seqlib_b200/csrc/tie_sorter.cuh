// tie_sorter.cuh -- complete ordering of suffix-like items by successive 63-bit key chunks.
//
// Items are opaque u64 ids; KeyFn(id, depth) returns chunk `depth` of the item's (conceptually unbounded) sort key.
// One radix sort on chunk 0, then only the runs of equal chunks are refined with chunk 1, 2, ... (each round sorts the
// surviving ties by (run, chunk)).  Items whose keys are equal at every depth up to KeyFn::max_depth() are true
// duplicates and keep an arbitrary relative order.  Used by the FM-index builder (index_build.cu: 29-base chunks of one
// text) and the FMD-index builder (fmd.cu: 27-symbol chunks of read rotations).
#pragma once
#include <cub/cub.cuh>
#include "engine.cuh"

namespace b200 {

// flags[i] = 1 when element i starts a new run of equal (group, key)
static __global__ void k_head_flags(const u64 *__restrict__ keys, const u32 *__restrict__ grp, u64 n, u8 *__restrict__ head)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    head[i] = i == 0 || keys[i] != keys[i - 1] || (grp && grp[i] != grp[i - 1]);
}

// tie[i] = 1 when element i belongs to a run of length > 1
static __global__ void k_tie_flags(const u8 *__restrict__ head, u64 n, u8 *__restrict__ tie)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool next_head = i + 1 >= n || head[i + 1];
    tie[i] = !(head[i] && next_head);
}

// for the compacted tie subset: the key chunk at depth d
template <class KeyFn>
static __global__ void k_tie_keys(KeyFn kf, const u64 *__restrict__ pos_sub, u64 n, int depth, u64 *__restrict__ keys)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    keys[i] = kf(pos_sub[i], depth);
}

// group id of each tie element = index (within the subset) of its run head; computed with a max-scan of head positions
struct MaxOp { __device__ __forceinline__ u32 operator()(u32 a, u32 b) const { return a > b ? a : b; } };
static __global__ void k_head_index(const u8 *__restrict__ head, u64 n, u32 *__restrict__ hidx)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    hidx[i] = head[i] ? (u32)i : 0u;
}

static __global__ void k_scatter_back(const u64 *__restrict__ slot, const u64 *__restrict__ pos_sorted, u64 n, u64 *__restrict__ bucket_pos)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bucket_pos[slot[i]] = pos_sorted[i];
}

static __global__ void k_iota(u64 *a, u64 n) { u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; if (i < n) a[i] = i; }
static __global__ void k_gather64(const u64 *__restrict__ src, const u64 *__restrict__ idx, u64 n, u64 *__restrict__ dst)
{ u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; if (i < n) dst[i] = src[idx[i]]; }
static __global__ void k_gather32(const u32 *__restrict__ src, const u64 *__restrict__ idx, u64 n, u32 *__restrict__ dst)
{ u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; if (i < n) dst[i] = src[idx[i]]; }

static __global__ void k_gather8(const u8 *__restrict__ src, const u64 *__restrict__ idx, u64 n, u8 *__restrict__ dst)
{ u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; if (i < n) dst[i] = src[idx[i]]; }

static inline unsigned nb(u64 n, int t) { return (unsigned)((n + t - 1) / t); }

// Sort the suffixes in pos[0..n) (keys = key(p,0) already computed) completely; result in pos (may swap buffers).
template <class KeyFn>
struct Sorter {
    DevBuf keys2, pos2, tmp, head, tie, sub_slot, sub_pos, sub_key, sub_grp, sub_idx, sub_idx2, sub_grp2, sub_key2, sub_pos2, nsel;
    KeyFn kf; int max_depth = 4000000;

    void sort_bucket(u64 *&keys, u64 *&pos, u64 n)
    {
        if (n == 0) return;
        keys2.reserve(n * 8); pos2.reserve(n * 8);
        cub::DoubleBuffer<u64> dk(keys, keys2.as<u64>()), dp(pos, pos2.as<u64>());
        size_t tb = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tb, dk, dp, (int)n, 0, 63);
        tmp.reserve(tb);
        CU_CHECK(cub::DeviceRadixSort::SortPairs(tmp.p, tb, dk, dp, (int)n, 0, 63));
        u64 *sk = dk.Current(), *sp = dp.Current();
        // refine runs of equal keys
        head.reserve(n); tie.reserve(n); nsel.reserve(16);
        k_head_flags<<<nb(n, 256), 256>>>(sk, nullptr, n, head.as<u8>());
        k_tie_flags<<<nb(n, 256), 256>>>(head.as<u8>(), n, tie.as<u8>());
        // subset = slots of tie elements
        sub_slot.reserve(n * 8);
        {
            cub::CountingInputIterator<u64> it(0);
            size_t t2 = 0;
            cub::DeviceSelect::Flagged(nullptr, t2, it, tie.as<u8>(), sub_slot.as<u64>(), nsel.as<u64>(), (int)n);
            tmp.reserve(t2);
            CU_CHECK(cub::DeviceSelect::Flagged(tmp.p, t2, it, tie.as<u8>(), sub_slot.as<u64>(), nsel.as<u64>(), (int)n));
        }
        u64 m = 0;
        CU_CHECK(cudaMemcpy(&m, nsel.p, 8, cudaMemcpyDeviceToHost));
        if (m) {
            // group id of every tie element: the slot of its run head (run heads are tie elements too)
            sub_grp.reserve(m * 4); sub_pos.reserve(m * 8); sub_key.reserve(m * 8);
            sub_idx.reserve(m * 8); sub_idx2.reserve(m * 8); sub_grp2.reserve(m * 4); sub_key2.reserve(m * 8); sub_pos2.reserve(m * 8);
            // head flags restricted to the subset, then an inclusive max-scan of head indices gives each element its head's subset index
            DevBuf sub_head; sub_head.reserve(m);
            gather_heads(sub_slot.as<u64>(), m, sub_head.as<u8>());
            k_head_index<<<nb(m, 256), 256>>>(sub_head.as<u8>(), m, sub_grp.as<u32>());
            {
                size_t t3 = 0;
                cub::DeviceScan::InclusiveScan(nullptr, t3, sub_grp.as<u32>(), sub_grp.as<u32>(), MaxOp(), (int)m);
                tmp.reserve(t3);
                CU_CHECK(cub::DeviceScan::InclusiveScan(tmp.p, t3, sub_grp.as<u32>(), sub_grp.as<u32>(), MaxOp(), (int)m));
            }
            k_gather64<<<nb(m, 256), 256>>>(sp, sub_slot.as<u64>(), m, sub_pos.as<u64>());
            int depth = 1;
            u64 cur = m;              // current subset size; arrays sub_slot / sub_pos / sub_grp hold it
            while (cur) {
                if (depth > max_depth) break;      // equal at every depth: true duplicates
                k_tie_keys<<<nb(cur, 256), 256>>>(kf, sub_pos.as<u64>(), cur, depth, sub_key.as<u64>());
                // order by (group, key): sort by key, then stable sort by group
                k_iota<<<nb(cur, 256), 256>>>(sub_idx.as<u64>(), cur);
                {
                    size_t t4 = 0;
                    cub::DeviceRadixSort::SortPairs(nullptr, t4, sub_key.as<u64>(), sub_key2.as<u64>(), sub_idx.as<u64>(), sub_idx2.as<u64>(), (int)cur, 0, 63);
                    tmp.reserve(t4);
                    CU_CHECK(cub::DeviceRadixSort::SortPairs(tmp.p, t4, sub_key.as<u64>(), sub_key2.as<u64>(), sub_idx.as<u64>(), sub_idx2.as<u64>(), (int)cur, 0, 63));
                }
                k_gather32<<<nb(cur, 256), 256>>>(sub_grp.as<u32>(), sub_idx2.as<u64>(), cur, sub_grp2.as<u32>());
                {
                    size_t t5 = 0;
                    cub::DeviceRadixSort::SortPairs(nullptr, t5, sub_grp2.as<u32>(), sub_grp.as<u32>(), sub_idx2.as<u64>(), sub_idx.as<u64>(), (int)cur, 0, 32);
                    tmp.reserve(t5);
                    CU_CHECK(cub::DeviceRadixSort::SortPairs(tmp.p, t5, sub_grp2.as<u32>(), sub_grp.as<u32>(), sub_idx2.as<u64>(), sub_idx.as<u64>(), (int)cur, 0, 32));
                }
                // sub_idx = permutation into (group, key) order; sub_grp = sorted groups
                k_gather64<<<nb(cur, 256), 256>>>(sub_pos.as<u64>(), sub_idx.as<u64>(), cur, sub_pos2.as<u64>());
                k_gather64<<<nb(cur, 256), 256>>>(sub_key.as<u64>(), sub_idx.as<u64>(), cur, sub_key2.as<u64>());
                // the j-th element in (group,key) order goes to the j-th slot (slots ascending, groups contiguous)
                k_scatter_back<<<nb(cur, 256), 256>>>(sub_slot.as<u64>(), sub_pos2.as<u64>(), cur, sp);
                // new runs inside the subset
                head.reserve(cur); tie.reserve(cur);
                k_head_flags<<<nb(cur, 256), 256>>>(sub_key2.as<u64>(), sub_grp.as<u32>(), cur, head.as<u8>());
                k_tie_flags<<<nb(cur, 256), 256>>>(head.as<u8>(), cur, tie.as<u8>());
                // new group ids = subset index of the new run head (then compacted below)
                k_head_index<<<nb(cur, 256), 256>>>(head.as<u8>(), cur, sub_grp2.as<u32>());
                {
                    size_t t6 = 0;
                    cub::DeviceScan::InclusiveScan(nullptr, t6, sub_grp2.as<u32>(), sub_grp2.as<u32>(), MaxOp(), (int)cur);
                    tmp.reserve(t6);
                    CU_CHECK(cub::DeviceScan::InclusiveScan(tmp.p, t6, sub_grp2.as<u32>(), sub_grp2.as<u32>(), MaxOp(), (int)cur));
                }
                // compact slot / pos / group of the still-tied elements
                u64 next = 0;
                {
                    size_t t7 = 0;
                    cub::DeviceSelect::Flagged(nullptr, t7, sub_slot.as<u64>(), tie.as<u8>(), sub_idx2.as<u64>(), nsel.as<u64>(), (int)cur);
                    tmp.reserve(t7);
                    CU_CHECK(cub::DeviceSelect::Flagged(tmp.p, t7, sub_slot.as<u64>(), tie.as<u8>(), sub_idx2.as<u64>(), nsel.as<u64>(), (int)cur));
                    CU_CHECK(cudaMemcpy(&next, nsel.p, 8, cudaMemcpyDeviceToHost));
                    if (next) {
                        CU_CHECK(cudaMemcpy(sub_slot.p, sub_idx2.p, next * 8, cudaMemcpyDeviceToDevice));
                        CU_CHECK(cub::DeviceSelect::Flagged(tmp.p, t7, sub_pos2.as<u64>(), tie.as<u8>(), sub_pos.as<u64>(), nsel.as<u64>(), (int)cur));
                        size_t t8 = 0;
                        cub::DeviceSelect::Flagged(nullptr, t8, sub_grp2.as<u32>(), tie.as<u8>(), sub_grp.as<u32>(), nsel.as<u64>(), (int)cur);
                        tmp.reserve(t8);
                        CU_CHECK(cub::DeviceSelect::Flagged(tmp.p, t8, sub_grp2.as<u32>(), tie.as<u8>(), sub_grp.as<u32>(), nsel.as<u64>(), (int)cur));
                    }
                }
                cur = next;
                ++depth;
            }
        }
        keys = sk; pos = sp;
    }

    void gather_heads(const u64 *slot, u64 m, u8 *out) { k_gather8<<<nb(m, 256), 256>>>(head.as<u8>(), slot, m, out); }
};

} // namespace b200
