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
template <class KeyFn, class IdT>
static __global__ void k_tie_keys(KeyFn kf, const IdT *__restrict__ pos_sub, u64 n, int depth, u64 *__restrict__ keys)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    keys[i] = kf((u64)pos_sub[i], depth);
}

// group id of each tie element = index (within the subset) of its run head; computed with a max-scan of head positions
struct MaxOp { __device__ __forceinline__ u32 operator()(u32 a, u32 b) const { return a > b ? a : b; } };
static __global__ void k_head_index(const u8 *__restrict__ head, u64 n, u32 *__restrict__ hidx)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    hidx[i] = head[i] ? (u32)i : 0u;
}

// dst[slot[i]] = v[i]
template <class IdT>
static __global__ void k_scatter_back(const u32 *__restrict__ slot, const IdT *__restrict__ v, u64 n, IdT *__restrict__ dst)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    dst[slot[i]] = v[i];
}

static __global__ void k_iota32(u32 *a, u64 n) { u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; if (i < n) a[i] = (u32)i; }
template <class T>
static __global__ void k_gather_by32(const T *__restrict__ src, const u32 *__restrict__ idx, u64 n, T *__restrict__ dst)
{ u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; if (i < n) dst[i] = src[idx[i]]; }

static inline unsigned nb(u64 n, int t) { return (unsigned)((n + t - 1) / t); }

// Sort the items in pos[0..n) (keys = chunk 0 already computed) completely; result in pos (may swap buffers).
// n < 2^31: slots and permutation indices inside a bucket are 32-bit; IdT is the caller's item id (u64 text positions for
// the reference index, u32 row numbers for the read index: the radix passes move 12 instead of 16 bytes per item).
template <class KeyFn, class IdT = u64>
struct Sorter {
    DevBuf keys2, pos2, tmp, head, tie, sub_slot, sub_pos, sub_key, sub_grp, sub_idx, sub_idx2, sub_grp2, sub_key2, sub_pos2, nsel, sub_head, sub_slot2;
    KeyFn kf; int max_depth = 4000000;

    template <class K, class V>
    void sort_pairs(K *k_in, K *k_out, V *v_in, V *v_out, u64 n, int bits)
    {
        size_t tb = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tb, k_in, k_out, v_in, v_out, (int)n, 0, bits);
        tmp.reserve(tb);
        CU_CHECK(cub::DeviceRadixSort::SortPairs(tmp.p, tb, k_in, k_out, v_in, v_out, (int)n, 0, bits));
    }
    template <class T>
    u64 select_flagged(const T *in, const u8 *flags, T *out, u64 n)
    {
        size_t tb = 0;
        cub::DeviceSelect::Flagged(nullptr, tb, in, flags, out, nsel.as<u64>(), (int)n);
        tmp.reserve(tb);
        CU_CHECK(cub::DeviceSelect::Flagged(tmp.p, tb, in, flags, out, nsel.as<u64>(), (int)n));
        u64 m = 0;
        CU_CHECK(cudaMemcpy(&m, nsel.p, 8, cudaMemcpyDeviceToHost));
        return m;
    }
    void max_scan(u32 *a, u64 n)
    {
        size_t tb = 0;
        cub::DeviceScan::InclusiveScan(nullptr, tb, a, a, MaxOp(), (int)n);
        tmp.reserve(tb);
        CU_CHECK(cub::DeviceScan::InclusiveScan(tmp.p, tb, a, a, MaxOp(), (int)n));
    }

    void sort_bucket(u64 *&keys, IdT *&pos, u64 n)
    {
        if (n == 0) return;
        if (n >= (1ull << 31)) throw std::length_error("tie sorter: more than 2^31 items in one bucket");
        keys2.reserve(n * 8); pos2.reserve(n * sizeof(IdT));
        cub::DoubleBuffer<u64> dk(keys, keys2.as<u64>());
        cub::DoubleBuffer<IdT> dp(pos, pos2.as<IdT>());
        {
            size_t tb = 0;
            cub::DeviceRadixSort::SortPairs(nullptr, tb, dk, dp, (int)n, 0, 63);
            tmp.reserve(tb);
            CU_CHECK(cub::DeviceRadixSort::SortPairs(tmp.p, tb, dk, dp, (int)n, 0, 63));
        }
        u64 *sk = dk.Current(); IdT *sp = dp.Current();
        // refine runs of equal keys
        head.reserve(n); tie.reserve(n); nsel.reserve(16);
        k_head_flags<<<nb(n, 256), 256>>>(sk, nullptr, n, head.as<u8>());
        k_tie_flags<<<nb(n, 256), 256>>>(head.as<u8>(), n, tie.as<u8>());
        // subset = slots of tie elements
        sub_slot.reserve(n * 4); sub_idx.reserve(n * 4);
        k_iota32<<<nb(n, 256), 256>>>(sub_idx.as<u32>(), n);
        u64 m = select_flagged(sub_idx.as<u32>(), tie.as<u8>(), sub_slot.as<u32>(), n);
        if (m) {
            // group id of every tie element: the subset index of its run head (run heads are tie elements too)
            sub_grp.reserve(m * 4); sub_pos.reserve(m * sizeof(IdT)); sub_key.reserve(m * 8); sub_slot2.reserve(m * 4);
            sub_idx2.reserve(m * 4); sub_grp2.reserve(m * 4); sub_key2.reserve(m * 8); sub_pos2.reserve(m * sizeof(IdT));
            sub_head.reserve(m);
            k_gather_by32<u8><<<nb(m, 256), 256>>>(head.as<u8>(), sub_slot.as<u32>(), m, sub_head.as<u8>());
            k_head_index<<<nb(m, 256), 256>>>(sub_head.as<u8>(), m, sub_grp.as<u32>());
            max_scan(sub_grp.as<u32>(), m);
            k_gather_by32<IdT><<<nb(m, 256), 256>>>(sp, sub_slot.as<u32>(), m, sub_pos.as<IdT>());
            int depth = 1;
            u64 cur = m;              // current subset size; arrays sub_slot / sub_pos / sub_grp hold it
            u64 grp_range = m;        // group ids are head indices in the numbering of the PREVIOUS subset
            while (cur) {
                if (depth > max_depth) break;      // equal at every depth: true duplicates
                k_tie_keys<<<nb(cur, 256), 256>>>(kf, sub_pos.as<IdT>(), cur, depth, sub_key.as<u64>());
                // order by (group, key): sort by key, then stable sort by group
                k_iota32<<<nb(cur, 256), 256>>>(sub_idx.as<u32>(), cur);
                sort_pairs(sub_key.as<u64>(), sub_key2.as<u64>(), sub_idx.as<u32>(), sub_idx2.as<u32>(), cur, 63);
                k_gather_by32<u32><<<nb(cur, 256), 256>>>(sub_grp.as<u32>(), sub_idx2.as<u32>(), cur, sub_grp2.as<u32>());
                int gbits = 1; while ((1ull << gbits) < grp_range) ++gbits;
                sort_pairs(sub_grp2.as<u32>(), sub_grp.as<u32>(), sub_idx2.as<u32>(), sub_idx.as<u32>(), cur, gbits);
                // sub_idx = permutation into (group, key) order; sub_grp = sorted groups
                k_gather_by32<IdT><<<nb(cur, 256), 256>>>(sub_pos.as<IdT>(), sub_idx.as<u32>(), cur, sub_pos2.as<IdT>());
                k_gather_by32<u64><<<nb(cur, 256), 256>>>(sub_key.as<u64>(), sub_idx.as<u32>(), cur, sub_key2.as<u64>());
                // the j-th element in (group,key) order goes to the j-th slot (slots ascending, groups contiguous)
                k_scatter_back<IdT><<<nb(cur, 256), 256>>>(sub_slot.as<u32>(), sub_pos2.as<IdT>(), cur, sp);
                // new runs inside the subset
                head.reserve(cur); tie.reserve(cur);
                k_head_flags<<<nb(cur, 256), 256>>>(sub_key2.as<u64>(), sub_grp.as<u32>(), cur, head.as<u8>());
                k_tie_flags<<<nb(cur, 256), 256>>>(head.as<u8>(), cur, tie.as<u8>());
                // new group ids = subset index of the new run head (then compacted below)
                k_head_index<<<nb(cur, 256), 256>>>(head.as<u8>(), cur, sub_grp2.as<u32>());
                max_scan(sub_grp2.as<u32>(), cur);
                // compact slot / pos / group of the still-tied elements
                u64 next = select_flagged(sub_slot.as<u32>(), tie.as<u8>(), sub_slot2.as<u32>(), cur);
                if (next) {
                    std::swap(sub_slot.p, sub_slot2.p); std::swap(sub_slot.cap, sub_slot2.cap);
                    select_flagged(sub_pos2.as<IdT>(), tie.as<u8>(), sub_pos.as<IdT>(), cur);
                    select_flagged(sub_grp2.as<u32>(), tie.as<u8>(), sub_grp.as<u32>(), cur);
                }
                grp_range = cur;
                cur = next;
                ++depth;
            }
        }
        keys = sk; pos = sp;
    }
};

} // namespace b200
