// common.cuh -- shared types of the seed-and-extend kernels.
//
// Everything marked HD is written once and compiled twice: by nvcc for sm_100a
// (the product) and by the host compiler inside tests/hostsim (a CPU emulation
// of the per-read device functions used only to debug logic in CI boxes that
// have no GPU; it is not linked into libseqlib_b200.so).
#pragma once
#include <stdint.h>
#include <stddef.h>

#if defined(__CUDACC__)
#define HD __host__ __device__ __forceinline__
#define HDN __host__ __device__ __noinline__
#else
#define HD inline
#define HDN
#endif

namespace b200 {

typedef int64_t i64;
typedef uint64_t u64;
typedef uint32_t u32;
typedef int32_t i32;
typedef uint8_t u8;
typedef int8_t i8;
typedef uint16_t u16;
typedef int16_t i16;

HD int popc64(u64 x)
{
#if defined(__CUDA_ARCH__)
    return __popcll(x);
#else
    return __builtin_popcountll(x);
#endif
}

// ---------------------------------------------------------------------------
// Device image of the FM-index (DESIGN.md "data layout in HBM").
//   occ  : one 32-byte block per 64 BWT symbols:
//            u32 cnt[4]  = number of A/C/G/T in bwt[0, 64*blk)   (16 B)
//            u64 sym[2]  = the 64 symbols as two bit planes: bit j of sym[0] = low bit, bit j of sym[1] = high bit
//                          of symbol j (16 B); a rank needs three 64-bit popcounts under one mask
//          ($ is removed exactly as in bwa: ranks >= primary shift down by one.)
//   sa   : u64 suffix-array samples every 2^sa_shift ranks, sa[0] = (u64)-1
//   text : forward + reverse-complement reference, 2 bits/base, 32 bases per u64,
//          base i in bits 2(i&31).. of text[i>>5]
// ---------------------------------------------------------------------------
struct OccBlock {
    u32 cnt[4];
    u64 sym[2];
};

HD int occ_sym(const OccBlock &b, int j) { return (int)((b.sym[0] >> j) & 1) | (int)((b.sym[1] >> j) & 1) << 1; }
HD void occ_set_sym(OccBlock &b, int j, u64 s) { b.sym[0] |= (s & 1) << j; b.sym[1] |= (s >> 1) << j; }

#define B200_MAX_CONTIGS_CONST 0

struct DevIndex {
    u64 primary;
    u64 L2[5];
    u64 seq_len;     // 2 * l_pac
    i64 l_pac;
    const OccBlock *occ;
    u64 n_occ;
    const u64 *sa;
    u64 n_sa;
    int sa_shift;
    const u64 *text;
    int n_seqs;
    const i64 *contig_off;   // n_seqs + 1 entries (last = l_pac)
    const i32 *contig_alt;   // n_seqs entries
};

// Options: the subset of mem_opt_t (bwa/bwamem.h:52-84) the path reads.
struct Opt {
    int a, b, o_del, e_del, o_ins, e_ins, pen_clip5, pen_clip3, w, zdrop;
    u64 max_mem_intv;
    int T, flag, min_seed_len, min_chain_weight, max_chain_extend;
    float split_factor;
    int split_width, max_occ, max_chain_gap;
    float mask_level, drop_ratio, mask_level_redun, mapQ_coef_len;
    int mapQ_coef_fac;
    i8 mat[25];
};

struct Intv {        // bwtintv_t (bwa/bwt.h:62-64)
    u64 x0, x1, x2, info;
};

struct Seed {        // mem_seed_t (bwa/bwamem.c:194-198)
    i64 rbeg;
    i32 qbeg, len, score;
    i32 next;        // link inside the per-read seed pool while chaining
};

struct Chain {       // mem_chain_t (bwa/bwamem.c:200-207)
    i64 pos;
    i32 n, first, rid, w;
    i32 kept, is_alt;
    i32 head, tail;  // seed list in the pool (chaining); head = offset of a contiguous run afterwards
};

struct Reg {         // mem_alnreg_t (bwa/bwamem.h:86-105)
    i64 rb, re;
    i32 qb, qe, rid, score, truesc, sub, alt_sc, csub, sub_n, w, seedcov;
    i32 secondary, secondary_all, seedlen0, n_comp, is_alt;
    float frac_rep;
    u32 pad_;
    u64 hash;
};

// per-stage work counters (atomically accumulated once per thread at exit)
struct Counters {
    unsigned long long occ_blocks, sa_reads, ref_bytes, sw_cells, n_ext, n_global, n_overflow;
};

// Error/overflow bits per read
enum {
    OVF_INTV = 1, OVF_SEED = 2, OVF_CHAIN = 4, OVF_REG = 8, OVF_OUT = 16, OVF_SCRATCH = 32
};

template <typename T> HD void swap_(T &a, T &b) { T t = a; a = b; b = t; }
template <typename T> HD T min_(T a, T b) { return a < b ? a : b; }
template <typename T> HD T max_(T a, T b) { return a > b ? a : b; }

} // namespace b200
