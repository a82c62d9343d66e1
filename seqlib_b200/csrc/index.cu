// index.cu -- FM-index handle: device image layout, bwa on-disk format I/O,
// conversions between bwa's host layout and the 32-byte Occ block image,
// and the index part of the C ABI.
//   b200_index_construct  <- BWAIndex::ConstructIndex (src/BWAIndex.cpp:83-180)
//   b200_index_load       <- BWAIndex::LoadIndex -> bwa_idx_load (bwa/bwa.c:289-316),
//                            bwt_restore_bwt/sa (bwa/bwt.c:421-462), bns_restore (bwa/bntseq.c:97-209)
//   b200_index_write      <- BWAIndex::WriteIndex (src/BWAIndex.cpp:360-406), bwt_dump_* (bwa/bwt.c:385-407), bns_dump (bwa/bntseq.c:65-95)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <sys/stat.h>
#include "engine.cuh"

using namespace b200;

namespace b200 {

static thread_local std::string g_err;
void set_error(const std::string &msg) { g_err = msg; }
int fail(int code, const std::string &msg) { g_err = msg; return code; }

static u64 align_up(u64 x, u64 a) { return (x + a - 1) / a * a; }

BlobHeader plan_blob(u64 seq_len, i64 l_pac, int sa_shift, const std::vector<ContigMeta> &contigs)
{
    BlobHeader h; memset(&h, 0, sizeof(h));
    h.magic = BLOB_MAGIC; h.seq_len = seq_len; h.l_pac = l_pac; h.sa_shift = sa_shift; h.n_seqs = (i32)contigs.size();
    h.n_occ = (seq_len + 63) / 64 + 1;
    h.n_sa = (seq_len >> sa_shift) + 1;
    h.n_text = (seq_len + 31) / 32 + 1;
    u64 off = align_up(sizeof(BlobHeader), 256);
    h.off_occ = off; off = align_up(off + h.n_occ * sizeof(OccBlock), 256);
    h.off_sa = off; off = align_up(off + h.n_sa * 8, 256);
    h.off_text = off; off = align_up(off + h.n_text * 8, 256);
    h.off_coff = off; off = align_up(off + ((u64)h.n_seqs + 1) * 8, 256);
    h.off_calt = off; off = align_up(off + (u64)h.n_seqs * 4 + 4, 256);
    u64 nb = 0;
    for (auto &c : contigs) nb += c.name.size() + 1 + c.anno.size() + 1;
    h.off_names = off; h.names_bytes = nb; off = align_up(off + nb + 8, 256);
    h.total_bytes = off;
    return h;
}

void bind_blob(b200_index *idx, void *d_blob, const BlobHeader &h)
{
    u8 *b = (u8 *)d_blob;
    DevIndex &d = idx->dev;
    d.primary = h.primary; for (int i = 0; i < 5; ++i) d.L2[i] = h.L2[i];
    d.seq_len = h.seq_len; d.l_pac = h.l_pac;
    d.occ = (const OccBlock *)(b + h.off_occ); d.n_occ = h.n_occ;
    d.sa = (const u64 *)(b + h.off_sa); d.n_sa = h.n_sa; d.sa_shift = h.sa_shift;
    d.text = (const u64 *)(b + h.off_text);
    d.n_seqs = h.n_seqs;
    d.contig_off = (const i64 *)(b + h.off_coff);
    d.contig_alt = (const i32 *)(b + h.off_calt);
    idx->d_blob = d_blob; idx->blob_bytes = (i64)h.total_bytes;
    idx->primary = h.primary; for (int i = 0; i < 5; ++i) idx->L2[i] = h.L2[i];
    idx->seq_len = h.seq_len; idx->l_pac = h.l_pac;
}

void upload_blob_meta(void *d_blob, const BlobHeader &h, const std::vector<ContigMeta> &contigs, cudaStream_t st)
{
    u8 *b = (u8 *)d_blob;
    std::vector<i64> coff(contigs.size() + 1);
    std::vector<i32> calt(contigs.size() + 1, 0);
    std::string names;
    for (size_t i = 0; i < contigs.size(); ++i) {
        coff[i] = contigs[i].offset; calt[i] = contigs[i].is_alt;
        names += contigs[i].name; names.push_back('\0'); names += contigs[i].anno; names.push_back('\0');
    }
    coff[contigs.size()] = h.l_pac;
    CU_CHECK(cudaMemcpyAsync(b, &h, sizeof(h), cudaMemcpyHostToDevice, st));
    CU_CHECK(cudaMemcpyAsync(b + h.off_coff, coff.data(), coff.size() * 8, cudaMemcpyHostToDevice, st));
    CU_CHECK(cudaMemcpyAsync(b + h.off_calt, calt.data(), contigs.size() * 4, cudaMemcpyHostToDevice, st));
    if (!names.empty()) CU_CHECK(cudaMemcpyAsync(b + h.off_names, names.data(), names.size(), cudaMemcpyHostToDevice, st));
    CU_CHECK(cudaStreamSynchronize(st));
}

// SA sampling density of the device image: as dense as fits comfortably.  A
// denser sample changes no result (bwt_sa, bwa/bwt.c:86-96, returns the same
// value) but removes the LF walk, i.e. ~31 dependent Occ-block reads per hit.
int pick_sa_shift(u64 seq_len, int max_shift)
{
    const char *e = getenv("B200_SA_SHIFT");
    if (e) { int v = atoi(e); if (v < 0) v = 0; if (v > max_shift) v = max_shift; return v; }
    size_t fr = 0, tot = 0;
    if (cudaMemGetInfo(&fr, &tot) != cudaSuccess) return max_shift;
    fr += dev_pool().pooled;          // pooled blocks are reclaimable
    for (int s = 0; s <= max_shift; ++s) {
        u64 bytes = ((seq_len >> s) + 1) * 8;
        if (bytes <= fr / 3) return s;
    }
    return max_shift;
}

// ---- conversion kernels -------------------------------------------------

// bwa Occ-interleaved bwt (bwa/bwtindex.c:149-171) -> 32-byte blocks; one thread per 64-symbol block
__global__ void k_occ_from_bwa(const u32 *__restrict__ bwt, u64 seq_len, OccBlock *__restrict__ occ, u64 n_occ)
{
    u64 b = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_occ) return;
    u64 x0 = b * 64;
    OccBlock o; o.sym[0] = o.sym[1] = 0; o.cnt[0] = o.cnt[1] = o.cnt[2] = o.cnt[3] = 0;
    if (x0 < seq_len) {
        const u32 *blk = bwt + ((x0 >> 7) << 4);
        const u64 *c64 = (const u64 *)blk;
        u64 c[4] = {c64[0], c64[1], c64[2], c64[3]};
        const u32 *w = blk + 8;
        if (x0 & 64) {                               // second half: add the first 64 symbols of the 128-block
            for (int j = 0; j < 4; ++j) {
                u32 v = w[j];
                for (int k = 0; k < 16; ++k) ++c[(v >> ((15 - k) << 1)) & 3];
            }
            w += 4;
        }
        for (int i = 0; i < 4; ++i) o.cnt[i] = (u32)c[i];
        for (int j = 0; j < 64; ++j) {
            if (x0 + j >= seq_len) break;
            u32 v = w[j >> 4];
            u64 s = (v >> ((15 - (j & 15)) << 1)) & 3;
            occ_set_sym(o, j, s);
        }
    } else {                                          // sentinel block past the end: totals
        // filled by the caller through L2 (kept zero symbols)
    }
    occ[b] = o;
}

__global__ void k_fix_tail_counts(OccBlock *occ, u64 n_occ, u64 seq_len, u64 c0, u64 c1, u64 c2, u64 c3)
{
    // blocks starting at or beyond seq_len carry the total counts
    u64 b = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_occ || b * 64 < seq_len) return;
    occ[b].cnt[0] = (u32)c0; occ[b].cnt[1] = (u32)c1; occ[b].cnt[2] = (u32)c2; occ[b].cnt[3] = (u32)c3;
}

// forward pac (4 bases/byte, base i in bits (~i&3)*2) -> forward + reverse-complement text, 32 bases per u64
__global__ void k_text_from_pac(const u8 *__restrict__ pac, i64 l_pac, u64 *__restrict__ text, u64 n_text)
{
    u64 wi = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (wi >= n_text) return;
    u64 N = (u64)l_pac * 2, v = 0;
    for (int j = 0; j < 32; ++j) {
        u64 p = wi * 32 + j;
        if (p >= N) break;
        u64 f = p < (u64)l_pac ? p : N - 1 - p;
        u64 c = (pac[f >> 2] >> ((~f & 3) << 1)) & 3;
        if (p >= (u64)l_pac) c = 3 - c;
        v |= c << (2 * j);
    }
    text[wi] = v;
}

// 32-byte blocks -> bwa Occ-interleaved bwt; one thread per 128-symbol block
__global__ void k_bwa_from_occ(const OccBlock *__restrict__ occ, u64 seq_len, u32 *__restrict__ bwt, u64 n_blk128)
{
    u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_blk128) return;
    u32 *blk = bwt + j * 16;
    const OccBlock &a = occ[2 * j];
    u64 *c64 = (u64 *)blk;
    for (int i = 0; i < 4; ++i) c64[i] = a.cnt[i];
    u64 x0 = j * 128;
    u64 rem = seq_len - x0 < 128 ? seq_len - x0 : 128;
    int nw = (int)((rem + 15) >> 4);
    for (int w = 0; w < nw; ++w) {
        u32 v = 0;
        for (int k = 0; k < 16; ++k) {
            u64 x = x0 + w * 16 + k;
            if (x >= seq_len) break;
            const OccBlock &b = occ[x >> 6];
            int jj = (int)(x & 63);
            u32 s = (u32)occ_sym(b, jj);
            v |= s << ((15 - k) << 1);
        }
        blk[8 + w] = v;
    }
}

struct NullCtr { unsigned long long occ_blocks, sa_reads; };

// denser SA samples from a sparse one by LF-walking (bwt_sa semantics)
__global__ void k_densify_sa(DevIndex src, u64 *__restrict__ dst, int dst_shift, u64 n_dst)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_dst) return;
    if (i == 0) { dst[0] = ~0ull; return; }
    NullCtr c; c.occ_blocks = 0; c.sa_reads = 0;
    dst[i] = sa_lookup(src, i << dst_shift, c);
}

__global__ void k_gather_sa(const u64 *__restrict__ src, int src_shift, u64 *__restrict__ dst, int dst_shift, u64 n_dst)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_dst) return;
    dst[i] = src[(i << dst_shift) >> src_shift];
}

static inline unsigned nblocks(u64 n, int t) { return (unsigned)((n + t - 1) / t); }

void image_from_bwa_arrays(b200_index *idx)
{
    u64 N = idx->seq_len;
    for (int c = 0; c < 4; ++c)
        if (idx->L2[c + 1] - idx->L2[c] >= (1ull << 32)) throw std::runtime_error("a single base occurs >= 2^32 times: outside the 32-bit Occ block layout");
    int src_shift = 0; while ((1 << src_shift) < idx->sa_intv) ++src_shift;
    int shift = pick_sa_shift(N, src_shift);
    BlobHeader h = plan_blob(N, idx->l_pac, shift, idx->contigs);
    h.primary = idx->primary; for (int i = 0; i < 5; ++i) h.L2[i] = idx->L2[i];
    void *blob = nullptr;
    if (cudaMalloc(&blob, h.total_bytes) != cudaSuccess) { cudaGetLastError(); dev_pool().trim(0); CU_CHECK(cudaMalloc(&blob, h.total_bytes)); }
    CU_CHECK(cudaMemset(blob, 0, h.total_bytes));
    idx->owns_blob = true;
    upload_blob_meta(blob, h, idx->contigs, 0);
    bind_blob(idx, blob, h);
    u8 *b = (u8 *)blob;
    DevBuf tmp;
    // occ
    tmp.reserve(idx->h_bwt.size() * 4);
    CU_CHECK(cudaMemcpy(tmp.p, idx->h_bwt.data(), idx->h_bwt.size() * 4, cudaMemcpyHostToDevice));
    k_occ_from_bwa<<<nblocks(h.n_occ, 256), 256>>>(tmp.as<u32>(), N, (OccBlock *)(b + h.off_occ), h.n_occ);
    k_fix_tail_counts<<<nblocks(h.n_occ, 256), 256>>>((OccBlock *)(b + h.off_occ), h.n_occ, N, idx->L2[1] - idx->L2[0], idx->L2[2] - idx->L2[1],
                                                      idx->L2[3] - idx->L2[2], idx->L2[4] - idx->L2[3]);
    CU_CHECK(cudaDeviceSynchronize());
    // text
    tmp.reserve(idx->h_pac.size() + 16);
    CU_CHECK(cudaMemcpy(tmp.p, idx->h_pac.data(), idx->h_pac.size(), cudaMemcpyHostToDevice));
    k_text_from_pac<<<nblocks(h.n_text, 256), 256>>>(tmp.as<u8>(), idx->l_pac, (u64 *)(b + h.off_text), h.n_text);
    CU_CHECK(cudaDeviceSynchronize());
    // sa
    if (shift == src_shift) {
        CU_CHECK(cudaMemcpy(b + h.off_sa, idx->h_sa.data(), idx->h_sa.size() * 8, cudaMemcpyHostToDevice));
    } else {
        tmp.reserve(idx->h_sa.size() * 8);
        CU_CHECK(cudaMemcpy(tmp.p, idx->h_sa.data(), idx->h_sa.size() * 8, cudaMemcpyHostToDevice));
        DevIndex src = idx->dev; src.sa = tmp.as<u64>(); src.sa_shift = src_shift; src.n_sa = idx->h_sa.size();
        k_densify_sa<<<nblocks(h.n_sa, 256), 256>>>(src, (u64 *)(b + h.off_sa), shift, h.n_sa);
        CU_CHECK(cudaDeviceSynchronize());
    }
    CU_CHECK(cudaGetLastError());
}

void host_copy_from_image(b200_index *idx)
{
    u64 N = idx->seq_len;
    u64 n128 = (N + 127) / 128;
    u64 bwt_size = ((N + 15) >> 4) + (n128 + 1) * 8;
    DevBuf tmp; tmp.reserve(bwt_size * 4 + 64);
    CU_CHECK(cudaMemset(tmp.p, 0, bwt_size * 4));
    k_bwa_from_occ<<<nblocks(n128, 128), 128>>>(idx->dev.occ, N, tmp.as<u32>(), n128);
    CU_CHECK(cudaDeviceSynchronize());
    idx->h_bwt.assign(bwt_size, 0);
    CU_CHECK(cudaMemcpy(idx->h_bwt.data(), tmp.p, bwt_size * 4, cudaMemcpyDeviceToHost));
    // the trailing totals (bwa/bwtindex.c:166) sit right after the last symbol word
    u64 k_end = bwt_size - 8;
    u64 tot[4] = {idx->L2[1] - idx->L2[0], idx->L2[2] - idx->L2[1], idx->L2[3] - idx->L2[2], idx->L2[4] - idx->L2[3]};
    memcpy(&idx->h_bwt[k_end], tot, 32);
    // SA at bwa's interval 32
    idx->sa_intv = 32;
    u64 n_sa = (N + 32) / 32;
    tmp.reserve(n_sa * 8);
    if (idx->dev.sa_shift > 5) throw std::runtime_error("device SA sparser than 32");
    k_gather_sa<<<nblocks(n_sa, 256), 256>>>(idx->dev.sa, idx->dev.sa_shift, tmp.as<u64>(), 5, n_sa);
    CU_CHECK(cudaDeviceSynchronize());
    idx->h_sa.assign(n_sa, 0);
    CU_CHECK(cudaMemcpy(idx->h_sa.data(), tmp.p, n_sa * 8, cudaMemcpyDeviceToHost));
    idx->h_sa[0] = ~0ull;
    idx->has_host = true;
}

} // namespace b200

b200_index::~b200_index() { if (owns_blob && d_blob) cudaFree(d_blob); if (d_seedtab) cudaFree(d_seedtab); }

// -------------------------------------------------------------------------
// C ABI
// -------------------------------------------------------------------------
extern "C" {

const char *b200_last_error(void) { return g_err.c_str(); }

int b200_device_count(void) { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) return 0; return n; }
int b200_set_device(int ordinal)
{
    cudaError_t e = cudaSetDevice(ordinal);
    if (e != cudaSuccess) return fail(B200_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
    // keep the local-memory reservation at its high-water mark: the overlap-record kernels have 1.5 KB stack frames, and
    // shrinking / regrowing the reservation between launches synchronises the device (concurrent windows would serialise)
    if (!getenv("B200_NO_LMEM_MAX")) { cudaSetDeviceFlags(cudaDeviceLmemResizeToMax); cudaGetLastError(); }
    return B200_OK;
}

void b200_fill_scmat(int a, int b, int8_t mat[25])      // bwa_fill_scmat (bwa/bwa.c:136-145)
{
    int k = 0;
    for (int i = 0; i < 4; ++i) {
        for (int j = 0; j < 4; ++j) mat[k++] = (int8_t)(i == j ? a : -b);
        mat[k++] = -1;
    }
    for (int j = 0; j < 5; ++j) mat[k++] = -1;
}

void b200_mem_opt_init(b200_mem_opt_t *o)               // mem_opt_init (bwa/bwamem.c:74-110)
{
    memset(o, 0, sizeof(*o));
    o->a = 1; o->b = 4; o->o_del = o->o_ins = 6; o->e_del = o->e_ins = 1; o->w = 100; o->T = 30; o->zdrop = 100;
    o->pen_unpaired = 17; o->pen_clip5 = o->pen_clip3 = 5; o->max_mem_intv = 20; o->min_seed_len = 19; o->split_width = 10;
    o->max_occ = 500; o->max_chain_gap = 10000; o->max_ins = 10000; o->mask_level = 0.50f; o->drop_ratio = 0.50f;
    o->XA_drop_ratio = 0.80f; o->split_factor = 1.5f; o->chunk_size = 10000000; o->n_threads = 1; o->max_XA_hits = 5;
    o->max_XA_hits_alt = 200; o->max_matesw = 50; o->mask_level_redun = 0.95f; o->min_chain_weight = 0;
    o->max_chain_extend = 1 << 30; o->mapQ_coef_len = 50; o->mapQ_coef_fac = 3;   /* (int)log(50) */
    b200_fill_scmat(o->a, o->b, o->mat);
}

static void fill_view_contigs(b200_index *idx)
{
    idx->view_contigs.resize(idx->contigs.size());
    for (size_t i = 0; i < idx->contigs.size(); ++i) {
        const ContigMeta &c = idx->contigs[i];
        b200_contig_t &v = idx->view_contigs[i];
        v.offset = c.offset; v.len = c.len; v.n_ambs = c.n_ambs; v.gi = c.gi; v.is_alt = c.is_alt;
        v.name = c.name.c_str(); v.anno = c.anno.c_str();
    }
}

static unsigned char nt4_of(unsigned char c)            // nst_nt4_table (bwa/bntseq.c:46-63)
{
    switch (c) {
    case 'A': case 'a': return 0; case 'C': case 'c': return 1;
    case 'G': case 'g': return 2; case 'T': case 't': return 3;
    case '-': return 5; default: return 4;
    }
}

static int construct_common(b200_index *idx, int flags)
{
    u64 N = (u64)idx->l_pac * 2;
    int shift = pick_sa_shift(N, 5);
    BlobHeader h = plan_blob(N, idx->l_pac, shift, idx->contigs);
    void *blob = nullptr;
    if (cudaMalloc(&blob, h.total_bytes) != cudaSuccess) { cudaGetLastError(); dev_pool().trim(0); CU_CHECK(cudaMalloc(&blob, h.total_bytes)); }
    CU_CHECK(cudaMemset(blob, 0, h.total_bytes));
    idx->owns_blob = true;
    idx->seq_len = N;
    bind_blob(idx, blob, h);
    {   // text from the forward pac
        DevBuf tmp; tmp.reserve(idx->h_pac.size() + 16);
        CU_CHECK(cudaMemcpy(tmp.p, idx->h_pac.data(), idx->h_pac.size(), cudaMemcpyHostToDevice));
        k_text_from_pac<<<nblocks(h.n_text, 256), 256>>>(tmp.as<u8>(), idx->l_pac, (u64 *)((u8 *)blob + h.off_text), h.n_text);
        CU_CHECK(cudaDeviceSynchronize());
    }
    build_fm_index_device(idx, h);                       // fills occ + sa, sets idx->primary / L2
    h.primary = idx->primary; for (int i = 0; i < 5; ++i) h.L2[i] = idx->L2[i];
    upload_blob_meta(blob, h, idx->contigs, 0);
    bind_blob(idx, blob, h);
    if (flags & 1) host_copy_from_image(idx);
    return B200_OK;
}

int b200_index_construct(int n, const char *const *names, const char *const *seqs, int flags, b200_index_t **out)
{
    if (!out) return fail(B200_ERR_ARG, "out is NULL");
    *out = nullptr;
    if (n <= 0) return fail(B200_ERR_ARG, "no reference sequences");
    for (int i = 0; i < n; ++i)
        if (!names[i] || !seqs[i] || !names[i][0] || !seqs[i][0]) return fail(B200_ERR_ARG, "each reference must have a non-empty name and sequence");
    b200_index *idx = new b200_index;
    try {
        i64 tot = 0;
        for (int i = 0; i < n; ++i) {
            ContigMeta c; c.name = names[i]; c.anno = "(null)"; c.offset = tot; c.len = (i32)strlen(seqs[i]); c.n_ambs = 0; c.gi = 0; c.is_alt = 0;
            tot += c.len; idx->contigs.push_back(c);
        }
        idx->l_pac = tot;
        // src/BWAIndex.cpp:107-113: the forward pac is made first, then a second pass packs forward + reverse for the BWT;
        // every ambiguous base draws lrand48()&3 in each pass (src/BWAIndex.cpp:217), so with Ns the two disagree.
        std::vector<u8> fwd((size_t)tot / 4 + 2, 0), bwt_src((size_t)tot / 4 + 2, 0);
        for (int pass = 0; pass < 2; ++pass) {
            std::vector<u8> &dst = pass == 0 ? fwd : bwt_src;
            i64 l = 0;
            for (int i = 0; i < n; ++i)
                for (const char *s = seqs[i]; *s; ++s, ++l) {
                    int c = nt4_of((unsigned char)*s);
                    if (c >= 4) c = (int)(lrand48() & 3);
                    dst[l >> 2] |= (u8)(c << ((~l & 3) << 1));
                }
        }
        idx->h_pac = bwt_src;                  // the BWT is built over the second pass
        int rc = construct_common(idx, flags | 1);
        (void)rc;
        idx->h_pac = fwd;                      // idx->pac of the reference = first pass
        // the device text used for extension must be the forward pac of the FIRST pass (what bns_get_seq reads)
        {
            BlobHeader h; CU_CHECK(cudaMemcpy(&h, idx->d_blob, sizeof(h), cudaMemcpyDeviceToHost));
            DevBuf tmp; tmp.reserve(idx->h_pac.size() + 16);
            CU_CHECK(cudaMemcpy(tmp.p, idx->h_pac.data(), idx->h_pac.size(), cudaMemcpyHostToDevice));
            k_text_from_pac<<<nblocks(h.n_text, 256), 256>>>(tmp.as<u8>(), idx->l_pac, (u64 *)((u8 *)idx->d_blob + h.off_text), h.n_text);
            CU_CHECK(cudaDeviceSynchronize());
        }
        fill_view_contigs(idx);
    } catch (const std::exception &e) { delete idx; return fail(B200_ERR_CUDA, e.what()); }
    *out = idx;
    return B200_OK;
}

int b200_index_construct_pac(int64_t l_pac, const uint8_t *pac, int n_seqs, const b200_contig_t *contigs, int flags, b200_index_t **out)
{
    if (!out) return fail(B200_ERR_ARG, "out is NULL");
    *out = nullptr;
    if (l_pac <= 0 || !pac || n_seqs <= 0) return fail(B200_ERR_ARG, "empty reference");
    b200_index *idx = new b200_index;
    try {
        for (int i = 0; i < n_seqs; ++i) {
            ContigMeta c; c.name = contigs[i].name; c.anno = contigs[i].anno ? contigs[i].anno : "(null)"; c.offset = contigs[i].offset;
            c.len = contigs[i].len; c.n_ambs = contigs[i].n_ambs; c.gi = contigs[i].gi; c.is_alt = contigs[i].is_alt;
            idx->contigs.push_back(c);
        }
        idx->l_pac = l_pac;
        idx->h_pac.assign((size_t)l_pac / 4 + 1, 0);                       // bwa's layout keeps l_pac / 4 + 1 bytes; the caller owns ceil(l_pac / 4)
        memcpy(idx->h_pac.data(), pac, (size_t)(l_pac + 3) / 4);
        idx->h_pac.push_back(0);
        construct_common(idx, flags);
        if (!(flags & 1)) { idx->h_pac.clear(); idx->h_pac.shrink_to_fit(); }
        fill_view_contigs(idx);
    } catch (const std::exception &e) { delete idx; return fail(B200_ERR_CUDA, e.what()); }
    *out = idx;
    return B200_OK;
}

static bool file_exists(const std::string &p) { struct stat st; return stat(p.c_str(), &st) == 0 && S_ISREG(st.st_mode); }

int b200_index_load(const char *hint, b200_index_t **out)
{
    if (!out || !hint) return fail(B200_ERR_ARG, "bad argument");
    *out = nullptr;
    std::string prefix;                                   // bwa_idx_infer_prefix (bwa/bwa.c:245-269)
    if (file_exists(std::string(hint) + ".64.bwt")) prefix = std::string(hint) + ".64";
    else if (file_exists(std::string(hint) + ".bwt")) prefix = hint;
    else return fail(B200_ERR_IO, std::string("fail to locate the index files for ") + hint);
    b200_index *idx = new b200_index;
    try {
        {   // .bwt (bwt_restore_bwt, bwa/bwt.c:443-462)
            FILE *fp = fopen((prefix + ".bwt").c_str(), "rb");
            if (!fp) throw std::runtime_error("cannot open .bwt");
            fseek(fp, 0, SEEK_END); long sz = ftell(fp); fseek(fp, 0, SEEK_SET);
            u64 bwt_size = (u64)(sz - 40) >> 2;
            if (fread(&idx->primary, 8, 1, fp) != 1 || fread(idx->L2 + 1, 8, 4, fp) != 4) { fclose(fp); throw std::runtime_error("short .bwt"); }
            idx->h_bwt.resize(bwt_size);
            if (fread(idx->h_bwt.data(), 4, bwt_size, fp) != bwt_size) { fclose(fp); throw std::runtime_error("short .bwt"); }
            fclose(fp);
            idx->L2[0] = 0; idx->seq_len = idx->L2[4];
        }
        {   // .sa (bwt_restore_sa, bwa/bwt.c:421-441)
            FILE *fp = fopen((prefix + ".sa").c_str(), "rb");
            if (!fp) throw std::runtime_error("cannot open .sa");
            u64 primary, skipped[4], intv64, seq_len;
            if (fread(&primary, 8, 1, fp) != 1 || fread(skipped, 8, 4, fp) != 4 || fread(&intv64, 8, 1, fp) != 1 || fread(&seq_len, 8, 1, fp) != 1) { fclose(fp); throw std::runtime_error("short .sa"); }
            if (primary != idx->primary) { fclose(fp); throw std::runtime_error("SA-BWT inconsistency: primary is not the same."); }
            if (seq_len != idx->seq_len) { fclose(fp); throw std::runtime_error("SA-BWT inconsistency: seq_len is not the same."); }
            idx->sa_intv = (int)(u32)intv64;               // written from an int field with sizeof(bwtint_t) (bwa/bwt.c:402)
            u64 n_sa = (idx->seq_len + idx->sa_intv) / idx->sa_intv;
            idx->h_sa.assign(n_sa, 0);
            idx->h_sa[0] = ~0ull;
            if (fread(idx->h_sa.data() + 1, 8, n_sa - 1, fp) != n_sa - 1) { fclose(fp); throw std::runtime_error("short .sa"); }
            fclose(fp);
            if (idx->sa_intv & (idx->sa_intv - 1)) throw std::runtime_error("SA sample interval is not a power of 2.");
        }
        {   // .ann / .amb (bns_restore_core, bwa/bntseq.c:97-166)
            FILE *fp = fopen((prefix + ".ann").c_str(), "r");
            if (!fp) throw std::runtime_error("cannot open .ann");
            long long l_pac; int n_seqs; unsigned seed;
            if (fscanf(fp, "%lld%d%u", &l_pac, &n_seqs, &seed) != 3) { fclose(fp); throw std::runtime_error("bad .ann header"); }
            idx->l_pac = l_pac; idx->seed = seed;
            for (int i = 0; i < n_seqs; ++i) {
                ContigMeta c; char name[8192]; unsigned gi;
                if (fscanf(fp, "%u%8191s", &gi, name) != 2) { fclose(fp); throw std::runtime_error("bad .ann record"); }
                c.gi = gi; c.name = name;
                std::string anno; int ch;
                while ((ch = fgetc(fp)) != '\n' && ch != EOF) anno.push_back((char)ch);
                // the reference stores the text after the separating blank; " (null)" becomes an empty annotation (bwa/bntseq.c:124-126)
                if (!anno.empty() && anno[0] == ' ') anno.erase(0, 1);
                if (anno == "(null)") anno.clear();
                c.anno = anno;
                long long off; int len, nambs;
                if (fscanf(fp, "%lld%d%d", &off, &len, &nambs) != 3) { fclose(fp); throw std::runtime_error("bad .ann record"); }
                c.offset = off; c.len = len; c.n_ambs = nambs; c.is_alt = 0;
                idx->contigs.push_back(c);
            }
            fclose(fp);
            fp = fopen((prefix + ".amb").c_str(), "r");
            if (fp) {
                long long lp; int ns, nh;
                if (fscanf(fp, "%lld%d%d", &lp, &ns, &nh) == 3)
                    for (int i = 0; i < nh; ++i) {
                        long long off; int len; char amb[8];
                        if (fscanf(fp, "%lld%d%7s", &off, &len, amb) != 3) break;
                        HoleMeta hm; hm.offset = off; hm.len = len; hm.amb = amb[0]; idx->holes.push_back(hm);
                    }
                fclose(fp);
            }
            fp = fopen((prefix + ".alt").c_str(), "r");   // bns_restore (bwa/bntseq.c:178-209)
            if (fp) {
                char line[8192];
                while (fgets(line, sizeof line, fp)) {
                    if (line[0] == '@') continue;
                    char *e = line; while (*e && *e != '\t' && *e != '\n' && *e != ' ') ++e; *e = 0;
                    for (auto &c : idx->contigs) if (c.name == line) c.is_alt = 1;
                }
                fclose(fp);
            }
        }
        {   // .pac (bwa/bwa.c:307-311)
            FILE *fp = fopen((prefix + ".pac").c_str(), "rb");
            if (!fp) throw std::runtime_error("cannot open .pac");
            idx->h_pac.assign((size_t)idx->l_pac / 4 + 2, 0);
            size_t got = fread(idx->h_pac.data(), 1, (size_t)idx->l_pac / 4 + 1, fp);
            fclose(fp);
            if (got < (size_t)(idx->l_pac + 3) / 4) throw std::runtime_error("short .pac");
        }
        idx->has_host = true;
        image_from_bwa_arrays(idx);
        fill_view_contigs(idx);
    } catch (const std::exception &e) { delete idx; return fail(B200_ERR_IO, e.what()); }
    *out = idx;
    return B200_OK;
}

int b200_index_write(const b200_index_t *idx, const char *prefix)
{
    if (!idx || !prefix) return fail(B200_ERR_ARG, "bad argument");
    if (!idx->has_host) return fail(B200_ERR_ARG, "index has no host copy (construct with flags bit0)");
    struct stat st;
    if (stat(prefix, &st) == 0 && S_ISDIR(st.st_mode)) return fail(B200_ERR_IO, "prefix is a directory");
    std::string p = prefix;
    FILE *fp = fopen((p + ".bwt").c_str(), "wb");
    if (!fp) return fail(B200_ERR_IO, "cannot write .bwt");
    fwrite(&idx->primary, 8, 1, fp); fwrite(idx->L2 + 1, 8, 4, fp); fwrite(idx->h_bwt.data(), 4, idx->h_bwt.size(), fp); fclose(fp);
    fp = fopen((p + ".sa").c_str(), "wb");
    if (!fp) return fail(B200_ERR_IO, "cannot write .sa");
    u64 intv64 = (u64)(u32)idx->sa_intv;
    fwrite(&idx->primary, 8, 1, fp); fwrite(idx->L2 + 1, 8, 4, fp); fwrite(&intv64, 8, 1, fp); fwrite(&idx->seq_len, 8, 1, fp);
    fwrite(idx->h_sa.data() + 1, 8, idx->h_sa.size() - 1, fp); fclose(fp);
    fp = fopen((p + ".ann").c_str(), "w");
    if (!fp) return fail(B200_ERR_IO, "cannot write .ann");
    fprintf(fp, "%lld %d %u\n", (long long)idx->l_pac, (int)idx->contigs.size(), idx->seed);
    for (auto &c : idx->contigs) {
        fprintf(fp, "%d %s", (int)c.gi, c.name.c_str());
        if (!c.anno.empty()) fprintf(fp, " %s\n", c.anno.c_str()); else fprintf(fp, "\n");
        fprintf(fp, "%lld %d %d\n", (long long)c.offset, c.len, c.n_ambs);
    }
    fclose(fp);
    fp = fopen((p + ".amb").c_str(), "w");
    if (!fp) return fail(B200_ERR_IO, "cannot write .amb");
    fprintf(fp, "%lld %d %u\n", (long long)idx->l_pac, (int)idx->contigs.size(), (unsigned)idx->holes.size());
    for (auto &h : idx->holes) fprintf(fp, "%lld %d %c\n", (long long)h.offset, h.len, h.amb);
    fclose(fp);
    fp = fopen((p + ".pac").c_str(), "wb");            // seqlib_write_pac_to_file (src/BWAIndex.cpp:360-380)
    if (!fp) return fail(B200_ERR_IO, "cannot write .pac");
    i64 l_pac = idx->l_pac; u8 ct;
    fwrite(idx->h_pac.data(), 1, (size_t)((l_pac >> 2) + ((l_pac & 3) == 0 ? 0 : 1)), fp);
    if (l_pac % 4 == 0) { ct = 0; fwrite(&ct, 1, 1, fp); }
    ct = (u8)(l_pac % 4); fwrite(&ct, 1, 1, fp);
    fclose(fp);
    return B200_OK;
}

void b200_index_destroy(b200_index_t *idx) { delete idx; }

int b200_index_view(const b200_index_t *idx, b200_index_view_t *v)
{
    if (!idx || !v) return fail(B200_ERR_ARG, "bad argument");
    if (!idx->has_host) return fail(B200_ERR_ARG, "index has no host copy");
    v->primary = idx->primary; memcpy(v->L2, idx->L2, sizeof(v->L2)); v->seq_len = idx->seq_len; v->bwt_size = idx->h_bwt.size();
    v->bwt = idx->h_bwt.data(); v->sa_intv = idx->sa_intv; v->n_sa = idx->h_sa.size(); v->sa = idx->h_sa.data();
    v->l_pac = idx->l_pac; v->pac = idx->h_pac.data(); v->n_seqs = (int32_t)idx->contigs.size(); v->contigs = idx->view_contigs.data();
    return B200_OK;
}

int64_t b200_index_blob_bytes(const b200_index_t *idx) { return idx ? idx->blob_bytes : 0; }

int b200_index_export_blob(const b200_index_t *idx, void *dev_dst)
{
    if (!idx || !dev_dst) return fail(B200_ERR_ARG, "bad argument");
    cudaError_t e = cudaMemcpy(dev_dst, idx->d_blob, (size_t)idx->blob_bytes, cudaMemcpyDeviceToDevice);
    if (e != cudaSuccess) return fail(B200_ERR_CUDA, cudaGetErrorString(e));
    return B200_OK;
}

int b200_index_attach_blob(void *dev_blob, int64_t nbytes, b200_index_t **out)
{
    if (!dev_blob || !out) return fail(B200_ERR_ARG, "bad argument");
    *out = nullptr;
    BlobHeader h;
    cudaError_t e = cudaMemcpy(&h, dev_blob, sizeof(h), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return fail(B200_ERR_CUDA, cudaGetErrorString(e));
    if (h.magic != BLOB_MAGIC || (int64_t)h.total_bytes > nbytes) return fail(B200_ERR_ARG, "not an index image");
    b200_index *idx = new b200_index;
    try {
        std::vector<i64> coff(h.n_seqs + 1); std::vector<i32> calt(h.n_seqs + 1); std::vector<char> names(h.names_bytes + 1);
        u8 *b = (u8 *)dev_blob;
        CU_CHECK(cudaMemcpy(coff.data(), b + h.off_coff, ((size_t)h.n_seqs + 1) * 8, cudaMemcpyDeviceToHost));
        CU_CHECK(cudaMemcpy(calt.data(), b + h.off_calt, (size_t)h.n_seqs * 4, cudaMemcpyDeviceToHost));
        if (h.names_bytes) CU_CHECK(cudaMemcpy(names.data(), b + h.off_names, h.names_bytes, cudaMemcpyDeviceToHost));
        const char *p = names.data();
        for (int i = 0; i < h.n_seqs; ++i) {
            ContigMeta c; c.name = p; p += c.name.size() + 1; c.anno = p; p += c.anno.size() + 1;
            c.offset = coff[i]; c.len = (i32)(coff[i + 1] - coff[i]); c.n_ambs = 0; c.gi = 0; c.is_alt = calt[i];
            idx->contigs.push_back(c);
        }
        bind_blob(idx, dev_blob, h);
        idx->owns_blob = false;
        fill_view_contigs(idx);
    } catch (const std::exception &ex) { delete idx; return fail(B200_ERR_CUDA, ex.what()); }
    *out = idx;
    return B200_OK;
}

int b200_index_n_seqs(const b200_index_t *idx) { return idx ? (int)idx->contigs.size() : 0; }
const char *b200_index_seq_name(const b200_index_t *idx, int rid)
{
    if (!idx || rid < 0 || rid >= (int)idx->contigs.size()) return nullptr;
    return idx->contigs[rid].name.c_str();
}
int64_t b200_index_seq_len(const b200_index_t *idx, int rid)
{
    if (!idx || rid < 0 || rid >= (int)idx->contigs.size()) return -1;
    return idx->contigs[rid].len;
}
int64_t b200_index_l_pac(const b200_index_t *idx) { return idx ? idx->l_pac : 0; }

} // extern "C"
