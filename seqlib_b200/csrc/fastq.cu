// fastq.cu -- FASTA/FASTQ ingest behind b200_fastq_* (SURVEY.md 8f row 2).
//   stream parser  <- FastqReader::GetNextSequence (src/FastqReader.cpp:37-59) = kseq_read (bwa/kseq.h:176-226) over
//                     ks_getc / ks_getuntil2 (bwa/kseq.h:69-141) on a gzFile: written from the documented behaviour of those
//                     three functions (what ends a name, when a trailing CR is dropped, how blank lines, multi-line records
//                     and short quality strings are treated), not from their text; records go straight into the flat
//                     (bytes, offsets) layout b200_mem_align_batch() takes.
//   device parser  strict four-line FASTQ: newline positions by a count / scan / write pass pair (16 bytes per thread), one thread per
//                     record for the checks and the name/comment split, one warp per record for the gather.
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <zlib.h>
#include <unistd.h>
#include <sys/stat.h>
#include <cuda_runtime.h>
#include <cub/cub.cuh>
#include "engine.cuh"          // DevPool: device buffers come from the process-wide pool, like the aligner's
#include "../../include/seqlib_b200.h"

namespace b200 {

namespace {

// growable byte array (std::vector<char> zero-fills on resize: too slow for GB-sized batches)
struct Bytes {
    char *p = nullptr; size_t n = 0, cap = 0;
    ~Bytes() { free(p); }
    void need(size_t extra)
    {
        if (n + extra <= cap) return;
        size_t c = cap ? cap : 4096;
        while (c < n + extra) c *= 2;
        p = (char *)realloc(p, c); cap = c;
    }
    void append(const unsigned char *s, size_t k) { need(k); memcpy(p + n, s, k); n += k; }
    void push(char c) { need(1); p[n++] = c; }
};

enum { SEP_SPACE = 0, SEP_LINE = 2 };

struct Stream {
    gzFile fp = nullptr;
    const unsigned char *buf = nullptr;
    std::vector<unsigned char> own;
    long beg = 0, end = 0;
    bool eof = false, mem = false, io_error = false;
    int last_char = 0;
    bool comment_buf = false, qual_buf = false;      // kseq's comment.s / qual.s have been allocated (FastqReader.cpp:49-56 tests them)

    bool refill()                                      // false: end of input
    {
        if (mem) { eof = true; return false; }
        beg = 0;
        int k = gzread(fp, own.data(), (unsigned)own.size());
        if (k <= 0) {                                   // corrupt / truncated gzip stream: not a clean end of input (next_batch reports B200_ERR_IO)
            int zerr = 0;
            gzerror(fp, &zerr);
            if (k < 0 || (zerr != Z_OK && zerr != Z_STREAM_END)) io_error = true;
        }
        end = k > 0 ? k : 0;
        if (end == 0) { eof = true; return false; }
        return true;
    }
    int getc_()
    {
        if (beg >= end) {
            if (eof) return -1;
            if (!refill()) return -1;
        }
        return buf[beg++];
    }
    // appends to `s` (the field started at s.n == f0) up to the delimiter; *dret = the delimiter met (0: none)
    long getuntil(int delim, Bytes &s, size_t f0, int *dret)
    {
        bool gotany = false;
        if (dret) *dret = 0;
        for (;;) {
            if (beg >= end) {
                if (eof || !refill()) break;
            }
            long i;
            if (delim == SEP_LINE) {
                const void *q = memchr(buf + beg, '\n', (size_t)(end - beg));
                i = q ? (long)((const unsigned char *)q - buf) : end;
            } else {
                for (i = beg; i < end; ++i) {
                    unsigned char c = buf[i];
                    if (c == ' ' || (c >= '\t' && c <= '\r')) break;
                }
            }
            gotany = true;
            s.append(buf + beg, (size_t)(i - beg));
            beg = i + 1;
            if (i < end) { if (dret) *dret = buf[i]; break; }
        }
        if (!gotany && eof && beg >= end) return -1;
        if (delim == SEP_LINE && s.n - f0 > 1 && s.p[s.n - 1] == '\r') --s.n;
        return (long)(s.n - f0);
    }
};

} // namespace
} // namespace b200

using namespace b200;

struct b200_fastq {
    Stream st;
    Bytes seq, qual, name, com;
    std::vector<int64_t> seq_off, qual_off, name_off, com_off;
    std::vector<uint8_t> has;
    // device parser buffers
    void *d_text = nullptr; size_t d_text_cap = 0;
    void *d_tmp = nullptr; size_t d_tmp_cap = 0;
    void *d_misc = nullptr; size_t d_misc_cap = 0;
    ~b200_fastq()
    {
        if (st.fp) gzclose(st.fp);
        // every parse call ends with blocking copies, so nothing is in flight on these blocks
        if (d_text) b200::dev_pool().put(d_text, d_text_cap);
        if (d_tmp) b200::dev_pool().put(d_tmp, d_tmp_cap);
        if (d_misc) b200::dev_pool().put(d_misc, d_misc_cap);
    }
};

// one record (kseq_read's return value: >= 0 length, -1 end of input, -2 truncated quality); fields appended to the flat buffers
static long read_record(b200_fastq &R)
{
    Stream &S = R.st;
    int c;
    if (S.last_char == 0) {
        while ((c = S.getc_()) != -1 && c != '>' && c != '@') {}
        if (c == -1) return -1;
        S.last_char = c;
    }
    const size_t n0 = R.name.n, c0 = R.com.n, s0 = R.seq.n, q0 = R.qual.n;
    if (S.getuntil(SEP_SPACE, R.name, n0, &c) < 0) return -1;
    if (c != '\n') { if (S.getuntil(SEP_LINE, R.com, c0, nullptr) >= 0) S.comment_buf = true; }
    while ((c = S.getc_()) != -1 && c != '>' && c != '+' && c != '@') {
        if (c == '\n') continue;
        R.seq.push((char)c);
        S.getuntil(SEP_LINE, R.seq, s0, nullptr);
    }
    if (c == '>' || c == '@') S.last_char = c;
    if (c != '+') return (long)(R.seq.n - s0);
    S.qual_buf = true;
    while ((c = S.getc_()) != -1 && c != '\n') {}
    if (c == -1) return -2;
    while (S.getuntil(SEP_LINE, R.qual, q0, nullptr) >= 0 && R.qual.n - q0 < R.seq.n - s0) {}
    S.last_char = 0;
    if (R.seq.n - s0 != R.qual.n - q0) return -2;
    return (long)(R.seq.n - s0);
}

static void fill_batch(b200_fastq &R, b200_fastq_batch_t *out, int status, int on_device)
{
    out->n = (int64_t)R.seq_off.size() - 1;
    out->seq = R.seq.p ? R.seq.p : ""; out->seq_off = R.seq_off.data();
    out->qual = R.qual.p ? R.qual.p : ""; out->qual_off = R.qual_off.data();
    out->name = R.name.p ? R.name.p : ""; out->name_off = R.name_off.data();
    out->comment = R.com.p ? R.com.p : ""; out->comment_off = R.com_off.data();
    out->status = status; out->parsed_on_device = on_device;
    out->has = R.has.data();
}

extern "C" {

int b200_fastq_open(const char *path, b200_fastq_t **out)
{
    if (!path || !out) { set_error("b200_fastq_open: null argument"); return B200_ERR_ARG; }
    *out = nullptr;
    gzFile fp;
    if (strcmp(path, "-") != 0) {
        struct stat sb;
        if (stat(path, &sb) != 0) { set_error(std::string("FastqReader: Failed to read non-existant file ") + path); return B200_ERR_IO; }
        fp = gzopen(path, "r");
    } else fp = gzdopen(fileno(stdin), "r");
    if (!fp) { set_error(std::string("FastqReader: Failed to read ") + path); return B200_ERR_IO; }
    gzbuffer(fp, 1u << 20);
    b200_fastq *R = new b200_fastq;
    R->st.fp = fp;
    R->st.own.resize(1u << 20);
    R->st.buf = R->st.own.data();
    *out = R;
    return 0;
}

int b200_fastq_open_mem(const char *text, int64_t len, b200_fastq_t **out)
{
    if ((!text && len) || len < 0 || !out) { set_error("b200_fastq_open_mem: bad argument"); return B200_ERR_ARG; }
    b200_fastq *R = new b200_fastq;
    R->st.mem = true; R->st.buf = (const unsigned char *)text; R->st.beg = 0; R->st.end = (long)len;
    *out = R;
    return 0;
}

int b200_fastq_next_batch(b200_fastq_t *R, int64_t max_records, b200_fastq_batch_t *out)
{
    if (!R || !out || max_records < 0) { set_error("b200_fastq_next_batch: bad argument"); return B200_ERR_ARG; }
    R->seq.n = R->qual.n = R->name.n = R->com.n = 0;
    R->seq_off.assign(1, 0); R->qual_off.assign(1, 0); R->name_off.assign(1, 0); R->com_off.assign(1, 0);
    int status = 0;
    R->has.clear();
    while ((int64_t)R->seq_off.size() - 1 < max_records) {
        const size_t n0 = R->name.n, c0 = R->com.n, s0 = R->seq.n, q0 = R->qual.n;
        long r = read_record(*R);
        if (r < 0) {                                  // drop what the failed record appended
            R->name.n = n0; R->com.n = c0; R->seq.n = s0; R->qual.n = q0;
            status = r == -1 ? 1 : -2;
            break;
        }
        R->seq_off.push_back((int64_t)R->seq.n); R->qual_off.push_back((int64_t)R->qual.n);
        R->name_off.push_back((int64_t)R->name.n); R->com_off.push_back((int64_t)R->com.n);
        R->has.push_back((uint8_t)((R->st.comment_buf ? 1 : 0) | (R->st.qual_buf ? 2 : 0)));
    }
    fill_batch(*R, out, status, 0);
    if (R->st.io_error) {      // gzread failed (corrupt or truncated gzip): the records before the damage are in the batch, the call says so
        set_error("b200_fastq_next_batch: read error in the compressed stream (truncated or corrupt input)");
        return B200_ERR_IO;
    }
    return 0;
}

// FastqReader::GetNextSequence only assigns Com / Qual when kseq has allocated those strings (src/FastqReader.cpp:49-56):
// bit 0 = comment buffer exists, bit 1 = quality buffer exists, as of the last record returned
int b200_fastq_buffers_seen(const b200_fastq_t *R) { return R ? (R->st.comment_buf ? 1 : 0) | (R->st.qual_buf ? 2 : 0) : 0; }

void b200_fastq_close(b200_fastq_t *R) { delete R; }

} // extern "C"

// ---------------------------------------------------------------------------------------------------------------------
// device parser
namespace b200 {

// newline positions: 16 text bytes per thread (one 128-bit load, SWAR byte compare), 4 KB per block; pass 1 counts per
// block, pass 2 (after an exclusive scan of the block counts) writes the positions in order
#define NL_THREADS 256
#define NL_BLOCK_BYTES (NL_THREADS * 16)
__device__ __forceinline__ void nl_load(const char *__restrict__ t, int64_t len, int64_t base, unsigned m[4])
{
    if (base + 16 <= len) {
        uint4 v = *reinterpret_cast<const uint4 *>(t + base);
        m[0] = __vcmpeq4(v.x, 0x0a0a0a0au); m[1] = __vcmpeq4(v.y, 0x0a0a0a0au); m[2] = __vcmpeq4(v.z, 0x0a0a0a0au); m[3] = __vcmpeq4(v.w, 0x0a0a0a0au);
    } else {
        for (int w = 0; w < 4; ++w) {
            m[w] = 0;
            for (int b = 0; b < 4; ++b) { int64_t p = base + 4 * w + b; if (p < len && t[p] == '\n') m[w] |= 0xffu << (8 * b); }
        }
    }
}
__global__ void __launch_bounds__(NL_THREADS) k_nl_count(const char *__restrict__ t, int64_t len, int64_t *__restrict__ blk_cnt)
{
    typedef cub::BlockReduce<unsigned, NL_THREADS> Red;
    __shared__ typename Red::TempStorage tmp;
    unsigned m[4];
    nl_load(t, len, (int64_t)blockIdx.x * NL_BLOCK_BYTES + threadIdx.x * 16, m);
    unsigned c = (__popc(m[0]) + __popc(m[1]) + __popc(m[2]) + __popc(m[3])) >> 3;
    unsigned tot = Red(tmp).Sum(c);
    if (threadIdx.x == 0) blk_cnt[blockIdx.x] = tot;
}
__global__ void __launch_bounds__(NL_THREADS) k_nl_write(const char *__restrict__ t, int64_t len, const int64_t *__restrict__ blk_off, int64_t *__restrict__ nl)
{
    typedef cub::BlockScan<unsigned, NL_THREADS> Scan;
    __shared__ typename Scan::TempStorage tmp;
    unsigned m[4];
    const int64_t base = (int64_t)blockIdx.x * NL_BLOCK_BYTES + threadIdx.x * 16;
    nl_load(t, len, base, m);
    unsigned c = (__popc(m[0]) + __popc(m[1]) + __popc(m[2]) + __popc(m[3])) >> 3, at;
    Scan(tmp).ExclusiveSum(c, at);
    if (!c) return;
    int64_t *o = nl + blk_off[blockIdx.x] + at;
#pragma unroll
    for (int w = 0; w < 4; ++w)
#pragma unroll
        for (int b = 0; b < 4; ++b) if (m[w] >> (8 * b) & 1u) *o++ = base + 4 * w + b;
}

struct RecSpan { int64_t name_b, com_b, seq_b, qual_b; int32_t name_l, com_l, seq_l, qual_l; };

__device__ __forceinline__ bool is_space_(unsigned char c) { return c == ' ' || (c >= '\t' && c <= '\r'); }

// one thread per record: lines 4r .. 4r+3; line k spans [start_k, end_k) where end_k is its '\n' (or len for an unterminated last line)
__global__ void k_fastq_records(const char *__restrict__ t, int64_t len, const int64_t *__restrict__ nl, int64_t n_nl, int64_t n_rec,
                                RecSpan *__restrict__ out, int64_t *__restrict__ lens /* 4 arrays of n_rec */, uint8_t *__restrict__ has_com, int *__restrict__ bad)
{
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rec) return;
    int64_t b[4], e[4];
    for (int k = 0; k < 4; ++k) {
        int64_t li = 4 * r + k;
        b[k] = li == 0 ? 0 : nl[li - 1] + 1;
        e[k] = li < n_nl ? nl[li] : len;
    }
    bool ok = e[0] > b[0] && t[b[0]] == '@' && e[2] > b[2] && t[b[2]] == '+' && e[1] > b[1];
    RecSpan s;
    s.name_b = b[0] + 1; s.name_l = 0; s.com_b = e[0]; s.com_l = 0; s.seq_b = b[1]; s.seq_l = 0; s.qual_b = b[3]; s.qual_l = 0;
    if (ok) {
        int64_t p = b[0] + 1;
        while (p < e[0] && !is_space_((unsigned char)t[p])) ++p;          // the name ends at the first isspace() character
        s.name_l = (int32_t)(p - s.name_b);
        if (p < e[0]) {                                                     // delimiter inside the line: the rest is the comment
            s.com_b = p + 1;
            int64_t l = e[0] - s.com_b;
            if (l > 1 && t[e[0] - 1] == '\r') --l;
            s.com_l = (int32_t)l;
        }
        char c0 = t[b[1]];
        if (c0 == '>' || c0 == '+' || c0 == '@' || c0 == '\r') ok = false;
        int64_t l = e[1] - b[1];
        if (l > 1 && t[e[1] - 1] == '\r') --l;
        s.seq_l = (int32_t)l;
        l = e[3] - b[3];
        if (l > 1 && t[e[3] - 1] == '\r') --l;
        s.qual_l = (int32_t)l;
        if (s.qual_l != s.seq_l) ok = false;
        if ((e[0] - b[0]) > 0x3fffffff || (e[1] - b[1]) > 0x3fffffff) ok = false;
    }
    if (!ok) { atomicExch(bad, 1); s.name_l = s.com_l = s.seq_l = s.qual_l = 0; }
    out[r] = s;
    has_com[r] = s.com_b < e[0] || (s.com_b == e[0] && s.name_b + s.name_l < e[0]) ? 1 : 0;      // the name ended at a delimiter inside the line
    lens[r] = s.name_l; lens[n_rec + r] = s.com_l; lens[2 * n_rec + r] = s.seq_l; lens[3 * n_rec + r] = s.qual_l;
}

// one warp per record: the four fields to their places in the contiguous buffers
__global__ void k_fastq_gather(const char *__restrict__ t, const RecSpan *__restrict__ spans, const int64_t *__restrict__ offs /* 4 x (n_rec + 1) */,
                               int64_t n_rec, char *__restrict__ name, char *__restrict__ com, char *__restrict__ seq, char *__restrict__ qual)
{
    int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (r >= n_rec) return;
    RecSpan s = spans[r];
    const int64_t *o = offs;
    for (int i = lane; i < s.name_l; i += 32) name[o[r] + i] = t[s.name_b + i];
    o += n_rec + 1;
    for (int i = lane; i < s.com_l; i += 32) com[o[r] + i] = t[s.com_b + i];
    o += n_rec + 1;
    for (int i = lane; i < s.seq_l; i += 32) seq[o[r] + i] = t[s.seq_b + i];
    o += n_rec + 1;
    for (int i = lane; i < s.qual_l; i += 32) qual[o[r] + i] = t[s.qual_b + i];
}

__global__ void k_fastq_last_off(const int64_t *__restrict__ lens, int64_t *__restrict__ offs, int64_t n_rec)
{
    int f = threadIdx.x;
    if (f < 4) offs[(n_rec + 1) * f + n_rec] = offs[(n_rec + 1) * f + n_rec - 1] + lens[n_rec * f + n_rec - 1];
}

static bool dev_reserve(void *&p, size_t &cap, size_t need)
{
    if (need <= cap) return true;
    if (p) dev_pool().put(p, cap);          // callers only grow a buffer after the work that used it was synchronised
    p = nullptr; cap = 0;
    try { p = dev_pool().get(need + need / 4 + 4096, cap); }
    catch (const CudaError &) { cudaGetLastError(); return false; }
    return true;
}

} // namespace b200

extern "C" int b200_fastq_parse_device(b200_fastq_t *R, const char *text, int64_t len, b200_fastq_batch_t *out)
{
    if (!R || !out || (!text && len) || len < 0) { set_error("b200_fastq_parse_device: bad argument"); return B200_ERR_ARG; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); set_error("b200_fastq_parse_device: no CUDA device"); return B200_ERR_CUDA; }
#define FQ_CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { set_error(std::string(#x) + ": " + cudaGetErrorString(e_)); cudaGetLastError(); return B200_ERR_CUDA; } } while (0)
    R->seq.n = R->qual.n = R->name.n = R->com.n = 0;
    R->seq_off.assign(1, 0); R->qual_off.assign(1, 0); R->name_off.assign(1, 0); R->com_off.assign(1, 0);
    if (len == 0) { fill_batch(*R, out, 1, 1); return 0; }
    if (!dev_reserve(R->d_text, R->d_text_cap, (size_t)len)) { set_error("b200_fastq_parse_device: out of device memory"); return B200_ERR_NOMEM; }
    char *d_t = (char *)R->d_text;
    FQ_CU(cudaMemcpy(d_t, text, (size_t)len, cudaMemcpyHostToDevice));
    // newline positions
    const int64_t max_nl = len / 2 + 2;            // a strict record has at least 2 bytes per line on two of its four lines; more newlines = not strict
    const int64_t n_blk = (len + NL_BLOCK_BYTES - 1) / NL_BLOCK_BYTES;
    size_t scan_tb0 = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, scan_tb0, (const int64_t *)nullptr, (int64_t *)nullptr, (int)(n_blk + 1));
    size_t misc_bytes = sizeof(int64_t) * (size_t)(max_nl + 8) + 64;
    if (n_blk + 1 > 0x7fffffff) { set_error("b200_fastq_parse_device: text larger than 8 TB"); return B200_ERR_LIMIT; }
    if (!dev_reserve(R->d_misc, R->d_misc_cap, misc_bytes)) { set_error("b200_fastq_parse_device: out of device memory"); return B200_ERR_NOMEM; }
    if (!dev_reserve(R->d_tmp, R->d_tmp_cap, sizeof(int64_t) * 2 * (size_t)(n_blk + 1) + scan_tb0 + 512)) { set_error("b200_fastq_parse_device: out of device memory"); return B200_ERR_NOMEM; }
    int64_t *d_nl = (int64_t *)R->d_misc;
    int *d_bad = (int *)(d_nl + max_nl + 2);
    int64_t n_nl = 0;
    {
        int64_t *d_cnt = (int64_t *)R->d_tmp, *d_off = d_cnt + (n_blk + 1);
        void *d_scan0 = (void *)(d_off + (n_blk + 1));
        FQ_CU(cudaMemset(d_cnt + n_blk, 0, sizeof(int64_t)));
        k_nl_count<<<(unsigned)n_blk, NL_THREADS>>>(d_t, len, d_cnt);
        FQ_CU(cudaGetLastError());
        FQ_CU(cub::DeviceScan::ExclusiveSum(d_scan0, scan_tb0, d_cnt, d_off, (int)(n_blk + 1)));
        FQ_CU(cudaMemcpy(&n_nl, d_off + n_blk, sizeof(n_nl), cudaMemcpyDeviceToHost));
        if (n_nl > max_nl) { set_error("b200_fastq_parse_device: not strict four-line FASTQ (blank lines)"); return B200_ERR_ARG; }
        k_nl_write<<<(unsigned)n_blk, NL_THREADS>>>(d_t, len, d_off, d_nl);
        FQ_CU(cudaGetLastError());
        FQ_CU(cudaStreamSynchronize(0));           // d_tmp is reused below
    }
    char last = text[len - 1];
    int64_t n_lines = n_nl + (last == '\n' ? 0 : 1);
    if (n_lines % 4 != 0) { set_error("b200_fastq_parse_device: not strict four-line FASTQ (line count)"); return B200_ERR_ARG; }
    const int64_t n_rec = n_lines / 4;
    // spans + lengths + offsets
    size_t spans_b = sizeof(RecSpan) * (size_t)n_rec, lens_b = sizeof(int64_t) * 4 * (size_t)n_rec, offs_b = sizeof(int64_t) * 4 * (size_t)(n_rec + 1);
    size_t scan_tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, scan_tb, (const int64_t *)nullptr, (int64_t *)nullptr, (int)n_rec);
    size_t need = spans_b + lens_b + offs_b + scan_tb + (size_t)n_rec + 2048;
    if (!dev_reserve(R->d_tmp, R->d_tmp_cap, need)) { set_error("b200_fastq_parse_device: out of device memory"); return B200_ERR_NOMEM; }
    char *w = (char *)R->d_tmp;
    RecSpan *d_spans = (RecSpan *)w; w += (spans_b + 255) & ~(size_t)255;
    int64_t *d_lens = (int64_t *)w; w += (lens_b + 255) & ~(size_t)255;
    int64_t *d_offs = (int64_t *)w; w += (offs_b + 255) & ~(size_t)255;
    void *d_scan = w; w += (scan_tb + 255) & ~(size_t)255;
    uint8_t *d_has = (uint8_t *)w;
    FQ_CU(cudaMemset(d_bad, 0, sizeof(int)));
    k_fastq_records<<<(unsigned)((n_rec + 255) / 256), 256>>>(d_t, len, d_nl, n_nl, n_rec, d_spans, d_lens, d_has, d_bad);
    FQ_CU(cudaGetLastError());
    for (int f = 0; f < 4; ++f)
        FQ_CU(cub::DeviceScan::ExclusiveSum(d_scan, scan_tb, d_lens + (size_t)f * n_rec, d_offs + (size_t)f * (n_rec + 1), (int)n_rec));
    k_fastq_last_off<<<1, 32>>>(d_lens, d_offs, n_rec);
    FQ_CU(cudaGetLastError());
    int bad = 0;
    FQ_CU(cudaMemcpy(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost));
    if (bad) { set_error("b200_fastq_parse_device: not strict four-line FASTQ (record check)"); return B200_ERR_ARG; }
    R->name_off.resize((size_t)n_rec + 1); R->com_off.resize((size_t)n_rec + 1); R->seq_off.resize((size_t)n_rec + 1); R->qual_off.resize((size_t)n_rec + 1);
    FQ_CU(cudaMemcpy(R->name_off.data(), d_offs, sizeof(int64_t) * (size_t)(n_rec + 1), cudaMemcpyDeviceToHost));
    FQ_CU(cudaMemcpy(R->com_off.data(), d_offs + (n_rec + 1), sizeof(int64_t) * (size_t)(n_rec + 1), cudaMemcpyDeviceToHost));
    FQ_CU(cudaMemcpy(R->seq_off.data(), d_offs + 2 * (n_rec + 1), sizeof(int64_t) * (size_t)(n_rec + 1), cudaMemcpyDeviceToHost));
    FQ_CU(cudaMemcpy(R->qual_off.data(), d_offs + 3 * (n_rec + 1), sizeof(int64_t) * (size_t)(n_rec + 1), cudaMemcpyDeviceToHost));
    const size_t nb = (size_t)R->name_off[n_rec], cb = (size_t)R->com_off[n_rec], sb = (size_t)R->seq_off[n_rec], qb = (size_t)R->qual_off[n_rec];
    // the gathered fields reuse the newline-position buffer (no longer needed) when it is large enough, else a fresh one
    size_t out_b = nb + cb + sb + qb + 1024;
    if (!dev_reserve(R->d_misc, R->d_misc_cap, out_b)) { set_error("b200_fastq_parse_device: out of device memory"); return B200_ERR_NOMEM; }
    char *d_name = (char *)R->d_misc, *d_com = d_name + nb, *d_seq = d_com + cb, *d_qual = d_seq + sb;
    k_fastq_gather<<<(unsigned)((n_rec * 32 + 255) / 256), 256>>>(d_t, d_spans, d_offs, n_rec, d_name, d_com, d_seq, d_qual);
    FQ_CU(cudaGetLastError());
    R->name.n = 0; R->name.need(nb + 1); R->name.n = nb;
    R->com.n = 0; R->com.need(cb + 1); R->com.n = cb;
    R->seq.n = 0; R->seq.need(sb + 1); R->seq.n = sb;
    R->qual.n = 0; R->qual.need(qb + 1); R->qual.n = qb;
    FQ_CU(cudaMemcpy(R->name.p, d_name, nb, cudaMemcpyDeviceToHost));
    FQ_CU(cudaMemcpy(R->com.p, d_com, cb, cudaMemcpyDeviceToHost));
    FQ_CU(cudaMemcpy(R->seq.p, d_seq, sb, cudaMemcpyDeviceToHost));
    FQ_CU(cudaMemcpy(R->qual.p, d_qual, qb, cudaMemcpyDeviceToHost));
    R->has.resize((size_t)n_rec);
    FQ_CU(cudaMemcpy(R->has.data(), d_has, (size_t)n_rec, cudaMemcpyDeviceToHost));
    {   // running "comment string exists" flag; every strict record has a '+' line
        uint8_t seen = (uint8_t)((R->st.comment_buf ? 1 : 0) | 2);
        for (int64_t i = 0; i < n_rec; ++i) { seen |= R->has[i] & 1; R->has[i] = seen; }
        if (n_rec) { R->st.comment_buf = (seen & 1) != 0; R->st.qual_buf = true; }
    }
    fill_batch(*R, out, 1, 1);
    return 0;
#undef FQ_CU
}
