// fml_emul.cpp -- TEST HARNESS ONLY.  The per-read device code of the BFC stage (seqlib_b200/csrc/bfc.cuh) compiled for
// the host, with the count table built the way fml.cu builds it on the device (records -> sort -> runs -> open-addressing
// slots) but serially.  Lets the CPU suite check the kernel logic against the reference library without a GPU; never
// linked into libseqlib_b200.so.
#include <cstdint>
#include <cstring>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include "../../include/seqlib_b200.h"
#include "../../seqlib_b200/csrc/fml_host.h"

using namespace b200;

namespace {

struct HostTable {
    std::vector<CountSlot> slots;
    int k, l_pre;
    uint64_t hist[256], hist_high[64];
    uint64_t n_kmers, n_distinct;
    CountTable view() const { CountTable t; t.slots = slots.data(); t.mask = slots.size() - 1; t.k = k; t.l_pre = l_pre; return t; }
};

struct Rec { u64 lo; u32 hi; };

void build_table(HostTable &T, int64_t n, const char *seqs, const char *quals, const int64_t *off, int k, int q, int l_pre_in)
{
    int l_pre = bfc_l_pre(k, l_pre_in);
    if (l_pre < 0) l_pre = 0;
    T.k = k; T.l_pre = l_pre;
    std::vector<Rec> recs;
    std::vector<u64> lo; std::vector<u32> hi;
    for (int64_t i = 0; i < n; ++i) {
        int len = (int)(off[i + 1] - off[i]);
        lo.resize(len + 1); hi.resize(len + 1);
        int m = count_read_kmers(k, l_pre, q, seqs + off[i], quals ? quals + off[i] : nullptr, len, lo.data(), hi.data());
        for (int j = 0; j < m; ++j) recs.push_back(Rec{lo[j], hi[j]});
    }
    std::sort(recs.begin(), recs.end(), [](const Rec &a, const Rec &b) {
        u32 ah = a.hi & 0x7fffffffu, bh = b.hi & 0x7fffffffu;
        return ah != bh ? ah < bh : a.lo < b.lo;
    });
    T.n_kmers = recs.size();
    memset(T.hist, 0, sizeof(T.hist)); memset(T.hist_high, 0, sizeof(T.hist_high));
    size_t nd = 0;
    for (size_t s = 0; s < recs.size();) {
        size_t e = s;
        while (e < recs.size() && recs[e].lo == recs[s].lo && ((recs[e].hi ^ recs[s].hi) & 0x7fffffffu) == 0) ++e;
        ++nd; s = e;
    }
    T.n_distinct = nd;
    u64 cap = 1024;
    while (cap < 2 * nd) cap <<= 1;
    T.slots.assign(cap, CountSlot{0, 0});
    for (size_t s = 0; s < recs.size();) {
        size_t e = s; u64 high = 0;
        while (e < recs.size() && recs[e].lo == recs[s].lo && ((recs[e].hi ^ recs[s].hi) & 0x7fffffffu) == 0) { high += recs[e].hi >> 31; ++e; }
        u64 tot = e - s;
        u32 cnt = tot > 255 ? 255u : (u32)tot, hc = high > 63 ? 63u : (u32)high;
        KmerKey key; key.lo = recs[s].lo; key.hi = recs[s].hi & 0x7fffffffu;
        u64 j = count_slot_hash(key) & (cap - 1);
        while (T.slots[j].w1) j = (j + 1) & (cap - 1);
        T.slots[j].w0 = key.lo; T.slots[j].w1 = key.hi << 16 | 1ull << 15 | (u64)(hc << 8 | cnt);
        ++T.hist[cnt]; ++T.hist_high[hc];
        s = e;
    }
}

} // namespace

extern "C" {

// Same contract as b200_fml_correct_flat.  stat[0] = largest search stack, stat[1] = largest heap, stat[2] = table probes,
// codes[i] (optional) = ec_code of read i.
int fml_emul_correct_flat(const b200_fml_opt_t *opt, int flt_uniq, int64_t n, char *seqs, char *quals, const int64_t *off,
                          int32_t *len_out, float *kcov_out, uint64_t *hist_out /* 320, optional */, int64_t *stat, uint8_t *codes)
{
    BfcOpt bo; bfc_opt_defaults(bo);
    bo.k = flt_uniq ? opt->min_asm_ovlp : opt->ec_k;
    if (bo.k <= 0) {
        if (kcov_out) *kcov_out = 255.0f;
        if (len_out) for (int64_t i = 0; i < n; ++i) len_out[i] = (int32_t)(off[i + 1] - off[i]);
        return 0;
    }
    uint64_t tot_len = n > 0 ? (uint64_t)off[n] : 0;
    bo.l_pre = fml_initial_l_pre(tot_len);
    HostTable T;
    build_table(T, n, seqs, quals, off, bo.k, bo.q, bo.l_pre);
    bo.l_pre = T.l_pre;
    if (hist_out) { memcpy(hist_out, T.hist, 256 * 8); memcpy(hist_out + 256, T.hist_high, 64 * 8); }
    int mode = fml_hist_mode(T.hist);
    float kcov; int min_cov;
    fml_kcov_min_cov(T.hist, opt->min_cnt, opt->max_cnt, kcov, min_cov);
    bo.min_cov = min_cov;
    if (kcov_out) *kcov_out = kcov;
    CountTable tab = T.view();
    int64_t maxlen = 1;
    for (int64_t i = 0; i < n; ++i) maxlen = std::max<int64_t>(maxlen, off[i + 1] - off[i]);
    const int heap_cap = bo.max_heap + 8, stack_cap = 1 << 22;
    std::vector<u8> scratch(ec_scratch_bytes((int)maxlen, heap_cap, stack_cap) + 16);
    u8 *sp = (u8 *)(((uintptr_t)scratch.data() + 15) & ~(uintptr_t)15);
    int64_t max_stack = 0, max_heap = 0;
    for (int64_t i = 0; i < n; ++i) {
        int len = (int)(off[i + 1] - off[i]);
        char *s = seqs + off[i], *q = quals ? quals + off[i] : nullptr;
        if (flt_uniq) len_out[i] = len > 0 ? fltuniq1(bo, tab, s, q, len) : 0;
        else {
            EcScratch e;
            ec_scratch_bind(e, sp, (int)maxlen, heap_cap, stack_cap);
            int code = ec1(bo, tab, mode, s, q, len, e);
            if (codes) codes[i] = (uint8_t)code;
            max_stack = std::max<int64_t>(max_stack, e.stack_hw);
            max_heap = std::max<int64_t>(max_heap, e.heap_hw);
            if (len_out) len_out[i] = len;
        }
    }
    if (stat) { stat[0] = max_stack; stat[1] = max_heap; stat[2] = 0; }
    return 0;
}

// worker_count + bfc_ch_hist for an explicit (k, q, l_pre): hist[0..255] totals, hist[256..319] high counts
int fml_emul_count_hist(int64_t n, const char *seqs, const char *quals, const int64_t *off, int k, int q, int l_pre,
                        uint64_t *hist, int64_t *n_distinct)
{
    HostTable T;
    build_table(T, n, seqs, quals, off, k, q, l_pre);
    memcpy(hist, T.hist, 256 * 8); memcpy(hist + 256, T.hist_high, 64 * 8);
    if (n_distinct) *n_distinct = (int64_t)T.n_distinct;
    return 0;
}

} // extern "C"
