// fmd_dev.h -- host-side handle of a device-resident FMD-index (internal to libseqlib_b200.so).
#pragma once
#include "engine.cuh"
#include "fmd.cuh"

#define FMD_KEY_SYMS 27          // 5^27 < 2^63: symbols per sort-key chunk

namespace b200 {

struct FmdDevice {
    DevBuf start;    // u64[n_str + 1]: first row of string j in `text` (string 2r = read r, 2r + 1 = its reverse complement)
    DevBuf text;     // u8: symbols 1..4 of every indexed string, each followed by 0
    DevBuf blocks;   // FmdBlock[n_blk]
    DevBuf bwt8;     // one symbol per BWT position (kept for parity dumps)
    FmdIndex idx{};
    u64 n_blk = 0;
    void release() { start.release(); text.release(); blocks.release(); bwt8.release(); idx = FmdIndex{}; n_blk = 0; }
};

// fmd.cu: builds F from reads resident on the device (ASCII pool + offsets [+ per-read kept lengths])
extern thread_local int g_fmd_launches;
void fmd_build_device(FmdDevice &F, const char *d_seq, const i64 *d_off, const i32 *d_len, i64 n_reads, cudaStream_t st);

} // namespace b200
