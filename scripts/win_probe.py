import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from seqlib_b200 import capi, synth
capi.set_device(0)
n, nw = 1000, 64
region = n
ws = []
for w in range(nw):
    pac = synth.reference(region, seed=0x5EED0100 + w)
    s_, o_, _, _ = synth.reads(pac, region, synth.contigs_for(region, 1, "w"), n, 150, 0.01, 0.0, seed=0x5EED0200 + w)
    ws.append((s_, o_))
seqs = np.concatenate([w[0] for w in ws]); quals = np.full(len(seqs), ord("I"), dtype=np.uint8)
off = np.concatenate([[0]] + [w[1][1:] + i * int(ws[0][1][-1]) for i, w in enumerate(ws)]).astype(np.int64)
win_off = np.arange(nw + 1, dtype=np.int64) * n
opt = capi.fml_default_opt()
nt = int(sys.argv[1])
capi.fml_assemble_windows(opt, seqs, quals, off, win_off, nt)
t0 = time.perf_counter(); capi.fml_assemble_windows(opt, seqs, quals, off, win_off, nt); dt = time.perf_counter() - t0
st = capi.fml_last_stats()
print("threads", nt, "ms/window", round(1e3 * dt / nw, 2), "per-window stage ms:", {k: round(st[k] / nw, 2) for k in ("ms_count", "ms_ec", "ms_flt", "ms_fmd", "ms_nodes", "ms_walk_host", "ms_clean_host", "ms_total")})
