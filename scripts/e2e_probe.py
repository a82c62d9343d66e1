"""Where the end-to-end time of b200_mem_align_batch goes: N reads vs a small random reference, B200_TRACE phases."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from seqlib_b200 import capi, synth
capi.set_device(0)
L = int(os.environ.get("REF", 200_000_000)); n = int(os.environ.get("READS", 4_000_000))
pac = synth.reference(L); ctg = synth.contigs_for(L, 4)
seqs, off, _, _ = synth.reads(pac, L, ctg, n, 150, 0.01, 0.0)
idx = capi.Index.construct_pac(pac, L, ctg, keep_host=False)
opt = capi.default_opt()
ids = np.arange(n, dtype=np.int64) * 7919 + 13
ps, po, pi = [torch.from_numpy(a).pin_memory().numpy() for a in (seqs, off, ids)]
for it in range(4):
    t0 = time.perf_counter()
    h = capi.align_raw(idx, ps, po, opt, pi)
    t1 = time.perf_counter()
    capi.results_free(h)
    print("iter %d: %.1f ms total (%.2f M reads/s), free %.1f ms" % (it, 1e3 * (t1 - t0), n / (t1 - t0) / 1e6, 1e3 * (time.perf_counter() - t1)), flush=True)
b = capi.Batch(idx, (seqs, off), opt, ids)
for it in range(3):
    b.run(); st = capi.last_stats()
    print("device-resident run: %.1f ms" % st["ms_total"], {k: round(st[k], 1) for k in ("ms_seed", "ms_chain", "ms_extend", "ms_finalize")})
