"""Latency of the single-read call (what BWAAligner::alignSequence does per read) and of small batches."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from seqlib_b200 import capi
import cases, goldenlib
capi.set_device(0)
idx = capi.Index.load(goldenlib.path("tiny", "tiny.fa"))
reads = cases.read_lines(goldenlib.path("sim1_5k.txt"))
opt = capi.default_opt()
for bs in (1, 8, 64, 1024):
    n_calls = max(20, min(400, 4096 // bs))
    capi.align(idx, reads[:bs], opt, np.arange(bs, dtype=np.int64))
    t0 = time.perf_counter()
    for c in range(n_calls):
        capi.align(idx, reads[c % 4 * bs:(c % 4 + 1) * bs], opt, np.arange(bs, dtype=np.int64))
    dt = (time.perf_counter() - t0) / n_calls
    print("batch %5d: %.3f ms per call, %.0f reads/s" % (bs, dt * 1e3, bs / dt), flush=True)
