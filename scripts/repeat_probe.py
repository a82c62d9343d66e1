"""Throughput on a reference with a human-like interspersed repeat family (FRAC of the sequence = copies of a 300-bp element at DIV
divergence) against the same size of random reference: device-resident run, stage times, spill counts."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from seqlib_b200 import capi, synth
capi.set_device(0)
L = int(os.environ.get("REF", 400_000_000)); n = int(os.environ.get("READS", 4_000_000))
FRAC = float(os.environ.get("FRAC", 0.10)); DIV = float(os.environ.get("DIV", 0.12))
rng = np.random.Generator(np.random.PCG64(5))
ref = rng.integers(0, 4, L, dtype=np.uint8)
el = 300
n_copy = int(L * FRAC / el)
if n_copy:
    elem = rng.integers(0, 4, el, dtype=np.uint8)
    pos = np.sort(rng.choice(L // el - 1, n_copy, replace=False)) * el
    cop = np.tile(elem, n_copy).reshape(n_copy, el)
    mut = rng.random((n_copy, el)) < DIV
    cop[mut] = (cop[mut] + rng.integers(1, 4, int(mut.sum()), dtype=np.uint8)) & 3
    idxs = (pos[:, None] + np.arange(el)[None, :]).ravel()
    ref[idxs] = cop.ravel()
pac = np.zeros((L + 3) // 4 + 1, np.uint8)
r4 = np.concatenate([ref, np.zeros((-L) % 4, np.uint8)]).reshape(-1, 4)
pac[: len(r4)] = (r4[:, 0] << 6 | r4[:, 1] << 4 | r4[:, 2] << 2 | r4[:, 3]).astype(np.uint8)
ctg = synth.contigs_for(L, 4)
seqs, off, _, _ = synth.reads(pac, L, ctg, n, 150, 0.01, 2e-4)
idx = capi.Index.construct_pac(pac, L, ctg, keep_host=False)
opt = capi.default_opt()
ids = np.arange(n, dtype=np.int64) * 7919 + 13
b = capi.Batch(idx, (seqs, off), opt, ids)
for it in range(3):
    t0 = time.perf_counter(); b.run(); dt = time.perf_counter() - t0
    st = capi.last_stats()
    print("FRAC %.2f DIV %.2f run %d: %.1f ms (%.2f M reads/s)" % (FRAC, DIV, it, 1e3 * dt, n / dt / 1e6),
          {k: round(st[k], 1) for k in ("ms_seed", "ms_chain", "ms_extend", "ms_finalize")}, "overflow", st["n_overflow"], "fallback", st["ext_fallback"],
          "occ/read %.0f chain/read %.0f" % (st["occ_blocks"] / n, st["tab_lookups_hi"] / n), flush=True)
