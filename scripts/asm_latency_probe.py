"""Latency of one FermiAssembler-sized problem (local assembly of a window: 10^3..10^5 reads) through b200_fml_assemble_flat,
next to the reference's fml_assemble on one host core."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from seqlib_b200 import capi, synth
from oracle import pyref_fml
capi.set_device(0)
for n in (1000, 5000, 20000, 100000):
    region = max(400, n)            # 150x
    pac = synth.reference(region, seed=0x5EED0005)
    seqs, off, _, _ = synth.reads(pac, region, synth.contigs_for(region, 1, "w"), n, 150, 0.01, 0.0, seed=0x5EED0006)
    quals = np.full(len(seqs), ord("I"), dtype=np.uint8)
    opt = capi.fml_default_opt()
    capi.fml_assemble_flat(opt, seqs, quals, off)
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); u = capi.fml_assemble_flat(opt, seqs, quals, off); ts.append(time.perf_counter() - t0)
    st = capi.fml_last_stats()
    print("   stages:", {k: round(st[k], 2) for k in ("ms_count", "ms_ec", "ms_flt", "ms_fmd", "ms_nodes", "ms_walk_host", "ms_clean_host", "ms_total")}, flush=True)
    line = "n=%6d: %.1f ms per assembly (%d unitigs, longest %d, %d launches)" % (n, 1e3 * np.median(ts), len(u), max([len(x["seq"]) for x in u] + [0]), st["n_launches"])
    if pyref_fml.have_ref() and n <= 20000:
        exp, sec = pyref_fml.assemble(pyref_fml.default_opt(), seqs, quals, off)
        line += "; reference %.1f ms, identical=%s" % (1e3 * sec, [x["seq"] for x in exp] == [x["seq"] for x in u])
    print(line, flush=True)

# many windows per call: b200_fml_assemble_windows against the same windows one call at a time
for n, nw in ((1000, 128), (5000, 64)):
    region = max(400, n)
    ws = []
    for w in range(nw):
        pac = synth.reference(region, seed=0x5EED0100 + w)
        s_, o_, _, _ = synth.reads(pac, region, synth.contigs_for(region, 1, "w"), n, 150, 0.01, 0.0, seed=0x5EED0200 + w)
        ws.append((s_, o_))
    seqs = np.concatenate([w[0] for w in ws]); quals = np.full(len(seqs), ord("I"), dtype=np.uint8)
    off = np.concatenate([[0]] + [w[1][1:] + i * int(ws[0][1][-1]) for i, w in enumerate(ws)]).astype(np.int64)
    win_off = np.arange(nw + 1, dtype=np.int64) * n
    opt = capi.fml_default_opt()
    t0 = time.perf_counter()
    one = [capi.fml_assemble_flat(opt, w[0], quals[:len(w[0])], w[1]) for w in ws]
    t_seq = time.perf_counter() - t0
    for nt in (1, 4, 8, 16):
        capi.fml_assemble_windows(opt, seqs, quals, off, win_off, nt)
        t0 = time.perf_counter(); got = capi.fml_assemble_windows(opt, seqs, quals, off, win_off, nt); dt = time.perf_counter() - t0
        same = all([x["seq"] for x in a] == [x["seq"] for x in b] for a, b in zip(one, got))
        print("windows: %d x %d reads, %2d threads: %.1f ms per window (%.0f windows/s; one call at a time %.1f ms per window), identical=%s"
              % (nw, n, nt, 1e3 * dt / nw, nw / dt, 1e3 * t_seq / nw, same), flush=True)
