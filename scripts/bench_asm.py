"""Config-4 shaped measurement of FermiAssembler::PerformAssembly's compute (b200_fml_assemble_flat) on one B200:
N x 150 bp synthetic reads from a random region at 150x coverage, per-stage device / host times, and optionally
the reference's fml_assemble on the host cores for a bounded sample.  Prints one JSON line; not the headline bench."""
import argparse
import json
import os
import sys
import time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=1000000)
    ap.add_argument("--region", type=int, default=0, help="0: reads * 150 / 150 (150x coverage)")
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--ref-reads", type=int, default=0, help="run the reference's fml_assemble on a same-coverage problem of this many reads")
    ap.add_argument("--check", action="store_true", help="compare the unitigs with the reference on the --ref-reads problem")
    a = ap.parse_args()
    from seqlib_b200 import capi, synth
    import fmlcases
    capi.set_device(0)

    def make(n):
        region = a.region or n
        pac = synth.reference(region, seed=0x5EED0005)
        ctg = synth.contigs_for(region, 1, "asm")
        seqs, off, _, _ = synth.reads(pac, region, ctg, n, 150, 0.01, 0.0, seed=0x5EED0006)
        return seqs, np.full(len(seqs), ord("I"), dtype=np.uint8), off, region

    seqs, quals, off, region = make(a.reads)
    opt = capi.fml_default_opt()
    out = {"workload": "%d x 150bp reads from a %d bp random region (150x), 1%% substitutions" % (a.reads, region)}
    ts, sts = [], []
    for it in range(a.warmup + a.steps):
        t0 = time.perf_counter()
        utgs = capi.fml_assemble_flat(opt, seqs, quals, off)
        t1 = time.perf_counter()
        if it >= a.warmup:
            ts.append(t1 - t0)
            sts.append(capi.fml_last_stats())
    keys = ("ms_count", "ms_ec", "ms_flt", "ms_fmd", "ms_nodes", "ms_walk_host", "ms_clean_host", "ms_total")
    st = sts[-1]
    out["assemble"] = {"e2e_seconds": float(np.mean(ts)), "e2e_reads_per_s": a.reads / float(np.mean(ts)),
                       "stage_ms": {k: float(np.mean([s[k] for s in sts])) for k in keys},
                       "n_utg": len(utgs), "longest": max([len(u["seq"]) for u in utgs] + [0]),
                       "fmd_symbols": st["fmd_symbols"], "n_strings": st["n_strings"], "n_vertices": st["n_vertices"],
                       "launches": st["n_launches"], "spill": st["n_spill"]}
    if a.ref_reads:
        from oracle import pyref_fml
        if pyref_fml.have_ref():
            s2, q2, o2, r2 = make(a.ref_reads)
            exp, sec = pyref_fml.assemble(pyref_fml.default_opt(), s2, q2, o2)
            out["cpu_reference"] = {"reads": a.ref_reads, "seconds": sec, "reads_per_s": a.ref_reads / sec, "cores": 1,
                                    "n_utg": len(exp), "longest": max([len(u["seq"]) for u in exp] + [0]),
                                    "what": "fml_assemble, n_threads=1, same coverage on a %d bp region" % r2}
            if a.check:
                got = capi.fml_assemble_flat(capi.fml_default_opt(), s2, q2, o2)
                out["check"] = {"reads": a.ref_reads, "identical_unitigs": fmlcases.utg_text(got) == fmlcases.utg_text(exp)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
