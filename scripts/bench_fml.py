"""Config-4 shaped measurement of the BFC stage (k-mer count + error correction + unique filter) on one B200:
N x 150 bp synthetic reads drawn from a 1 Mb random region, timed through the C ABI with host buffers
(e2e, copies included) and per stage on the device (CUDA events); optionally the reference's fml_correct / fml_fltuniq on
the host cores for a bounded sample.  Prints one JSON line; not the headline bench (that is bench.py)."""
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
    ap.add_argument("--region", type=int, default=1000000)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--cpu-sample", type=int, default=0, help="reads given to the reference library (0 = skip)")
    ap.add_argument("--check", type=int, default=0, help="compare the first N reads with the reference library")
    a = ap.parse_args()
    from seqlib_b200 import capi, synth
    import fmlcases
    capi.set_device(0)
    pac = synth.reference(a.region, seed=0x5EED0005)
    ctg = synth.contigs_for(a.region, 1, "asm")
    seqs, off, _, _ = synth.reads(pac, a.region, ctg, a.reads, 150, 0.01, 0.0, seed=0x5EED0006)
    quals = np.full(len(seqs), ord("I"), dtype=np.uint8)
    opt = capi.fml_default_opt()
    opt.ec_k = fmlcases.adjusted_ec_k(int(off[-1]), 0)
    out = {"workload": "%d x 150bp reads from a %d bp random region, 1%% substitutions, ec_k=%d" % (a.reads, a.region, opt.ec_k)}
    for name, flt in (("correct", False), ("fltuniq", True)):
        ts, sts = [], []
        for it in range(a.warmup + a.steps):
            t0 = time.perf_counter()
            r = capi.fml_correct_flat(opt, seqs, quals, off, flt_uniq=flt)
            t1 = time.perf_counter()
            if it >= a.warmup:
                ts.append(t1 - t0)
                sts.append(capi.fml_last_stats())
        st = sts[-1]
        ms = {k: float(np.mean([s[k] for s in sts])) for k in ("ms_count", "ms_ec", "ms_flt", "ms_total")}
        out[name] = {"e2e_reads_per_s": a.reads / float(np.mean(ts)), "device_ms": ms, "kcov": r[3],
                     "n_kmers": st["n_kmers"], "n_distinct": st["n_distinct"], "table_bytes": st["table_bytes"],
                     "lookups": st["n_lookups"], "spill": st["n_spill"], "ec_codes": st["ec_codes"], "launches": st["n_launches"],
                     "changed_bases": int((r[0] != seqs).sum()) if not flt else None,
                     "dropped": int((r[2] == 0).sum()) if flt else None}
        stage_ms = ms["ms_flt"] if flt else ms["ms_ec"]
        if stage_ms > 0:
            out[name]["lookup_GBps_32B_sectors"] = st["n_lookups"] * 32 / stage_ms / 1e6
        if not flt:
            corrected = r
    if a.check or a.cpu_sample:
        from oracle import pyref_fml
        if pyref_fml.have_ref():
            ro = pyref_fml.default_opt()
            ro.ec_k = opt.ec_k
            if a.cpu_sample:
                m = min(a.cpu_sample, a.reads)
                # a prefix of the same reads keeps the coverage of the sample proportional; timing only
                _, _, _, _, sec = pyref_fml.correct_flat(ro, seqs[:off[m]], quals[:off[m]], off[:m + 1])
                out["cpu_reference"] = {"reads": m, "seconds": sec, "reads_per_s": m / sec, "cores": 1,
                                        "what": "fml_correct (count + bfc_ec1), n_threads=1, on a prefix of the reads"}
            if a.check:
                rr = pyref_fml.correct_flat(ro, seqs, quals, off)
                m = min(a.check, a.reads)
                out["check"] = {"reads": a.reads, "seq_equal": bool(np.array_equal(rr[0], corrected[0])),
                                "qual_equal": bool(np.array_equal(rr[1], corrected[1])), "kcov_equal": bool(rr[3] == corrected[3]),
                                "ref_seconds": rr[4]}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
