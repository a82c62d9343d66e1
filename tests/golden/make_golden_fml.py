"""Generates the committed parity fixtures of the fermi-lite half from the reference itself (run in the build container,
where /root/reference and oracle/_ref exist).  Expected outputs come from oracle/_ref/libseqref_fml.so = the reference's
unmodified fermi-lite C.  Inputs: the reference's own test reads (tests/data/sim1_bcr.fq, fermi-lite/test/MT-simu.fq.gz;
data, not code) and the seeded synthetic set of tests/cases.py.

    python tests/golden/make_golden_fml.py
"""
import gzip
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import pyref_fml                    # noqa: E402
from seqlib_b200.abi import pack_reads_quals    # noqa: E402
import cases                                    # noqa: E402
import fmlcases                                 # noqa: E402

REF = os.environ.get("SEQLIB_REF", "/root/reference")


def read_fq(path, n):
    op = gzip.open if path.endswith(".gz") else open
    with op(path, "rt") as f:
        L = f.read().split("\n")
    return L[1::4][:n], L[3::4][:n]


def main():
    sets = {}
    s, q = read_fq(os.path.join(REF, "tests/data/sim1_bcr.fq"), 2000)
    assert s == cases.read_lines(os.path.join(HERE, "bcr_2k.txt"))
    sets["fml_bcr_2k"] = pack_reads_quals(s, q)
    s, q = read_fq(os.path.join(REF, "fermi-lite/test/MT-simu.fq.gz"), 2000)
    sets["fml_mt_2k"] = pack_reads_quals(s, q)
    sets["fml_mixed"] = cases.fml_reads(3000, region=5000, seed=1)
    sq, _, so = cases.fml_reads(1500, region=3000, seed=3)
    sets["fml_mixed_noqual"] = (sq, None, so)
    for name, (seqs, quals, off) in sets.items():
        exp = fmlcases.reference_pipeline(pyref_fml, seqs, quals, off)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), seqs=seqs, quals=quals if quals is not None else np.zeros(0, np.uint8),
                            off=off, **exp)
        print(name, "ec_k", int(exp["ec_k"]), "kcov", float(exp["ec_kcov"]), float(exp["flt_kcov"]),
              "changed", int((exp["ec_seqs"] != seqs).sum()), "dropped", int((exp["flt_lens"] == 0).sum()))


if __name__ == "__main__":
    main()
