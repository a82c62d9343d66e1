"""Generates the committed parity fixtures of the assembly half (FMD-index, unitig graph, cleaned graph, unitigs) from the
reference's own fermi-lite C (oracle/_ref/libseqref_fml.so), for the read sets of make_golden_fml.py.

    python tests/golden/make_golden_asm.py
"""
import hashlib
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import pyref_fml      # noqa: E402
import fmlcases                   # noqa: E402


def main():
    for name in fmlcases.FML_SETS:
        seqs, quals, off, z = fmlcases.load(name)
        utgs, _ = pyref_fml.assemble(pyref_fml.default_opt(), seqs, quals, off)
        fs, foff = fmlcases.filtered_reads(z, off)
        kcov = float(z["flt_kcov"])
        rng = np.random.default_rng(7)
        bwt, cnt, mcnt, _ = pyref_fml.bwt(fs, foff)
        q = np.concatenate([rng.integers(0, len(bwt) + 1, 500).astype(np.uint64), np.array([0, 1, 127, 128, 129, len(bwt)], dtype=np.uint64)])
        _, _, _, (rr, rs) = pyref_fml.bwt(fs, foff, q)
        m0, rd, _ = pyref_fml.mag_text(pyref_fml.default_opt(), 0, kcov, fs, foff)
        m1, _, _ = pyref_fml.mag_text(pyref_fml.default_opt(), 1, kcov, fs, foff)
        np.savez_compressed(os.path.join(HERE, name + "_asm.npz"), bwt_md5=hashlib.md5(bwt.tobytes()).hexdigest(), cnt=cnt,
                            rank_q=q, rank_r=rr, rank_s=rs, rdist=np.float32(rd),
                            mag0=np.frombuffer(m0.encode(), dtype=np.uint8), mag1=np.frombuffer(m1.encode(), dtype=np.uint8),
                            utg=np.frombuffer(fmlcases.utg_text(utgs).encode(), dtype=np.uint8))
        print(name, "bwt", len(bwt), "vertices", m0.count("\n+\n"), "->", m1.count("\n+\n"), "unitigs", len(utgs),
              "longest", max([len(u["seq"]) for u in utgs] + [0]))


if __name__ == "__main__":
    main()
