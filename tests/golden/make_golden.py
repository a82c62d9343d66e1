"""Generates the committed parity fixtures from the reference itself (run in the build container,
where /root/reference and oracle/_ref exist).  Inputs copied from the reference's own test data are
data files, not code: the pre-built bwa index of tests/data/tiny.fa and subsets of its wgsim reads.
Expected outputs come from oracle/_ref/libseqref_bwa.so = the reference's unmodified bwa C.

    python tests/golden/make_golden.py
"""
import os
import shutil
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import pyref          # noqa: E402
from seqlib_b200 import synth     # noqa: E402
import parity                     # noqa: E402
import cases                      # noqa: E402

REF = os.environ.get("SEQLIB_REF", "/root/reference")


def save_results(path, res, **extra):
    np.savez_compressed(path, hit_off=res.hit_off, hits=res.hits, cigar=res.cigar,
                        md=np.frombuffer(res.md, dtype=np.uint8), **extra)


def main():
    opt = pyref.default_opt()
    # 1. tiny.fa index (the reference ships it pre-built) and read subsets
    os.makedirs(os.path.join(HERE, "tiny"), exist_ok=True)
    for ext in ("bwt", "sa", "pac", "ann", "amb"):
        shutil.copy(os.path.join(REF, "tests/data/tiny.fa." + ext), os.path.join(HERE, "tiny", "tiny.fa." + ext))
    tidx = pyref.RefIndex.load(os.path.join(HERE, "tiny", "tiny.fa"))
    for fq, n, out in (("sim1.fq", 5000, "sim1_5k"), ("sim1_bcr.fq", 2000, "bcr_2k")):
        _, rs, _ = parity.read_fastq(os.path.join(REF, "tests/data", fq), n)
        with open(os.path.join(HERE, out + ".txt"), "w") as f:
            f.write("\n".join(rs) + "\n")
        ids = cases.ids_for(len(rs))
        res, _ = pyref.align(tidx, rs, opt, ids)
        ioff, intv = pyref.collect_intv(tidx, rs, opt)
        save_results(os.path.join(HERE, out + ".npz"), res, intv_off=ioff, intv=intv)
    # 2. the reference's only known-answer test (seq_test/seq_test.cpp:848-911)
    names, seqs = cases.KAT_NAMES, cases.KAT_SEQS
    pyref.srand48(cases.KAT_SRAND)
    kidx = pyref.RefIndex.construct(names, seqs)
    a = kidx.arrays()
    ids = np.array([pyref.lib().refdrv_lrand48() for _ in cases.KAT_QUERIES], dtype=np.int64)
    res, _ = pyref.align(kidx, cases.KAT_QUERIES, opt, ids)
    save_results(os.path.join(HERE, "kat.npz"), res, ids=ids, bwt=a["bwt"], sa=a["sa"], pac=a["pac"], primary=a["primary"],
                 L2=np.array(a["L2"], dtype=np.uint64))
    # 3. config 1: 1k x 150 bp synthetic reads vs a 10 kb in-memory ConstructIndex
    pac, ctg, ref_ascii = cases.c1_reference()
    cidx = pyref.RefIndex.construct(["ref10k"], [ref_ascii])
    a = cidx.arrays()
    seqs_a, off = cases.c1_reads(pac, ctg)
    ids = cases.ids_for(len(off) - 1)
    res, _ = pyref.align(cidx, (seqs_a, off), opt, ids)
    save_results(os.path.join(HERE, "c1.npz"), res, bwt=a["bwt"], sa=a["sa"], primary=a["primary"], L2=np.array(a["L2"], dtype=np.uint64))
    # 4. ksw_extend2 tuples (config-3 shaped, 4000 of them)
    jobs, qp, tp = cases.c3_tuples(4000)
    out, _ = pyref.ksw_extend2_batch(jobs, qp, tp, np.array(list(opt.mat), dtype=np.int8))
    np.savez_compressed(os.path.join(HERE, "ksw_c3.npz"), out=out)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
