"""Generates tests/golden/sam_<name>_<lo>_<hi>.sam: the reference's mem_reg2sam text (oracle/_ref, refdrv_sam, SeqLib's
option set) for the read slices test_cpu_sam.GOLDEN_SLICES of the golden read sets.  Run from the repo root: python tests/golden/make_golden_sam.py"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import cases
import goldenlib
import test_cpu_sam
from oracle import pyref
from seqlib_b200.abi import pack_reads

tidx = pyref.RefIndex.load(goldenlib.path("tiny", "tiny.fa"))
for name, lo, hi in test_cpu_sam.GOLDEN_SLICES:
    reads, names, quals, comments = test_cpu_sam._inputs(name, lo, hi)
    seqs, off = pack_reads(reads)
    ids = cases.ids_for(5000)[lo:hi]
    opt = pyref.default_opt()
    text = pyref.sam(tidx, (seqs, off), opt, ids, names, quals, comments)
    open(goldenlib.path("sam_%s_%d_%d.sam" % (name, lo, hi)), "wb").write(text)
    print(name, len(text), text.count(b"\n"), text.count(b"SA:Z:"), text.count(b"XA:Z:"))
