"""SAM text: b200_results_to_sam is host-side formatting of the aligner's hits, so its parity with the
reference's mem_reg2sam (oracle/_ref, refdrv_sam) is checked on the CPU: the reference's own regions go in, the text must be
the reference's text -- for SeqLib's option set (soft clipping), for bwa's defaults (hard-clipped supplementary records),
with secondary records (MEM_F_ALL), XB instead of XA, a threshold nothing reaches (unaligned records), with and without
qualities and comments -- and against the committed text (tests/golden/sam_bcr_200.*.sam)."""
import os

import numpy as np
import pytest

import cases
import goldenlib

F_ALL, F_SOFTCLIP, F_XB, F_NO_MULTI, F_KEEP_SUPP_MAPQ = 0x8, 0x200, 0x2000, 0x10, 0x1000
TINY_RNAMES = None


GOLDEN_SLICES = (("bcr_2k", 900, 1100), ("sim1_5k", 800, 900))      # split reads (SA tags, hard clips) / multi-hit reads (XA tags)


def _inputs(name, lo, hi):
    reads = cases.read_lines(goldenlib.path(name + ".txt"))[lo:hi]
    names = [b"read%d/%s" % (lo + i, name.encode()) for i in range(len(reads))]
    rng = np.random.default_rng(5)
    quals = [bytes(rng.integers(33, 74, size=len(r), dtype=np.uint8)) for r in reads]
    comments = [b"BC:Z:%d" % i if i % 3 == 0 else b"" for i in range(len(reads))]
    return reads, names, quals, comments


def _variants(opt_factory):
    out = []
    for label, setf in (("seqlib", lambda o: None),
                        ("bwa_default", lambda o: setattr(o, "flag", o.flag & ~F_SOFTCLIP)),
                        ("all", lambda o: setattr(o, "flag", (o.flag & ~F_SOFTCLIP) | F_ALL)),
                        ("xb_nomulti", lambda o: setattr(o, "flag", (o.flag & ~F_SOFTCLIP) | F_XB | F_NO_MULTI | F_KEEP_SUPP_MAPQ)),
                        ("high_T", lambda o: setattr(o, "T", 140))):
        o = opt_factory()
        setf(o)
        out.append((label, o))
    return out


def test_sam_text_vs_live_reference():
    from oracle import pyref
    if not pyref.have_ref():
        pytest.skip("oracle/_ref not built")
    from seqlib_b200 import sam
    from seqlib_b200.abi import pack_reads
    tidx = pyref.RefIndex.load(goldenlib.path("tiny", "tiny.fa"))
    rnames = [l.split()[1] for l in open(goldenlib.path("tiny", "tiny.fa.ann")).read().splitlines()[1::2]]
    seen = set()
    for name, n in (("bcr_2k", 2000), ("sim1_5k", 5000)):
        reads, names, quals, comments = _inputs(name, 0, n)
        seqs, off = pack_reads(reads)
        ids = cases.ids_for(len(reads))
        for label, opt in _variants(pyref.default_opt):
            res, _ = pyref.align(tidx, (seqs, off), opt, ids)
            for q, c in ((quals, comments), (None, None), (quals, None)):
                exp = pyref.sam(tidx, (seqs, off), opt, ids, names, q, c)
                got = sam.results_to_sam(res, opt, rnames, seqs, off, names, q, c)
                assert got == exp, (name, label, q is not None, c is not None)
                for tag in (b"SA:Z:", b"XA:Z:", b"XB:Z:", b"\t2048\t", b"\t256\t", b"\t4\t*\t0\t0"):
                    if tag in exp:
                        seen.add(tag)
    assert len(seen) == 6, seen          # the inputs exercise split reads, multi-hit reads, secondary and unaligned records


@pytest.mark.parametrize("name,lo,hi", GOLDEN_SLICES)
def test_sam_text_vs_golden(name, lo, hi):
    """the committed reference text (tests/golden/sam_<name>_<lo>_<hi>.sam, SeqLib's option set) from the golden hits"""
    from seqlib_b200 import sam, capi
    from seqlib_b200.abi import pack_reads
    gold, _ = goldenlib.load(name)
    reads, names, quals, comments = _inputs(name, lo, hi)
    seqs, off = pack_reads(reads)
    opt = capi.default_opt()

    class Sub:
        pass
    r = Sub()
    r.hit_off = gold.hit_off[lo:hi + 1] - gold.hit_off[lo]
    r.hits = gold.hits[int(gold.hit_off[lo]):int(gold.hit_off[hi])]
    r.cigar = gold.cigar; r.md = gold.md
    rnames = [l.split()[1] for l in open(goldenlib.path("tiny", "tiny.fa.ann")).read().splitlines()[1::2]]
    got = sam.results_to_sam(r, opt, rnames, seqs, off, names, quals, comments)
    assert got == open(goldenlib.path("sam_%s_%d_%d.sam" % (name, lo, hi)), "rb").read()
