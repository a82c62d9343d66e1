"""End to end on the GPU: FASTQ text -> device line parser -> b200_mem_align_batch -> b200_results_to_sam equals the committed
text of the reference's mem_reg2sam (tests/golden/sam_*.sam, made by tests/golden/make_golden_sam.py)."""
import numpy as np
import pytest

import cases
import goldenlib
import test_cpu_sam

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,lo,hi", test_cpu_sam.GOLDEN_SLICES)
def test_fastq_to_sam_on_gpu(name, lo, hi):
    from seqlib_b200 import capi, fastq, sam
    capi.set_device(0)
    reads, names, quals, comments = test_cpu_sam._inputs(name, lo, hi)
    text = b"".join(b"@" + n + (b" " + c if c else b"") + b"\n" + (r.encode() if isinstance(r, str) else r) + b"\n+\n" + q + b"\n"
                    for n, c, r, q in zip(names, comments, reads, quals))
    rd = fastq.FastqReader(text=b"")
    b = rd.parse_device(text)
    assert b.n == hi - lo and b.parsed_on_device == 1
    idx = capi.Index.load(goldenlib.path("tiny", "tiny.fa"))
    opt = capi.default_opt()
    ids = cases.ids_for(5000)[lo:hi]
    res = capi.align(idx, (b.seq, b.seq_off), opt, ids)
    rnames = [idx.seq_name(i) for i in range(idx.n_seqs())]
    recs = b.records()
    got = sam.results_to_sam(res, opt, rnames, b.seq, b.seq_off, [r[0] for r in recs], [r[3] for r in recs], [r[1] for r in recs])
    assert got == open(goldenlib.path("sam_%s_%d_%d.sam" % (name, lo, hi)), "rb").read()
    rd.close()
