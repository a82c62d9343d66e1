"""The wrapper layer: the product's record assembly (seqlib_b200/cxx/BWA.cpp) against the CPU restatement
of src/BWAAligner.cpp:111-248 (oracle/oracle_wrap.cpp), both fed the reference's own regions (committed golden vectors made by
the reference's bwa C).  No GPU: only the host-side C++ runs."""
import ctypes as C
import os
import subprocess
import numpy as np
import pytest

import cases
import goldenlib

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _lib():
    so = os.path.join(HERE, "cxx", "libwraptest.so")
    srcs = [os.path.join(HERE, "cxx", "wraptest.cpp"), os.path.join(ROOT, "oracle", "oracle_wrap.cpp")]
    deps = srcs + [os.path.join(ROOT, "seqlib_b200", "libSeqLibB200.so")]
    if not os.path.exists(so) or any(os.path.getmtime(so) < os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-I" + os.path.join(ROOT, "include"), "-o", so] + srcs +
                              ["-L" + os.path.join(ROOT, "seqlib_b200"), "-lSeqLibB200", "-lseqlib_b200",
                               "-Wl,-rpath," + os.path.join(ROOT, "seqlib_b200")])
    L = C.CDLL(so)
    for f in (L.wraptest_records, L.oracle_wrap_records):
        f.restype = C.c_int64
        f.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_void_p, C.c_int64,
                      C.POINTER(C.c_int)]
    return L


def _run(fn, seq, name, regs, cigar, hardclip, frac, maxsec):
    buf = np.zeros(1 << 16, dtype=np.uint8)
    n = C.c_int(0)
    sz = fn(seq, len(seq), name, len(regs), regs.ctypes.data_as(C.c_void_p), cigar.ctypes.data_as(C.c_void_p), int(hardclip), float(frac),
            int(maxsec), buf.ctypes.data_as(C.c_void_p), buf.size, C.byref(n))
    return n.value, (bytes(buf[:sz]) if 0 <= sz <= buf.size else None)


@pytest.mark.parametrize("name", ["bcr_2k", "sim1_5k"])
def test_wrapper_records_equal_restated_reference(name):
    if not os.path.exists(os.path.join(ROOT, "seqlib_b200", "libSeqLibB200.so")):
        pytest.skip("libSeqLibB200.so not built")
    L = _lib()
    gold, z = goldenlib.load(name)
    reads = cases.read_lines(goldenlib.path(name + ".txt"))
    cigar = np.ascontiguousarray(gold.cigar, dtype=np.uint32)
    n_multi = n_sec = n_rev_clip = n_records = 0
    for r in range(len(reads)):
        regs = np.ascontiguousarray(gold.hits[gold.hit_off[r]:gold.hit_off[r + 1]])
        if len(regs) == 0:
            continue
        n_multi += len(regs) > 1
        n_sec += int(((regs["flag"] & 256) != 0).sum())
        for h in regs:                                     # reverse-strand hits with clips of different lengths at the two ends
            cg = gold.cigar_of(h)
            if h["is_rev"] and len(cg) > 1 and (cg[0] & 0xf) == 3 and ((cg[-1] & 0xf) != 3 or (cg[0] >> 4) != (cg[-1] >> 4)):
                n_rev_clip += 1
        seq = reads[r].encode()
        combos = [(False, 0.9, 10), (True, 0.9, 10)] if len(regs) == 1 else \
            [(hc, f, m) for hc in (False, True) for f in (-1.0, 0.0, 0.5, 0.9, 1.0, 2.0) for m in (0, 1, 10)]
        for hardclip, frac, maxsec in combos:
            a = _run(L.wraptest_records, seq, b"read%d" % r, regs, cigar, hardclip, frac, maxsec)
            b = _run(L.oracle_wrap_records, seq, b"read%d" % r, regs, cigar, hardclip, frac, maxsec)
            assert a == b, "read %d hardclip %s keepSecFrac %s maxSecondary %d: %d vs %d records" % (r, hardclip, frac, maxsec, a[0], b[0])
            n_records += b[0]
    # the fixture must actually exercise the quirks
    assert n_multi > 20 and n_records > 2000
    if name == "bcr_2k":
        assert n_sec > 0 and n_rev_clip > 0
