"""ctypes binding of tests/hostsim/libfmlemul.so: the BFC per-read device code (seqlib_b200/csrc/bfc.cuh) compiled for the
host (test harness only, lets the CPU suite check kernel logic without a GPU)."""
import ctypes as C
import os
import subprocess
import numpy as np
from seqlib_b200.abi import FmlOpt

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "hostsim", "libfmlemul.so")
_lib = None


def build():
    src = os.path.join(_HERE, "hostsim", "fml_emul.cpp")
    csrc = os.path.join(_HERE, "..", "seqlib_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in ("bfc.cuh", "common.cuh", "fml_host.h")]
    if os.path.exists(_SO) and all(os.path.getmtime(_SO) >= os.path.getmtime(d) for d in deps):
        return
    subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-o", _SO, src])


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.fml_emul_correct_flat.argtypes = [C.POINTER(FmlOpt), C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.POINTER(C.c_float), C.c_void_p, C.c_void_p, C.c_void_p]
        L.fml_emul_count_hist.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                          C.POINTER(C.c_int64)]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def correct_flat(opt, seqs, quals, off, flt_uniq=False):
    seqs = np.array(seqs, dtype=np.uint8, copy=True)
    quals = None if quals is None else np.array(quals, dtype=np.uint8, copy=True)
    off = np.ascontiguousarray(off, dtype=np.int64)
    n = len(off) - 1
    lens = np.zeros(max(n, 1), dtype=np.int32)
    kcov = C.c_float(0)
    hist = np.zeros(320, dtype=np.uint64)
    stat = np.zeros(3, dtype=np.int64)
    codes = np.zeros(max(n, 1), dtype=np.uint8)
    lib().fml_emul_correct_flat(C.byref(opt), int(bool(flt_uniq)), n, _p(seqs), _p(quals), _p(off), _p(lens), C.byref(kcov),
                                _p(hist), _p(stat), _p(codes))
    return seqs, quals, lens[:n], kcov.value, dict(hist=hist, max_stack=int(stat[0]), max_heap=int(stat[1]), codes=codes[:n])


def count_hist(seqs, quals, off, k, q=20, l_pre=20):
    seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
    quals = None if quals is None else np.ascontiguousarray(quals, dtype=np.uint8)
    off = np.ascontiguousarray(off, dtype=np.int64)
    h = np.zeros(320, dtype=np.uint64)
    nd = C.c_int64(0)
    lib().fml_emul_count_hist(len(off) - 1, _p(seqs), _p(quals), _p(off), k, q, l_pre, _p(h), C.byref(nd))
    cnt = h[:256].copy()
    mode = -1
    mx = 0
    for i in range(3, 256):
        if cnt[i] > mx:
            mx, mode = cnt[i], i
    return cnt, h[256:].copy(), mode, nd.value
