"""ctypes binding of tests/hostsim/libfmdemul.so: FMD rank code, per-string overlap records, seed-order walk and graph cleaning
compiled for the host over a BWT supplied by the test (test harness only)."""
import ctypes as C
import os
import subprocess
import numpy as np
from seqlib_b200.abi import FmlOpt

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "hostsim", "libfmdemul.so")
_lib = None


def build():
    src = os.path.join(_HERE, "hostsim", "fmd_emul.cpp")
    csrc = os.path.join(_HERE, "..", "seqlib_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in ("fmd.cuh", "unitig.cuh", "utg_walk.h", "mag_host.h", "sort.cuh", "common.cuh")]
    if os.path.exists(_SO) and all(os.path.getmtime(_SO) >= os.path.getmtime(d) for d in deps):
        return
    subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-w", "-o", _SO, src])


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.fmd_emul_rank.argtypes = [C.c_void_p, C.c_uint64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.fmd_emul_mag_text.restype = C.c_void_p
        L.fmd_emul_mag_text.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(FmlOpt), C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_float), C.c_void_p]
        L.fmd_emul_free.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def rank(bwt, q):
    bwt = np.ascontiguousarray(bwt, dtype=np.uint8)
    q = np.ascontiguousarray(q, dtype=np.uint64)
    ranks = np.zeros((max(len(q), 1), 6), dtype=np.uint64)
    sym = np.zeros(max(len(q), 1), dtype=np.int32)
    lib().fmd_emul_rank(bwt.ctypes.data, len(bwt), len(q), q.ctypes.data, ranks.ctypes.data, sym.ctypes.data)
    return ranks[:len(q)], sym[:len(q)]


def mag_text(bwt, opt, stage):
    bwt = np.ascontiguousarray(bwt, dtype=np.uint8)
    ln = C.c_int64(0)
    rd = C.c_float(0)
    stat = np.zeros(4, dtype=np.int64)
    p = lib().fmd_emul_mag_text(bwt.ctypes.data, len(bwt), C.byref(opt), stage, C.byref(ln), C.byref(rd), stat.ctypes.data)
    txt = C.string_at(p, ln.value).decode()
    lib().fmd_emul_free(p)
    return txt, rd.value, stat
