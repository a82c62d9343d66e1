"""ctypes binding of tests/hostsim/libhostsim.so (CPU emulation of the per-read device functions; test harness only)."""
import ctypes as C
import os
import subprocess
import numpy as np
from seqlib_b200.abi import MemOpt, IndexView, ResultsView, Results, HIT_DTYPE, INTV_DTYPE, np_from_ptr, pack_reads

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "hostsim", "libhostsim.so")
_lib = None


def build():
    src = os.path.join(_HERE, "hostsim", "hostsim.cpp")
    csrc = os.path.join(_HERE, "..", "seqlib_b200", "csrc")
    deps = [src, os.path.join(_HERE, "hostsim", "hostindex.h")] + \
        [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cuh", ".h"))]
    if os.path.exists(_SO) and all(os.path.getmtime(_SO) >= os.path.getmtime(d) for d in deps):
        return
    subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-x", "c++", "-o", _SO, src])


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.hostsim_index_from_view.restype = C.c_void_p
        L.hostsim_index_from_view.argtypes = [C.POINTER(IndexView), C.c_int]
        L.hostsim_index_destroy.argtypes = [C.c_void_p]
        L.hostsim_align.restype = C.c_void_p
        L.hostsim_align.argtypes = [C.c_void_p, C.POINTER(MemOpt), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.hostsim_results_view.argtypes = [C.c_void_p, C.POINTER(ResultsView)]
        L.hostsim_ovf.restype = C.POINTER(C.c_uint32)
        L.hostsim_ovf.argtypes = [C.c_void_p]
        L.hostsim_stage_views.argtypes = [C.c_void_p] + [C.POINTER(C.c_void_p)] * 8
        L.hostsim_results_free.argtypes = [C.c_void_p]
        _lib = L
    return _lib


class SimIndex:
    def __init__(self, view, keep=None):
        self._keep = keep
        self.h = lib().hostsim_index_from_view(C.byref(view), 5)

    def __del__(self):
        if self.h:
            lib().hostsim_index_destroy(self.h)
            self.h = None


def align(idx, reads, opt, ids, small=False):
    seqs, off = pack_reads(reads) if not isinstance(reads, tuple) else reads
    n = len(off) - 1
    ids = np.ascontiguousarray(ids, dtype=np.int64)
    h = lib().hostsim_align(idx.h, C.byref(opt), n, seqs.ctypes.data, off.ctypes.data, ids.ctypes.data, int(small))
    v = ResultsView()
    lib().hostsim_results_view(h, C.byref(v))
    res = Results(v)
    res.ovf = np_from_ptr(lib().hostsim_ovf(h), n, np.uint32)
    ps = [C.c_void_p() for _ in range(8)]
    lib().hostsim_stage_views(h, *[C.byref(p) for p in ps])
    res.intv_off = np_from_ptr(ps[0], n + 1, np.int64)
    res.intv = np_from_ptr(ps[1], int(res.intv_off[-1]), INTV_DTYPE)
    res.chn_off = np_from_ptr(ps[2], n + 1, np.int64)
    nc = int(res.chn_off[-1])
    res.chn = np_from_ptr(ps[3], nc * 6, np.int64).reshape(-1, 6)
    res.seed_off = np_from_ptr(ps[4], nc + 1, np.int64)
    res.seeds = np_from_ptr(ps[5], int(res.seed_off[-1]) * 4, np.int64).reshape(-1, 4)
    res.reg_off = np_from_ptr(ps[6], n + 1, np.int64)
    res.regs = np_from_ptr(ps[7], int(res.reg_off[-1]), HIT_DTYPE)
    lib().hostsim_results_free(h)
    return res
