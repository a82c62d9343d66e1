"""ctypes binding of b200_results_to_sam (include/seqlib_b200.h): SAM text of single-end alignments = bwa's mem_reg2sam."""
import ctypes as C
import numpy as np

from .abi import MemOpt, ResultsView, HIT_DTYPE
from .capi import lib, _check

_bound = False


def flat(items):
    """list of bytes -> (uint8 array, int64 offsets)"""
    off = np.zeros(len(items) + 1, dtype=np.int64)
    if items:
        off[1:] = np.cumsum([len(x) for x in items])
    data = np.frombuffer(b"".join(items), dtype=np.uint8).copy() if items else np.zeros(0, np.uint8)
    return data, off


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None and len(a) else None


def view_of(res):
    """ResultsView over the numpy arrays of a Results object (kept alive by the caller)"""
    v = ResultsView()
    keep = [np.ascontiguousarray(res.hit_off, dtype=np.int64), np.ascontiguousarray(res.hits, dtype=HIT_DTYPE),
            np.ascontiguousarray(res.cigar, dtype=np.uint32), np.frombuffer(res.md + b"\0", dtype=np.uint8).copy()]
    v.n_reads = len(keep[0]) - 1
    v.hit_off = keep[0].ctypes.data_as(C.POINTER(C.c_int64))
    v.hits = keep[1].ctypes.data_as(C.c_void_p)
    v.cigar = keep[2].ctypes.data_as(C.POINTER(C.c_uint32))
    v.md = keep[3].ctypes.data_as(C.POINTER(C.c_char))
    v.n_hits = len(keep[1]); v.n_cigar = len(keep[2]); v.n_md = len(keep[3])
    return v, keep


def results_to_sam(res, opt, rnames, seqs, seq_off, names, quals=None, comments=None):
    """res: abi.Results; rnames: contig names; seqs/seq_off: flat reads; names/quals/comments: lists of bytes -> bytes"""
    return results_to_sam_flat(res, opt, rnames, seqs, seq_off, flat(names), flat(quals) if quals is not None else None,
                               flat(comments) if comments is not None else None)


def results_to_sam_flat(res, opt, rnames, seqs, seq_off, names, quals=None, comments=None, timing=None):
    """the same with names / quals / comments already as (bytes array, offsets) pairs; timing: dict that receives the seconds
    spent inside the C call"""
    import time
    global _bound
    L = lib()
    if not _bound:
        L.b200_results_to_sam.argtypes = [C.POINTER(ResultsView), C.POINTER(MemOpt), C.POINTER(C.c_char_p), C.c_int] + [C.c_void_p] * 8 + \
            [C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
        _bound = True
    v, keep = view_of(res)
    rn = (C.c_char_p * len(rnames))(*[r.encode() if isinstance(r, str) else r for r in rnames])
    seqs = np.ascontiguousarray(seqs, dtype=np.uint8); seq_off = np.ascontiguousarray(seq_off, dtype=np.int64)
    nm, nmo = names
    q, qo = quals if quals is not None else (None, None)
    c, co = comments if comments is not None else (None, None)
    empty = np.zeros(1, np.uint8)
    out = C.c_void_p(); n = C.c_int64()
    t0 = time.perf_counter()
    _check(L.b200_results_to_sam(C.byref(v), C.byref(opt), rn, len(rnames), _p(seqs), _p(seq_off),
                                 (_p(q) or empty.ctypes.data_as(C.c_void_p)) if q is not None else None, _p(qo),
                                 _p(nm) or empty.ctypes.data_as(C.c_void_p), _p(nmo),
                                 (_p(c) or empty.ctypes.data_as(C.c_void_p)) if c is not None else None, _p(co),
                                 C.byref(out), C.byref(n)))
    if timing is not None:
        timing["seconds"] = time.perf_counter() - t0
    text = C.string_at(out, n.value)
    C.CDLL(None).free(out)
    return text
