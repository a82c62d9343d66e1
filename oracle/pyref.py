"""TEST INFRASTRUCTURE ONLY: ctypes bindings for oracle/_ref/libseqref_bwa.so
(the reference's own bwa C compiled from the mount, see oracle/Makefile) and
oracle/liboracle.so (our plain-C restatement).

Imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import numpy as np

from seqlib_b200.abi import (MemOpt, IndexView, ResultsView, Results, Contig, HIT_DTYPE, INTV_DTYPE,
                             EXT_JOB_DTYPE, EXT_OUT_DTYPE, np_from_ptr, pack_reads)

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF_SO = os.path.join(_HERE, "_ref", "libseqref_bwa.so")
_FML_SO = os.path.join(_HERE, "_ref", "libseqref_fml.so")
_ORACLE_SO = os.path.join(_HERE, "liboracle.so")

_lib = None


def have_ref():
    return os.path.exists(_REF_SO)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(_REF_SO)
        L.refdrv_index_construct.restype = C.c_void_p
        L.refdrv_index_construct.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p)]
        L.refdrv_index_load.restype = C.c_void_p
        L.refdrv_index_load.argtypes = [C.c_char_p]
        L.refdrv_index_from_view.restype = C.c_void_p
        L.refdrv_index_from_view.argtypes = [C.POINTER(IndexView)]
        L.refdrv_index_view.argtypes = [C.c_void_p, C.POINTER(IndexView)]
        L.refdrv_index_write.argtypes = [C.c_void_p, C.c_char_p]
        L.refdrv_index_destroy.argtypes = [C.c_void_p]
        L.refdrv_opt_init.argtypes = [C.POINTER(MemOpt)]
        L.refdrv_align.restype = C.c_void_p
        L.refdrv_align.argtypes = [C.c_void_p, C.POINTER(MemOpt), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_int, C.POINTER(C.c_double)]
        L.refdrv_results_view.argtypes = [C.c_void_p, C.POINTER(ResultsView)]
        L.refdrv_results_free.argtypes = [C.c_void_p]
        L.refdrv_process_seqs.restype = C.c_double
        L.refdrv_process_seqs.argtypes = [C.c_void_p, C.POINTER(MemOpt), C.c_int64, C.c_void_p, C.c_void_p, C.c_int]
        L.refdrv_collect_intv.argtypes = [C.c_void_p, C.POINTER(MemOpt), C.c_int64, C.c_void_p, C.c_void_p,
                                          C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
        L.refdrv_chains.argtypes = [C.c_void_p, C.POINTER(MemOpt), C.c_int64, C.c_void_p, C.c_void_p] + \
            [C.POINTER(C.c_void_p)] * 4
        L.refdrv_regs_raw.argtypes = [C.c_void_p, C.POINTER(MemOpt), C.c_int64, C.c_void_p, C.c_void_p,
                                      C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
        L.refdrv_ksw_extend2_batch.restype = C.c_double
        L.refdrv_ksw_extend2_batch.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                               C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.refdrv_free.argtypes = [C.c_void_p]
        L.refdrv_srand48.argtypes = [C.c_long]
        L.refdrv_lrand48.restype = C.c_long
        _lib = L
    return _lib


def default_opt():
    o = MemOpt()
    lib().refdrv_opt_init(C.byref(o))
    return o


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class RefIndex:
    def __init__(self, handle):
        self.h = handle
        self._keep = None

    @classmethod
    def construct(cls, names, seqs):
        n = len(names)
        an = (C.c_char_p * n)(*[s.encode() for s in names])
        asq = (C.c_char_p * n)(*[s.encode() for s in seqs])
        return cls(lib().refdrv_index_construct(n, an, asq))

    @classmethod
    def load(cls, prefix):
        h = lib().refdrv_index_load(prefix.encode())
        if not h:
            raise RuntimeError("bwa_idx_load failed for " + prefix)
        return cls(h)

    @classmethod
    def from_view(cls, view, keep=None):
        r = cls(lib().refdrv_index_from_view(C.byref(view)))
        r._keep = keep
        return r

    def view(self):
        v = IndexView()
        lib().refdrv_index_view(self.h, C.byref(v))
        return v

    def arrays(self):
        """dict of numpy copies of the bwa-layout arrays."""
        v = self.view()
        return dict(primary=v.primary, L2=list(v.L2), seq_len=v.seq_len, bwt_size=v.bwt_size,
                    bwt=np_from_ptr(v.bwt, v.bwt_size, np.uint32), sa_intv=v.sa_intv, n_sa=v.n_sa,
                    sa=np_from_ptr(v.sa, v.n_sa, np.uint64), l_pac=v.l_pac,
                    pac=np_from_ptr(v.pac, v.l_pac // 4 + 1, np.uint8), n_seqs=v.n_seqs,
                    contigs=[(v.contigs[i].name.decode(), v.contigs[i].offset, v.contigs[i].len) for i in range(v.n_seqs)])

    def write(self, prefix):
        return lib().refdrv_index_write(self.h, prefix.encode())

    def __del__(self):
        if self.h:
            lib().refdrv_index_destroy(self.h)
            self.h = None


def align(idx, reads, opt=None, ids=None, n_threads=1):
    """mem_align1 + mem_reg2aln per read; returns (Results, seconds)."""
    seqs, off = pack_reads(reads) if not isinstance(reads, tuple) else reads
    n = len(off) - 1
    ids_a = None if ids is None else np.ascontiguousarray(ids, dtype=np.int64)
    sec = C.c_double(0)
    h = lib().refdrv_align(idx.h, C.byref(opt) if opt is not None else None, n, _p(seqs), _p(off), _p(ids_a),
                           n_threads, C.byref(sec))
    v = ResultsView()
    lib().refdrv_results_view(h, C.byref(v))
    res = Results(v)
    lib().refdrv_results_free(h)
    return res, sec.value


def process_seqs(idx, reads, opt=None, n_threads=1):
    seqs, off = pack_reads(reads) if not isinstance(reads, tuple) else reads
    return lib().refdrv_process_seqs(idx.h, C.byref(opt) if opt is not None else None, len(off) - 1, _p(seqs), _p(off),
                                     n_threads)


def collect_intv(idx, reads, opt=None):
    seqs, off = pack_reads(reads) if not isinstance(reads, tuple) else reads
    n = len(off) - 1
    po, pi = C.c_void_p(), C.c_void_p()
    lib().refdrv_collect_intv(idx.h, C.byref(opt) if opt is not None else None, n, _p(seqs), _p(off), C.byref(po), C.byref(pi))
    ioff = np_from_ptr(po, n + 1, np.int64)
    intv = np_from_ptr(pi, int(ioff[-1]), INTV_DTYPE)
    lib().refdrv_free(po)
    lib().refdrv_free(pi)
    return ioff, intv


def chains(idx, reads, opt=None):
    seqs, off = pack_reads(reads) if not isinstance(reads, tuple) else reads
    n = len(off) - 1
    ps = [C.c_void_p() for _ in range(4)]
    lib().refdrv_chains(idx.h, C.byref(opt) if opt is not None else None, n, _p(seqs), _p(off), *[C.byref(p) for p in ps])
    coff = np_from_ptr(ps[0], n + 1, np.int64)
    nc = int(coff[-1])
    chn = np_from_ptr(ps[1], nc * 6, np.int64).reshape(-1, 6)
    soff = np_from_ptr(ps[2], nc + 1, np.int64)
    seeds = np_from_ptr(ps[3], int(soff[-1]) * 4 if nc else 0, np.int64).reshape(-1, 4)
    for p in ps:
        lib().refdrv_free(p)
    return coff, chn, soff, seeds


def regs_raw(idx, reads, opt=None):
    seqs, off = pack_reads(reads) if not isinstance(reads, tuple) else reads
    n = len(off) - 1
    po, pr = C.c_void_p(), C.c_void_p()
    lib().refdrv_regs_raw(idx.h, C.byref(opt) if opt is not None else None, n, _p(seqs), _p(off), C.byref(po), C.byref(pr))
    roff = np_from_ptr(po, n + 1, np.int64)
    regs = np_from_ptr(pr, int(roff[-1]), HIT_DTYPE)
    lib().refdrv_free(po)
    lib().refdrv_free(pr)
    return roff, regs


def ksw_extend2_batch(jobs, qpool, tpool, mat, o_del=6, e_del=1, o_ins=6, e_ins=1, n_threads=1):
    jobs = np.ascontiguousarray(jobs, dtype=EXT_JOB_DTYPE)
    out = np.zeros(len(jobs), dtype=EXT_OUT_DTYPE)
    mat = np.ascontiguousarray(mat, dtype=np.int8)
    sec = lib().refdrv_ksw_extend2_batch(len(jobs), _p(jobs), _p(qpool), _p(tpool), _p(mat), o_del, e_del, o_ins, e_ins,
                                         _p(out), n_threads)
    return out, sec


def srand48(seed):
    lib().refdrv_srand48(seed)


def sam(idx, reads, opt, ids, names, quals=None, comments=None):
    """SAM text of mem_align1 + mem_reg2sam per read (oracle/refdrv_bwa.c refdrv_sam); names/quals/comments: lists of bytes"""
    seqs, off = pack_reads(reads) if not isinstance(reads, tuple) else reads
    n = len(off) - 1

    def flat(items):
        o = np.zeros(len(items) + 1, dtype=np.int64)
        o[1:] = np.cumsum([len(x) for x in items])
        return np.frombuffer(b"".join(items) + b"\0", dtype=np.uint8).copy(), o
    nm, nmo = flat(names)
    q, qo = flat(quals) if quals is not None else (None, None)
    c, co = flat(comments) if comments is not None else (None, None)
    ids_a = None if ids is None else np.ascontiguousarray(ids, dtype=np.int64)
    out = C.c_void_p(); ln = C.c_int64()
    L = lib()
    L.refdrv_sam.argtypes = [C.c_void_p, C.c_void_p, C.c_int64] + [C.c_void_p] * 9 + [C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
    L.refdrv_sam(idx.h, C.byref(opt), n, _p(seqs), _p(off), _p(ids_a), _p(nm), _p(nmo), _p(q), _p(qo), _p(c), _p(co), C.byref(out), C.byref(ln))
    text = C.string_at(out, ln.value)
    L.refdrv_free(out)
    return text
