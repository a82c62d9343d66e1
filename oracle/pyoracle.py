"""TEST INFRASTRUCTURE ONLY: ctypes binding of oracle/liboracle.so, the plain-C restatement of the reference's
seed-and-extend path (oracle/oracle_bwa.c).  Works on bwa-layout arrays, so it needs neither the GPU nor oracle/_ref."""
import ctypes as C
import os
import numpy as np

from seqlib_b200.abi import MemOpt, IndexView, Contig, ResultsView, Results, EXT_JOB_DTYPE, EXT_OUT_DTYPE, pack_reads

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")
_lib = None


def have():
    return os.path.exists(_SO)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(_SO)
        L.oracle_align.restype = C.c_void_p
        L.oracle_align.argtypes = [C.POINTER(IndexView), C.POINTER(MemOpt), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_results_view.argtypes = [C.c_void_p, C.POINTER(ResultsView)]
        L.oracle_results_free.argtypes = [C.c_void_p]
        L.oracle_ksw_extend2_batch.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        _lib = L
    return _lib


class View:
    """An IndexView over numpy arrays (kept alive by this object)."""

    def __init__(self, primary, L2, bwt, sa, sa_intv, l_pac, pac, contigs):
        self.bwt = np.ascontiguousarray(bwt, dtype=np.uint32)
        self.sa = np.ascontiguousarray(sa, dtype=np.uint64)
        self.pac = np.ascontiguousarray(pac, dtype=np.uint8)
        self._names = [c[0].encode() for c in contigs]
        self._ctg = (Contig * len(contigs))()
        for i, (name, off, ln) in enumerate(contigs):
            self._ctg[i].offset = off
            self._ctg[i].len = ln
            self._ctg[i].name = self._names[i]
            self._ctg[i].anno = b""
        v = IndexView()
        v.primary = int(primary)
        for i in range(5):
            v.L2[i] = int(L2[i])
        v.seq_len = int(L2[4])
        v.bwt_size = len(self.bwt)
        v.bwt = self.bwt.ctypes.data_as(C.POINTER(C.c_uint32))
        v.sa_intv = sa_intv
        v.n_sa = len(self.sa)
        v.sa = self.sa.ctypes.data_as(C.POINTER(C.c_uint64))
        v.l_pac = l_pac
        v.pac = self.pac.ctypes.data_as(C.POINTER(C.c_uint8))
        v.n_seqs = len(contigs)
        v.contigs = self._ctg
        self.v = v


def load_bwa_index(prefix):
    """Parse .bwt/.sa/.pac/.ann written by bwa into a View."""
    raw = np.fromfile(prefix + ".bwt", dtype=np.uint8)
    primary = int(raw[:8].view(np.uint64)[0])
    L2 = [0] + [int(x) for x in raw[8:40].view(np.uint64)]
    bwt = raw[40:].view(np.uint32)
    sraw = np.fromfile(prefix + ".sa", dtype=np.uint8)
    hdr = sraw[:56].view(np.uint64)
    sa_intv = int(hdr[5] & 0xffffffff)
    sa = np.concatenate([np.array([0xFFFFFFFFFFFFFFFF], dtype=np.uint64), sraw[56:].view(np.uint64)])
    with open(prefix + ".ann") as f:
        toks = f.read().split("\n")
    l_pac, n_seqs, _seed = toks[0].split()
    l_pac, n_seqs = int(l_pac), int(n_seqs)
    contigs = []
    for i in range(n_seqs):
        name = toks[1 + 2 * i].split()[1]
        off, ln, _ = toks[2 + 2 * i].split()
        contigs.append((name, int(off), int(ln)))
    pac = np.fromfile(prefix + ".pac", dtype=np.uint8)[: l_pac // 4 + 1]
    return View(primary, L2, bwt, sa, sa_intv, l_pac, pac, contigs)


def align(view, reads, opt, ids):
    seqs, off = pack_reads(reads) if not isinstance(reads, tuple) else reads
    n = len(off) - 1
    ids = np.ascontiguousarray(ids, dtype=np.int64)
    h = lib().oracle_align(C.byref(view.v), C.byref(opt), n, seqs.ctypes.data, off.ctypes.data, ids.ctypes.data)
    rv = ResultsView()
    lib().oracle_results_view(h, C.byref(rv))
    res = Results(rv)
    lib().oracle_results_free(h)
    return res


def ksw_extend2_batch(jobs, qpool, tpool, mat, o_del=6, e_del=1, o_ins=6, e_ins=1):
    jobs = np.ascontiguousarray(jobs, dtype=EXT_JOB_DTYPE)
    out = np.zeros(len(jobs), dtype=EXT_OUT_DTYPE)
    mat = np.ascontiguousarray(mat, dtype=np.int8)
    lib().oracle_ksw_extend2_batch(len(jobs), jobs.ctypes.data, qpool.ctypes.data, tpool.ctypes.data, mat.ctypes.data, o_del, e_del, o_ins, e_ins, out.ctypes.data)
    return out


# ---------------------------------------------------------------------------------------------------- fermi-lite BFC stages
def fml_default_opt():
    """fml_opt_init values (fermi-lite/misc.c:31-41) without needing any library."""
    from seqlib_b200.abi import FmlOpt
    o = FmlOpt()
    o.n_threads, o.ec_k, o.min_cnt, o.max_cnt, o.min_asm_ovlp, o.min_merge_len = 1, 0, 4, 8, 33, 0
    return o


def fml_correct_flat(opt, seqs, quals, off, flt_uniq=False):
    """oracle_fml_correct_flat (oracle/oracle_fml.c): (seqs, quals, lens, kcov, hist[320]); inputs are not modified."""
    from seqlib_b200.abi import FmlOpt  # noqa: F401
    L = lib()
    L.oracle_fml_correct_flat.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.POINTER(C.c_float), C.c_void_p]
    seqs = np.array(seqs, dtype=np.uint8, copy=True)
    quals = None if quals is None else np.array(quals, dtype=np.uint8, copy=True)
    off = np.ascontiguousarray(off, dtype=np.int64)
    n = len(off) - 1
    lens = np.zeros(max(n, 1), dtype=np.int32)
    hist = np.zeros(320, dtype=np.uint64)
    kcov = C.c_float(0)
    L.oracle_fml_correct_flat(C.byref(opt), int(bool(flt_uniq)), n, seqs.ctypes.data, quals.ctypes.data if quals is not None else None,
                              off.ctypes.data, lens.ctypes.data, C.byref(kcov), hist.ctypes.data)
    return seqs, quals, lens[:n], kcov.value, hist
