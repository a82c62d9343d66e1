"""TEST INFRASTRUCTURE ONLY: ctypes binding of oracle/_ref/libseqref_fml.so = the reference's unmodified
fermi-lite C compiled from the mount (oracle/Makefile) behind the flat-buffer driver oracle/refdrv_fml.c.

Imported only by tests/, tests/golden/make_golden_fml.py and bench.py's cpu_baseline / --impl reference legs.
"""
import ctypes as C
import os
import numpy as np

from seqlib_b200.abi import FmlOpt

_HERE = os.path.dirname(os.path.abspath(__file__))
_FML_SO = os.path.join(_HERE, "_ref", "libseqref_fml.so")
_lib = None


def have_ref():
    return os.path.exists(_FML_SO)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(_FML_SO)
        L.refdrv_fml_opt_init.argtypes = [C.POINTER(FmlOpt)]
        L.refdrv_fml_opt_size.restype = C.c_int
        L.refdrv_fml_correct.restype = C.c_float
        L.refdrv_fml_correct.argtypes = [C.POINTER(FmlOpt), C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
        L.refdrv_fml_count_hist.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                            C.POINTER(C.c_int64)]
        L.refdrv_fml_kmer_occ.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                          C.c_int64, C.c_void_p, C.c_void_p]
        L.refdrv_fml_assemble.restype = C.c_int
        L.refdrv_fml_assemble.argtypes = [C.POINTER(FmlOpt), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                          C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_double)]
        L.refdrv_fml_free.argtypes = [C.c_void_p]
        assert L.refdrv_fml_opt_size() == C.sizeof(FmlOpt)
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def default_opt():
    o = FmlOpt()
    lib().refdrv_fml_opt_init(C.byref(o))
    return o


def correct_flat(opt, seqs, quals, off, flt_uniq=False, adjust=False):
    """fml_correct / fml_fltuniq of the reference.  Returns (seqs, quals, lens, kcov, seconds)."""
    seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
    quals = None if quals is None else np.ascontiguousarray(quals, dtype=np.uint8)
    off = np.ascontiguousarray(off, dtype=np.int64)
    n = len(off) - 1
    so = seqs.copy()
    qo = None if quals is None else quals.copy()
    lens = np.zeros(max(n, 1), dtype=np.int32)
    sec = C.c_double(0)
    kcov = lib().refdrv_fml_correct(C.byref(opt), int(adjust), int(bool(flt_uniq)), n, _p(seqs), _p(quals), _p(off),
                                    _p(so), _p(qo), _p(lens), C.byref(sec))
    return so, qo, lens[:n], kcov, sec.value


def count_hist(seqs, quals, off, k, q=20, l_pre=20):
    """fml_count + bfc_ch_hist of the reference: (cnt[256], high[64], mode, n_distinct)."""
    seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
    quals = None if quals is None else np.ascontiguousarray(quals, dtype=np.uint8)
    off = np.ascontiguousarray(off, dtype=np.int64)
    h = np.zeros(320, dtype=np.uint64)
    nd = C.c_int64(0)
    mode = lib().refdrv_fml_count_hist(len(off) - 1, _p(seqs), _p(quals), _p(off), k, q, l_pre, _p(h), C.byref(nd))
    return h[:256].copy(), h[256:].copy(), mode, nd.value


def kmer_occ(seqs, quals, off, k, kmers, q=20, l_pre=20):
    """bfc_ch_kmer_occ of the reference for a list of k-long strings."""
    seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
    quals = None if quals is None else np.ascontiguousarray(quals, dtype=np.uint8)
    off = np.ascontiguousarray(off, dtype=np.int64)
    pool = np.frombuffer(b"".join(x.encode() if isinstance(x, str) else x for x in kmers), dtype=np.uint8).copy()
    occ = np.zeros(max(len(kmers), 1), dtype=np.int32)
    lib().refdrv_fml_kmer_occ(len(off) - 1, _p(seqs), _p(quals), _p(off), k, q, l_pre, len(kmers), _p(pool), _p(occ))
    return occ[:len(kmers)]


def bwt(seqs, off, queries=None):
    """fml_seq2fmi of the reference, decoded: (bwt symbols u8 0..5, cnt[7], mcnt[7], (ranks[nq,6], sym[nq]) for rld_rank1a queries)."""
    L = lib()
    L.refdrv_fml_bwt.restype = C.c_int64
    L.refdrv_fml_bwt.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p,
                                 C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
    off = np.ascontiguousarray(off, dtype=np.int64)
    q = np.ascontiguousarray(queries if queries is not None else np.zeros(0), dtype=np.uint64)
    ranks = np.zeros((max(len(q), 1), 6), dtype=np.uint64)
    sym = np.zeros(max(len(q), 1), dtype=np.int32)
    cnt = np.zeros(7, dtype=np.uint64)
    mcnt = np.zeros(7, dtype=np.uint64)
    p = C.c_void_p()
    n = L.refdrv_fml_bwt(len(off) - 1, _p(seqs), _p(off), C.byref(p), _p(cnt), _p(mcnt), len(q), _p(q), _p(ranks), _p(sym))
    out = np.zeros(0, dtype=np.uint8)
    if n:
        out = np.frombuffer((C.c_char * n).from_address(p.value), dtype=np.uint8, count=n).copy()
        L.refdrv_fml_free(p)
    return out, cnt, mcnt, (ranks[:len(q)], sym[:len(q)])


def mag_text(opt, stage, kcov, seqs, off):
    """fml_seq2fmi + fml_fmi2mag (+ fml_mag_clean for stage 1) of the reference as mag_g_print text."""
    L = lib()
    L.refdrv_fml_mag_text.restype = C.c_void_p
    L.refdrv_fml_mag_text.argtypes = [C.POINTER(FmlOpt), C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_void_p,
                                      C.POINTER(C.c_int64), C.POINTER(C.c_float), C.POINTER(C.c_double)]
    seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
    off = np.ascontiguousarray(off, dtype=np.int64)
    ln = C.c_int64(0)
    rd = C.c_float(0)
    sec = C.c_double(0)
    p = L.refdrv_fml_mag_text(C.byref(opt), stage, kcov, len(off) - 1, _p(seqs), _p(off), C.byref(ln), C.byref(rd), C.byref(sec))
    if not p:
        return "", 0.0, sec.value
    txt = C.string_at(p, ln.value).decode()
    L.refdrv_fml_free(p)
    return txt, rd.value, sec.value


def assemble(opt, seqs, quals, off):
    """fml_assemble of the reference: (list of dicts like seqlib_b200.abi.utgs_to_py, seconds)."""
    L = lib()
    seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
    quals = None if quals is None else np.ascontiguousarray(quals, dtype=np.uint8)
    off = np.ascontiguousarray(off, dtype=np.int64)
    ps = [C.c_void_p() for _ in range(6)]
    sec = C.c_double(0)
    n = L.refdrv_fml_assemble(C.byref(opt), len(off) - 1, _p(seqs), _p(quals), _p(off), *[C.byref(p) for p in ps], C.byref(sec))
    out = []
    if n >= 0 and ps[0].value:
        uo = np.frombuffer((C.c_char * (8 * (n + 1))).from_address(ps[0].value), dtype=np.int64).copy()
        tot = int(uo[-1])
        us = C.string_at(ps[1].value, tot)
        uc = C.string_at(ps[2].value, tot)
        nsr = np.frombuffer((C.c_char * (4 * max(n, 1))).from_address(ps[3].value), dtype=np.int32).copy()
        nv = np.frombuffer((C.c_char * (8 * max(n, 1))).from_address(ps[4].value), dtype=np.int32).copy()
        no = int(nv[:2 * n].sum())
        ov = np.frombuffer((C.c_char * (16 * max(no, 1))).from_address(ps[5].value), dtype=np.int32).copy().reshape(-1, 4)
        k = 0
        for i in range(n):
            m = int(nv[2 * i] + nv[2 * i + 1])
            out.append(dict(seq=us[uo[i]:uo[i + 1]], cov=uc[uo[i]:uo[i + 1]], nsr=int(nsr[i]), n_ovlp=(int(nv[2 * i]), int(nv[2 * i + 1])),
                            ovlp=[tuple(int(x) for x in ov[k + j]) for j in range(m)]))
            k += m
        for p in ps:
            L.refdrv_fml_free(p)
    return out, sec.value
