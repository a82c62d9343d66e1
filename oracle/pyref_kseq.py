"""TEST INFRASTRUCTURE ONLY: ctypes binding of oracle/_ref/libseqref_kseq.so = the reference's bwa/kseq.h instantiated over
gzread and driven like FastqReader::GetNextSequence (oracle/refdrv_kseq.c).  Imported only by tests/ and
tests/golden/make_golden_fastq.py and bench scripts' CPU baseline legs."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libseqref_kseq.so")
_lib = None


def have_ref():
    return os.path.exists(_SO)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(_SO)
        L.refdrv_kseq_parse.restype = C.c_int64
        L.refdrv_kseq_parse.argtypes = [C.c_char_p, C.c_int64, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int)]
        L.refdrv_kseq_free.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def parse(path, max_rec=1 << 62):
    """-> (records [(name, comment, seq, qual) bytes], has [int per record], last kseq_read value)"""
    L = lib()
    fields = (C.c_void_p * 4)()
    offs = (C.c_void_p * 4)()
    has = C.c_void_p()
    last = C.c_int()
    n = L.refdrv_kseq_parse(path.encode(), max_rec, fields, offs, C.byref(has), C.byref(last))
    if n < 0:
        raise IOError(path)
    cols = []
    for f in range(4):
        o = C.cast(offs[f], C.POINTER(C.c_int64))
        tot = o[n]
        data = C.string_at(fields[f], tot)
        cols.append([data[o[i]:o[i + 1]] for i in range(n)])
    h = C.cast(has, C.POINTER(C.c_int32))
    hv = [h[i] for i in range(n)]
    for f in range(4):
        L.refdrv_kseq_free(fields[f]); L.refdrv_kseq_free(offs[f])
    L.refdrv_kseq_free(has)
    return list(zip(*cols)) if n else [], hv, last.value
