"""Deterministic synthetic references / reads; thin ctypes wrapper around synth.c."""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libb200synth.so")
_lib = None

REF_SEED = 0x5EED0001
READ_SEED = 0x5EED0002


def build():
    src = os.path.join(_HERE, "synth.c")
    if os.path.exists(_SO) and os.path.getmtime(_SO) >= os.path.getmtime(src):
        return
    subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", _SO, src, "-lpthread"])


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.synth_reference.argtypes = [C.c_uint64, C.c_int64, C.c_void_p, C.c_int]
        L.synth_pac_to_ascii.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]
        L.synth_reads.argtypes = [C.c_uint64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_double, C.c_double,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        _lib = L
    return _lib


def _threads():
    return max(1, min(32, os.cpu_count() or 1))


def reference(l_pac, seed=REF_SEED):
    """Forward 2-bit pac (bwa order) of a uniform random reference."""
    pac = np.zeros(l_pac // 4 + 2, dtype=np.uint8)
    lib().synth_reference(seed, l_pac, pac.ctypes.data, _threads())
    return pac


def contigs_for(l_pac, n_contigs, prefix="chr"):
    """n equal contigs named chr1..chrN (last one takes the remainder)."""
    per = l_pac // n_contigs
    out = []
    for i in range(n_contigs):
        off = i * per
        ln = per if i < n_contigs - 1 else l_pac - off
        out.append((prefix + str(i + 1), off, ln))
    return out


def ascii_of(pac, beg, end):
    out = np.zeros(end - beg, dtype=np.uint8)
    lib().synth_pac_to_ascii(pac.ctypes.data, beg, end, out.ctypes.data)
    return out.tobytes().decode()


def reads(pac, l_pac, contigs, n, length=150, sub=0.01, indel=0.0, seed=READ_SEED):
    """Returns (ascii uint8 array of n*length, int64 offsets, truth positions, strands)."""
    coff = np.array([c[1] for c in contigs] + [l_pac], dtype=np.int64)
    out = np.zeros(n * length, dtype=np.uint8)
    pos = np.zeros(n, dtype=np.int64)
    strand = np.zeros(n, dtype=np.int8)
    lib().synth_reads(seed, pac.ctypes.data, l_pac, coff.ctypes.data, len(contigs), n, length, sub, indel,
                      out.ctypes.data, pos.ctypes.data, strand.ctypes.data, _threads())
    off = np.arange(n + 1, dtype=np.int64) * length
    return out, off, pos, strand
