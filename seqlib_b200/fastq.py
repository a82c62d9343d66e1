"""ctypes binding of the FASTA/FASTQ ingest entry points (b200_fastq_*, include/seqlib_b200.h): the host-side mirror of
SeqLib::FastqReader (SeqLib/FastqReader.h:22-60) plus the batch / device-parser forms that feed b200_mem_align_batch."""
import ctypes as C
import numpy as np

from .capi import lib, _check


class FastqBatch(C.Structure):
    _fields_ = [("n", C.c_int64),
                ("seq", C.c_void_p), ("seq_off", C.c_void_p),
                ("qual", C.c_void_p), ("qual_off", C.c_void_p),
                ("name", C.c_void_p), ("name_off", C.c_void_p),
                ("comment", C.c_void_p), ("comment_off", C.c_void_p),
                ("status", C.c_int32), ("parsed_on_device", C.c_int32), ("has", C.c_void_p)]


_bound = False


def _bind():
    global _bound
    L = lib()
    if not _bound:
        L.b200_fastq_open.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
        L.b200_fastq_open_mem.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_void_p)]
        L.b200_fastq_next_batch.argtypes = [C.c_void_p, C.c_int64, C.POINTER(FastqBatch)]
        L.b200_fastq_parse_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(FastqBatch)]
        L.b200_fastq_buffers_seen.argtypes = [C.c_void_p]
        L.b200_fastq_close.argtypes = [C.c_void_p]
        _bound = True
    return L


def _field(ptr, off_ptr, n):
    off = np.ctypeslib.as_array(C.cast(off_ptr, C.POINTER(C.c_int64)), shape=(n + 1,)).copy()
    tot = int(off[n])
    data = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(max(tot, 1),))[:tot].copy()
    return data, off


class Batch:
    """One batch, copied out of the reader's buffers: four (bytes, offsets) pairs."""

    def __init__(self, b):
        self.n = int(b.n)
        self.status = int(b.status)
        self.parsed_on_device = int(b.parsed_on_device)
        self.seq, self.seq_off = _field(b.seq, b.seq_off, self.n)
        self.qual, self.qual_off = _field(b.qual, b.qual_off, self.n)
        self.name, self.name_off = _field(b.name, b.name_off, self.n)
        self.comment, self.comment_off = _field(b.comment, b.comment_off, self.n)
        self.has = np.ctypeslib.as_array(C.cast(b.has, C.POINTER(C.c_uint8)), shape=(max(self.n, 1),))[:self.n].copy() if self.n else np.zeros(0, np.uint8)

    def records(self):
        out = []
        for i in range(self.n):
            out.append(tuple(bytes(a[o[i]:o[i + 1]]) for a, o in ((self.name, self.name_off), (self.comment, self.comment_off),
                                                                  (self.seq, self.seq_off), (self.qual, self.qual_off))))
        return out


class FastqReader:
    def __init__(self, path=None, text=None):
        L = _bind()
        self.h = C.c_void_p()
        self._keep = None
        if path is not None:
            _check(L.b200_fastq_open(path.encode(), C.byref(self.h)))
        else:
            self._keep = np.frombuffer(text if text is not None else b"", dtype=np.uint8)
            _check(L.b200_fastq_open_mem(self._keep.ctypes.data_as(C.c_void_p) if len(self._keep) else None, len(self._keep), C.byref(self.h)))

    def next_batch(self, max_records):
        b = FastqBatch()
        _check(_bind().b200_fastq_next_batch(self.h, max_records, C.byref(b)))
        return Batch(b)

    def parse_device(self, text):
        a = np.frombuffer(text, dtype=np.uint8)
        b = FastqBatch()
        _check(_bind().b200_fastq_parse_device(self.h, a.ctypes.data_as(C.c_void_p) if len(a) else None, len(a), C.byref(b)))
        return Batch(b)

    def buffers_seen(self):
        return int(_bind().b200_fastq_buffers_seen(self.h))

    def close(self):
        if self.h:
            _bind().b200_fastq_close(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
