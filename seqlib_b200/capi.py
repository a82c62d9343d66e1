"""ctypes binding of libseqlib_b200.so (the CUDA engine behind the C ABI in include/seqlib_b200.h).

This is what tests and bench.py call: every function goes through the same
extern "C" entry points a C++/cgo/JNI binding would use.  There is no CPU
fallback: if the shared library is missing, or no CUDA device is usable, the
calls raise.
"""
import ctypes as C
import os
import numpy as np

from .abi import (MemOpt, Contig, IndexView, ResultsView, Results, StageStats, INTV_DTYPE, EXT_JOB_DTYPE, EXT_OUT_DTYPE,
                  np_from_ptr, pack_reads, FmlOpt, Fseq1, FmlStats, Utg, utgs_to_py)

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libseqlib_b200.so")
_lib = None

EXPORTS = [
    "b200_last_error", "b200_mem_opt_init", "b200_fill_scmat", "b200_index_construct", "b200_index_construct_pac",
    "b200_index_load", "b200_index_write", "b200_index_destroy", "b200_index_view", "b200_index_blob_bytes",
    "b200_index_export_blob", "b200_index_attach_blob", "b200_index_n_seqs", "b200_index_seq_name", "b200_index_seq_len",
    "b200_index_l_pac", "b200_mem_align_batch", "b200_results_view", "b200_results_free", "b200_batch_create",
    "b200_batch_run", "b200_batch_fetch", "b200_batch_destroy", "b200_last_stats", "b200_debug_collect_intv",
    "b200_ksw_extend2_batch", "b200_set_device", "b200_device_count",
    "b200_shard_bounds", "b200_comm_unique_id", "b200_comm_init", "b200_comm_destroy", "b200_index_bcast", "b200_comm_last_bcast_ms",
    "b200_reads_scatter",
    "b200_fml_opt_init", "b200_fml_opt_adjust", "b200_fml_opt_adjust_lens", "b200_fml_correct", "b200_fml_fltuniq",
    "b200_fml_correct_flat", "b200_fml_count", "b200_kmer_table_hist", "b200_kmer_table_size", "b200_kmer_table_lookup",
    "b200_kmer_correct_flat", "b200_kmer_table_destroy", "b200_fml_last_stats",
    "b200_fml_assemble_flat", "b200_fml_seqs2utg_flat", "b200_fml_assemble", "b200_fml_utg_destroy", "b200_utgs_view",
    "b200_utgs_free", "b200_fmd_build", "b200_fmd_len", "b200_fmd_info", "b200_fmd_bwt", "b200_fmd_rank1a", "b200_fmd_destroy",
    "b200_fml_mag_text",
]


class B200Error(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise B200Error("libseqlib_b200.so is not built (run __graft_entry__.build()); there is no CPU fallback")
        L = C.CDLL(SO_PATH)
        L.b200_last_error.restype = C.c_char_p
        L.b200_mem_opt_init.argtypes = [C.POINTER(MemOpt)]
        L.b200_index_construct.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.c_int, C.POINTER(C.c_void_p)]
        L.b200_index_construct_pac.argtypes = [C.c_int64, C.c_void_p, C.c_int, C.POINTER(Contig), C.c_int, C.POINTER(C.c_void_p)]
        L.b200_index_load.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
        L.b200_index_write.argtypes = [C.c_void_p, C.c_char_p]
        L.b200_index_destroy.argtypes = [C.c_void_p]
        L.b200_index_view.argtypes = [C.c_void_p, C.POINTER(IndexView)]
        L.b200_index_blob_bytes.restype = C.c_int64
        L.b200_index_blob_bytes.argtypes = [C.c_void_p]
        L.b200_index_export_blob.argtypes = [C.c_void_p, C.c_void_p]
        L.b200_index_attach_blob.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_void_p)]
        L.b200_index_n_seqs.argtypes = [C.c_void_p]
        L.b200_index_seq_name.restype = C.c_char_p
        L.b200_index_seq_name.argtypes = [C.c_void_p, C.c_int]
        L.b200_index_seq_len.restype = C.c_int64
        L.b200_index_seq_len.argtypes = [C.c_void_p, C.c_int]
        L.b200_index_l_pac.restype = C.c_int64
        L.b200_index_l_pac.argtypes = [C.c_void_p]
        L.b200_mem_align_batch.argtypes = [C.c_void_p, C.POINTER(MemOpt), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
        L.b200_results_view.argtypes = [C.c_void_p, C.POINTER(ResultsView)]
        L.b200_results_free.argtypes = [C.c_void_p]
        L.b200_batch_create.argtypes = [C.c_void_p, C.POINTER(MemOpt), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
        L.b200_batch_run.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        L.b200_batch_fetch.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
        L.b200_batch_destroy.argtypes = [C.c_void_p]
        L.b200_last_stats.argtypes = [C.POINTER(StageStats)]
        L.b200_debug_collect_intv.argtypes = [C.c_void_p, C.POINTER(MemOpt), C.c_int64, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
        L.b200_ksw_extend2_batch.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p,
                                             C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_float)]
        L.b200_set_device.argtypes = [C.c_int]
        L.b200_fml_opt_init.argtypes = [C.POINTER(FmlOpt)]
        L.b200_fml_opt_adjust.argtypes = [C.POINTER(FmlOpt), C.c_int, C.POINTER(Fseq1)]
        L.b200_fml_opt_adjust_lens.argtypes = [C.POINTER(FmlOpt), C.c_int64, C.c_int64]
        L.b200_fml_correct.argtypes = [C.POINTER(FmlOpt), C.c_int, C.POINTER(Fseq1), C.POINTER(C.c_float)]
        L.b200_fml_fltuniq.argtypes = [C.POINTER(FmlOpt), C.c_int, C.POINTER(Fseq1), C.POINTER(C.c_float)]
        L.b200_fml_correct_flat.argtypes = [C.POINTER(FmlOpt), C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.POINTER(C.c_float)]
        L.b200_fml_count.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.b200_kmer_table_hist.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
        L.b200_kmer_table_size.restype = C.c_int64
        L.b200_kmer_table_size.argtypes = [C.c_void_p]
        L.b200_kmer_table_lookup.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.b200_kmer_correct_flat.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.b200_kmer_table_destroy.argtypes = [C.c_void_p]
        L.b200_fml_last_stats.argtypes = [C.POINTER(FmlStats)]
        L.b200_fml_assemble_flat.argtypes = [C.POINTER(FmlOpt), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
        L.b200_fml_seqs2utg_flat.argtypes = [C.POINTER(FmlOpt), C.c_int64, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
        L.b200_fml_assemble.argtypes = [C.POINTER(FmlOpt), C.c_int, C.POINTER(Fseq1), C.POINTER(C.c_int), C.POINTER(C.POINTER(Utg))]
        L.b200_fml_utg_destroy.argtypes = [C.c_int, C.POINTER(Utg)]
        L.b200_utgs_view.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.POINTER(Utg))]
        L.b200_utgs_free.argtypes = [C.c_void_p]
        L.b200_fmd_build.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
        L.b200_fmd_len.restype = C.c_int64
        L.b200_fmd_len.argtypes = [C.c_void_p]
        L.b200_fmd_info.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.b200_fmd_bwt.argtypes = [C.c_void_p, C.c_void_p]
        L.b200_fmd_rank1a.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.b200_fmd_destroy.argtypes = [C.c_void_p]
        L.b200_fml_mag_text.argtypes = [C.POINTER(FmlOpt), C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p),
                                        C.POINTER(C.c_int64), C.POINTER(C.c_float)]
        L.b200_shard_bounds.restype = None
        L.b200_shard_bounds.argtypes = [C.c_int64, C.c_int, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.b200_comm_unique_id.argtypes = [C.c_char_p]
        L.b200_comm_init.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.b200_comm_destroy.restype = None
        L.b200_comm_destroy.argtypes = [C.c_void_p]
        L.b200_index_bcast.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
        L.b200_comm_last_bcast_ms.restype = C.c_float
        L.b200_comm_last_bcast_ms.argtypes = [C.c_void_p]
        L.b200_reads_scatter.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise B200Error("b200 error %d: %s" % (rc, lib().b200_last_error().decode(errors="replace")))


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def default_opt(softclip=True):
    """mem_opt_init() plus SeqLib's MEM_F_SOFTCLIP (SeqLib/BWAAligner.h:15-17)."""
    o = MemOpt()
    lib().b200_mem_opt_init(C.byref(o))
    if softclip:
        o.flag |= 0x200
    return o


def set_device(i):
    _check(lib().b200_set_device(i))


def shard_bounds(n_total, world, rank):
    """Reads [b, e) of `rank`: contiguous shards by global read index (b200_shard_bounds)."""
    b, e = C.c_int64(), C.c_int64()
    lib().b200_shard_bounds(n_total, world, rank, C.byref(b), C.byref(e))
    return b.value, e.value


def comm_unique_id():
    """128-byte NCCL id made by rank 0; ship it to the other ranks over any channel, then Comm(id, rank, world) everywhere."""
    buf = C.create_string_buffer(128)
    _check(lib().b200_comm_unique_id(buf))
    return buf.raw


class Comm:
    """One process per GPU: the index image is replicated by one NCCL broadcast, read batches are scattered by contiguous shard."""

    def __init__(self, uid, rank, world):
        self.h = C.c_void_p()
        self.rank, self.world = rank, world
        _check(lib().b200_comm_init(C.c_char_p(uid), rank, world, C.byref(self.h)))

    def index_bcast(self, idx, root=0):
        """root passes its Index and gets it back; the other ranks pass None and get a replica."""
        out = C.c_void_p()
        _check(lib().b200_index_bcast(self.h, idx.h if idx is not None else None, root, C.byref(out)))
        return idx if self.rank == root else Index(out)

    def last_bcast_ms(self):
        return float(lib().b200_comm_last_bcast_ms(self.h))

    def reads_scatter(self, seqs, n_total, read_len, root=0):
        """seqs: uint8 array of n_total * read_len bases on root (None elsewhere) -> this rank's shard (uint8), b, e."""
        b, e = shard_bounds(n_total, self.world, self.rank)
        shard = np.empty((e - b) * read_len, dtype=np.uint8)
        src = np.ascontiguousarray(seqs, dtype=np.uint8) if seqs is not None else None
        _check(lib().b200_reads_scatter(self.h, root, n_total, read_len, _p(src), _p(shard)))
        return shard, b, e

    def close(self):
        if self.h:
            lib().b200_comm_destroy(self.h)
            self.h = C.c_void_p()


class Index:
    def __init__(self, handle, keep=None):
        self.h = C.c_void_p(handle) if not isinstance(handle, C.c_void_p) else handle
        self._keep = keep

    @classmethod
    def construct(cls, names, seqs, keep_host=True):
        n = len(names)
        an = (C.c_char_p * n)(*[s.encode() for s in names])
        asq = (C.c_char_p * n)(*[s.encode() for s in seqs])
        h = C.c_void_p()
        _check(lib().b200_index_construct(n, an, asq, 1 if keep_host else 0, C.byref(h)))
        return cls(h)

    @classmethod
    def construct_pac(cls, pac, l_pac, contigs, keep_host=False):
        """pac: np.uint8 forward 2-bit (bwa order); contigs: list of (name, offset, len)."""
        n = len(contigs)
        arr = (Contig * n)()
        keep = []
        for i, (name, off, ln) in enumerate(contigs):
            nb = name.encode()
            keep.append(nb)
            arr[i].offset = off
            arr[i].len = ln
            arr[i].name = nb
            arr[i].anno = b"(null)"
        pac = np.ascontiguousarray(pac, dtype=np.uint8)
        h = C.c_void_p()
        _check(lib().b200_index_construct_pac(l_pac, _p(pac), n, arr, 1 if keep_host else 0, C.byref(h)))
        return cls(h, keep=keep)

    @classmethod
    def load(cls, prefix):
        h = C.c_void_p()
        _check(lib().b200_index_load(prefix.encode(), C.byref(h)))
        return cls(h)

    @classmethod
    def attach_blob(cls, dev_ptr, nbytes, keep=None):
        h = C.c_void_p()
        _check(lib().b200_index_attach_blob(C.c_void_p(dev_ptr), nbytes, C.byref(h)))
        return cls(h, keep=keep)

    def write(self, prefix):
        _check(lib().b200_index_write(self.h, prefix.encode()))

    def view(self):
        v = IndexView()
        _check(lib().b200_index_view(self.h, C.byref(v)))
        return v

    def arrays(self):
        v = self.view()
        return dict(primary=v.primary, L2=list(v.L2), seq_len=v.seq_len, bwt_size=v.bwt_size,
                    bwt=np_from_ptr(v.bwt, v.bwt_size, np.uint32), sa_intv=v.sa_intv, n_sa=v.n_sa,
                    sa=np_from_ptr(v.sa, v.n_sa, np.uint64), l_pac=v.l_pac,
                    pac=np_from_ptr(v.pac, v.l_pac // 4 + 1, np.uint8), n_seqs=v.n_seqs,
                    contigs=[(v.contigs[i].name.decode(), v.contigs[i].offset, v.contigs[i].len) for i in range(v.n_seqs)])

    def blob_bytes(self):
        return lib().b200_index_blob_bytes(self.h)

    def export_blob(self, dev_ptr):
        _check(lib().b200_index_export_blob(self.h, C.c_void_p(dev_ptr)))

    def n_seqs(self):
        return lib().b200_index_n_seqs(self.h)

    def seq_name(self, rid):
        s = lib().b200_index_seq_name(self.h, rid)
        return None if s is None else s.decode()

    def seq_len(self, rid):
        return lib().b200_index_seq_len(self.h, rid)

    def l_pac(self):
        return lib().b200_index_l_pac(self.h)

    def close(self):
        if self.h:
            lib().b200_index_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _results_from(handle):
    v = ResultsView()
    _check(lib().b200_results_view(handle, C.byref(v)))
    res = Results(v)
    lib().b200_results_free(handle)
    return res


def align(idx, reads, opt=None, ids=None):
    """b200_mem_align_batch on host buffers; returns Results."""
    seqs, off = pack_reads(reads) if not isinstance(reads, tuple) else reads
    n = len(off) - 1
    opt = opt if opt is not None else default_opt()
    ids_a = None if ids is None else np.ascontiguousarray(ids, dtype=np.int64)
    h = C.c_void_p()
    _check(lib().b200_mem_align_batch(idx.h, C.byref(opt), n, _p(seqs), _p(off), _p(ids_a), C.byref(h)))
    return _results_from(h)


def align_raw(idx, seqs, off, opt, ids):
    """Same call, returns the opaque results handle (for timing without the numpy copies)."""
    h = C.c_void_p()
    _check(lib().b200_mem_align_batch(idx.h, C.byref(opt), len(off) - 1, _p(seqs), _p(off), _p(ids), C.byref(h)))
    return h


def results_free(h):
    lib().b200_results_free(h)


def results_summary(h):
    v = ResultsView()
    _check(lib().b200_results_view(h, C.byref(v)))
    return dict(n_reads=v.n_reads, n_hits=v.n_hits, n_cigar=v.n_cigar, n_md=v.n_md)


class Batch:
    """Device-resident reads (bench 'value' leg)."""

    def __init__(self, idx, reads, opt=None, ids=None):
        seqs, off = pack_reads(reads) if not isinstance(reads, tuple) else reads
        self.n = len(off) - 1
        self.opt = opt if opt is not None else default_opt()
        ids_a = None if ids is None else np.ascontiguousarray(ids, dtype=np.int64)
        self.h = C.c_void_p()
        self._idx = idx
        _check(lib().b200_batch_create(idx.h, C.byref(self.opt), self.n, _p(seqs), _p(off), _p(ids_a), C.byref(self.h)))

    def run(self):
        nl = C.c_int(0)
        _check(lib().b200_batch_run(self.h, C.byref(nl)))
        return nl.value

    def fetch(self):
        h = C.c_void_p()
        _check(lib().b200_batch_fetch(self.h, C.byref(h)))
        return _results_from(h)

    def close(self):
        if self.h:
            lib().b200_batch_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def last_stats():
    s = StageStats()
    _check(lib().b200_last_stats(C.byref(s)))
    return {f: getattr(s, f) for f, _ in StageStats._fields_}


def collect_intv(idx, reads, opt=None):
    seqs, off = pack_reads(reads) if not isinstance(reads, tuple) else reads
    n = len(off) - 1
    opt = opt if opt is not None else default_opt()
    po, pi = C.c_void_p(), C.c_void_p()
    _check(lib().b200_debug_collect_intv(idx.h, C.byref(opt), n, _p(seqs), _p(off), C.byref(po), C.byref(pi)))
    ioff = np_from_ptr(po, n + 1, np.int64)
    intv = np_from_ptr(pi, int(ioff[-1]), INTV_DTYPE)
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    libc.free(po)
    libc.free(pi)
    return ioff, intv


def ksw_extend2_batch(jobs, qpool, tpool, mat, o_del=6, e_del=1, o_ins=6, e_ins=1):
    jobs = np.ascontiguousarray(jobs, dtype=EXT_JOB_DTYPE)
    qpool = np.ascontiguousarray(qpool, dtype=np.uint8)
    tpool = np.ascontiguousarray(tpool, dtype=np.uint8)
    mat = np.ascontiguousarray(mat, dtype=np.int8)
    out = np.zeros(len(jobs), dtype=EXT_OUT_DTYPE)
    cells = C.c_uint64(0)
    ms = C.c_float(0)
    _check(lib().b200_ksw_extend2_batch(len(jobs), _p(jobs), _p(qpool), len(qpool), _p(tpool), len(tpool), _p(mat),
                                        o_del, e_del, o_ins, e_ins, _p(out), C.byref(cells), C.byref(ms)))
    return out, cells.value, ms.value


# ---------------------------------------------------------------------------------------------------- fermi-lite half
def fml_default_opt():
    o = FmlOpt()
    lib().b200_fml_opt_init(C.byref(o))
    return o


def fml_opt_adjust(opt, n_seqs, tot_len):
    lib().b200_fml_opt_adjust_lens(C.byref(opt), int(n_seqs), int(tot_len))
    return opt


def fml_correct_flat(opt, seqs, quals, off, flt_uniq=False):
    """b200_fml_correct_flat on numpy pools.  Returns (seqs, quals, lens, kcov); the input arrays are not modified."""
    seqs = np.array(seqs, dtype=np.uint8, copy=True)
    quals = None if quals is None else np.array(quals, dtype=np.uint8, copy=True)
    off = np.ascontiguousarray(off, dtype=np.int64)
    n = len(off) - 1
    lens = np.zeros(max(n, 1), dtype=np.int32)
    kcov = C.c_float(0)
    _check(lib().b200_fml_correct_flat(C.byref(opt), int(bool(flt_uniq)), n, _p(seqs), _p(quals), _p(off), _p(lens), C.byref(kcov)))
    return seqs, quals, lens[:n], kcov.value


def fml_last_stats():
    st = FmlStats()
    _check(lib().b200_fml_last_stats(C.byref(st)))
    d = {k: getattr(st, k) for k, _ in FmlStats._fields_ if k != "ec_codes"}
    d["ec_codes"] = list(st.ec_codes)
    return d


class KmerTable:
    """b200_kmer_table_t: the device-resident k-mer count table (bfc_ch_t)."""
    h = None

    def __init__(self, seqs, quals, off, k, q=20, l_pre=20):
        seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
        quals = None if quals is None else np.ascontiguousarray(quals, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.int64)
        h = C.c_void_p()
        _check(lib().b200_fml_count(len(off) - 1, _p(seqs), _p(quals), _p(off), k, q, l_pre, C.byref(h)))
        self.h, self.k = h, k

    def hist(self):
        cnt = np.zeros(256, dtype=np.uint64)
        high = np.zeros(64, dtype=np.uint64)
        mode = C.c_int(0)
        _check(lib().b200_kmer_table_hist(self.h, _p(cnt), _p(high), C.byref(mode)))
        return cnt, high, mode.value

    def size(self):
        return lib().b200_kmer_table_size(self.h)

    def lookup(self, kmers):
        """kmers: list of k-long strings -> int32 occ (-1 absent, else high << 8 | total)."""
        pool, off = pack_reads(kmers)
        assert len(pool) == len(kmers) * self.k
        occ = np.zeros(max(len(kmers), 1), dtype=np.int32)
        _check(lib().b200_kmer_table_lookup(self.h, len(kmers), _p(pool), _p(occ)))
        return occ[:len(kmers)]

    def correct(self, seqs, quals, off, min_cov, mode, flt_uniq=False):
        seqs = np.array(seqs, dtype=np.uint8, copy=True)
        quals = None if quals is None else np.array(quals, dtype=np.uint8, copy=True)
        off = np.ascontiguousarray(off, dtype=np.int64)
        n = len(off) - 1
        lens = np.zeros(max(n, 1), dtype=np.int32)
        _check(lib().b200_kmer_correct_flat(self.h, min_cov, mode, int(bool(flt_uniq)), n, _p(seqs), _p(quals), _p(off), _p(lens)))
        return seqs, quals, lens[:n]

    def close(self):
        if self.h:
            lib().b200_kmer_table_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()


def _utgs_out(h):
    n = C.c_int(0)
    arr = C.POINTER(Utg)()
    _check(lib().b200_utgs_view(h, C.byref(n), C.byref(arr)))
    out = utgs_to_py(n.value, arr)
    lib().b200_utgs_free(h)
    return out


def fml_assemble_flat(opt, seqs, quals, off):
    """b200_fml_assemble_flat: list of unitig dicts (seq, cov, nsr, n_ovlp, ovlp)."""
    seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
    quals = None if quals is None else np.ascontiguousarray(quals, dtype=np.uint8)
    off = np.ascontiguousarray(off, dtype=np.int64)
    h = C.c_void_p()
    _check(lib().b200_fml_assemble_flat(C.byref(opt), len(off) - 1, _p(seqs), _p(quals), _p(off), C.byref(h)))
    return _utgs_out(h)


def fml_assemble_windows(opt, seqs, quals, off, win_off, n_threads=0):
    """b200_fml_assemble_windows: window w = reads [win_off[w], win_off[w+1]); returns one unitig list per window."""
    seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
    quals = None if quals is None else np.ascontiguousarray(quals, dtype=np.uint8)
    off = np.ascontiguousarray(off, dtype=np.int64)
    win_off = np.ascontiguousarray(win_off, dtype=np.int64)
    nw = len(win_off) - 1
    hs = (C.c_void_p * max(nw, 1))()
    L = lib()
    L.b200_fml_assemble_windows.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
    rc = L.b200_fml_assemble_windows(C.byref(opt), nw, _p(win_off), _p(seqs), _p(quals), _p(off), n_threads, hs)
    outs = []
    for w in range(nw):
        outs.append(_utgs_out(C.c_void_p(hs[w])) if hs[w] else None)
    _check(rc)
    return outs


def fml_seqs2utg_flat(opt, seqs, off):
    seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
    off = np.ascontiguousarray(off, dtype=np.int64)
    h = C.c_void_p()
    _check(lib().b200_fml_seqs2utg_flat(C.byref(opt), len(off) - 1, _p(seqs), _p(off), C.byref(h)))
    return _utgs_out(h)


def fml_mag_text(opt, stage, seqs, off):
    seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
    off = np.ascontiguousarray(off, dtype=np.int64)
    p = C.c_void_p()
    ln = C.c_int64(0)
    rd = C.c_float(0)
    _check(lib().b200_fml_mag_text(C.byref(opt), stage, len(off) - 1, _p(seqs), _p(off), C.byref(p), C.byref(ln), C.byref(rd)))
    txt = C.string_at(p, ln.value).decode() if p.value else ""
    if p.value:
        C.CDLL(None).free(p)
    return txt, rd.value


class Fmd:
    """b200_fmd_t: device-resident FMD-index of a read set."""
    h = None

    def __init__(self, seqs, off):
        seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.int64)
        h = C.c_void_p()
        _check(lib().b200_fmd_build(len(off) - 1, _p(seqs), _p(off), C.byref(h)))
        self.h = h

    def info(self):
        cnt = np.zeros(7, dtype=np.uint64)
        mcnt = np.zeros(7, dtype=np.uint64)
        _check(lib().b200_fmd_info(self.h, _p(cnt), _p(mcnt)))
        return cnt, mcnt

    def bwt(self):
        out = np.zeros(lib().b200_fmd_len(self.h), dtype=np.uint8)
        _check(lib().b200_fmd_bwt(self.h, _p(out)))
        return out

    def rank1a(self, q):
        q = np.ascontiguousarray(q, dtype=np.uint64)
        ranks = np.zeros((max(len(q), 1), 6), dtype=np.uint64)
        sym = np.zeros(max(len(q), 1), dtype=np.int32)
        _check(lib().b200_fmd_rank1a(self.h, len(q), _p(q), _p(ranks), _p(sym)))
        return ranks[:len(q)], sym[:len(q)]

    def close(self):
        if self.h:
            lib().b200_fmd_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()
