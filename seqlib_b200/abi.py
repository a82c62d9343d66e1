"""ctypes/numpy mirrors of the C ABI structs in include/seqlib_b200.h.

Pure type definitions (no library loading) so that both the product binding
(seqlib_b200.capi) and the test-only oracle bindings (oracle/pyref.py) agree
on layouts.
"""
import ctypes as C
import numpy as np


class MemOpt(C.Structure):
    """b200_mem_opt_t == bwa mem_opt_t (bwa/bwamem.h:52-84)."""
    _fields_ = [
        ("a", C.c_int), ("b", C.c_int),
        ("o_del", C.c_int), ("e_del", C.c_int),
        ("o_ins", C.c_int), ("e_ins", C.c_int),
        ("pen_unpaired", C.c_int),
        ("pen_clip5", C.c_int), ("pen_clip3", C.c_int),
        ("w", C.c_int), ("zdrop", C.c_int),
        ("max_mem_intv", C.c_uint64),
        ("T", C.c_int), ("flag", C.c_int),
        ("min_seed_len", C.c_int), ("min_chain_weight", C.c_int),
        ("max_chain_extend", C.c_int),
        ("split_factor", C.c_float),
        ("split_width", C.c_int), ("max_occ", C.c_int),
        ("max_chain_gap", C.c_int), ("n_threads", C.c_int),
        ("chunk_size", C.c_int),
        ("mask_level", C.c_float), ("drop_ratio", C.c_float),
        ("XA_drop_ratio", C.c_float), ("mask_level_redun", C.c_float),
        ("mapQ_coef_len", C.c_float), ("mapQ_coef_fac", C.c_int),
        ("max_ins", C.c_int), ("max_matesw", C.c_int),
        ("max_XA_hits", C.c_int), ("max_XA_hits_alt", C.c_int),
        ("mat", C.c_int8 * 25),
    ]


class Contig(C.Structure):
    _fields_ = [
        ("offset", C.c_int64), ("len", C.c_int32), ("n_ambs", C.c_int32),
        ("gi", C.c_uint32), ("is_alt", C.c_int32),
        ("name", C.c_char_p), ("anno", C.c_char_p),
    ]


class IndexView(C.Structure):
    _fields_ = [
        ("primary", C.c_uint64), ("L2", C.c_uint64 * 5),
        ("seq_len", C.c_uint64), ("bwt_size", C.c_uint64),
        ("bwt", C.POINTER(C.c_uint32)),
        ("sa_intv", C.c_int), ("n_sa", C.c_uint64),
        ("sa", C.POINTER(C.c_uint64)),
        ("l_pac", C.c_int64), ("pac", C.POINTER(C.c_uint8)),
        ("n_seqs", C.c_int32), ("contigs", C.POINTER(Contig)),
    ]


HIT_DTYPE = np.dtype([
    ("rb", "<i8"), ("re", "<i8"), ("pos", "<i8"), ("hash", "<u8"),
    ("qb", "<i4"), ("qe", "<i4"), ("rid", "<i4"),
    ("score", "<i4"), ("truesc", "<i4"), ("sub", "<i4"), ("alt_sc", "<i4"),
    ("csub", "<i4"), ("sub_n", "<i4"), ("w", "<i4"), ("seedcov", "<i4"),
    ("secondary", "<i4"), ("secondary_all", "<i4"), ("seedlen0", "<i4"),
    ("n_comp", "<i4"), ("is_alt", "<i4"), ("frac_rep", "<f4"),
    ("flag", "<i4"), ("is_rev", "<i4"), ("mapq", "<i4"), ("NM", "<i4"),
    ("aln_sub", "<i4"), ("n_cigar", "<i4"), ("md_len", "<i4"),
    ("cigar_off", "<i8"), ("md_off", "<i8"),
], align=True)
assert HIT_DTYPE.itemsize == 144, HIT_DTYPE.itemsize


class ResultsView(C.Structure):
    _fields_ = [
        ("n_reads", C.c_int64),
        ("hit_off", C.POINTER(C.c_int64)),
        ("hits", C.c_void_p),
        ("cigar", C.POINTER(C.c_uint32)),
        ("md", C.POINTER(C.c_char)),
        ("n_hits", C.c_int64), ("n_cigar", C.c_int64), ("n_md", C.c_int64),
    ]


class StageStats(C.Structure):
    _fields_ = [
        ("ms_seed", C.c_float), ("ms_chain", C.c_float), ("ms_extend", C.c_float),
        ("ms_finalize", C.c_float), ("ms_total", C.c_float),
        ("occ_blocks", C.c_uint64), ("sa_reads", C.c_uint64), ("ref_bytes", C.c_uint64),
        ("sw_cells", C.c_uint64), ("n_ext", C.c_uint64), ("n_global", C.c_uint64),
        ("n_overflow", C.c_uint64), ("n_launches", C.c_int),
        ("tab_lookups_lo", C.c_uint64), ("tab_lookups_hi", C.c_uint64), ("ext_fallback", C.c_uint64), ("n_failed", C.c_uint64),
    ]


INTV_DTYPE = np.dtype([("x0", "<u8"), ("x1", "<u8"), ("x2", "<u8"), ("info", "<u8")])

EXT_JOB_DTYPE = np.dtype([
    ("qlen", "<i4"), ("tlen", "<i4"), ("q_off", "<i8"), ("t_off", "<i8"),
    ("w", "<i4"), ("end_bonus", "<i4"), ("zdrop", "<i4"), ("h0", "<i4"),
], align=True)
assert EXT_JOB_DTYPE.itemsize == 40
EXT_OUT_DTYPE = np.dtype([
    ("score", "<i4"), ("qle", "<i4"), ("tle", "<i4"), ("gtle", "<i4"),
    ("gscore", "<i4"), ("max_off", "<i4"),
])


def np_from_ptr(ptr, n, dtype):
    """Copy n items of dtype from a C pointer into a fresh numpy array."""
    dtype = np.dtype(dtype)
    if n == 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    addr = C.cast(ptr, C.c_void_p).value
    buf = (C.c_char * (n * dtype.itemsize)).from_address(addr)
    return np.frombuffer(buf, dtype=dtype, count=n).copy()


def pack_reads(reads):
    """list[str|bytes] -> (concatenated bytes as np.uint8, int64 offsets)."""
    bs = [r.encode() if isinstance(r, str) else bytes(r) for r in reads]
    off = np.zeros(len(bs) + 1, dtype=np.int64)
    if bs:
        off[1:] = np.cumsum([len(b) for b in bs])
    return np.frombuffer(b"".join(bs), dtype=np.uint8).copy(), off


class Results:
    """Host copy of a results view: hit_off, hits (structured), cigar, md."""

    def __init__(self, view):
        n = view.n_reads
        self.hit_off = np_from_ptr(view.hit_off, n + 1, np.int64)
        self.hits = np_from_ptr(view.hits, view.n_hits, HIT_DTYPE)
        self.cigar = np_from_ptr(view.cigar, view.n_cigar, np.uint32)
        self.md = np_from_ptr(view.md, view.n_md, np.uint8).tobytes()

    def read_hits(self, i):
        return self.hits[self.hit_off[i]:self.hit_off[i + 1]]

    def cigar_of(self, h):
        return self.cigar[h["cigar_off"]:h["cigar_off"] + h["n_cigar"]]

    def cigar_str(self, h):
        return "".join("%d%s" % (c >> 4, "MIDSH"[c & 0xf]) for c in self.cigar_of(h))

    def md_of(self, h):
        return self.md[h["md_off"]:h["md_off"] + h["md_len"]].decode()


# ---------------------------------------------------------------------------------------------------- fermi-lite half
class MagOpt(C.Structure):
    """b200_magopt_t == magopt_t (fermi-lite/fml.h:17-20)."""
    _fields_ = [(n, C.c_int) for n in ("flag", "min_ovlp", "min_elen", "min_ensr", "min_insr", "max_bdist", "max_bdiff",
                                        "max_bvtx", "min_merge_len", "trim_len", "trim_depth")] + \
               [(n, C.c_float) for n in ("min_dratio1", "max_bcov", "max_bfrac")]


class FmlOpt(C.Structure):
    """b200_fml_opt_t == fml_opt_t (fermi-lite/fml.h:22-29)."""
    _fields_ = [("n_threads", C.c_int), ("ec_k", C.c_int), ("min_cnt", C.c_int), ("max_cnt", C.c_int),
                ("min_asm_ovlp", C.c_int), ("min_merge_len", C.c_int), ("mag_opt", MagOpt)]


class Fseq1(C.Structure):
    """b200_fseq1_t == fseq1_t (fermi-lite/fml.h:8-11)."""
    _fields_ = [("l_seq", C.c_int32), ("seq", C.c_void_p), ("qual", C.c_void_p)]


class FmlStats(C.Structure):
    _fields_ = [("ms_count", C.c_float), ("ms_table", C.c_float), ("ms_ec", C.c_float), ("ms_flt", C.c_float),
                ("ms_total", C.c_float), ("ms_fmd", C.c_float), ("ms_nodes", C.c_float), ("ms_walk_host", C.c_float),
                ("ms_clean_host", C.c_float), ("fmd_symbols", C.c_uint64), ("n_strings", C.c_uint64), ("n_vertices", C.c_uint64),
                ("n_utg", C.c_uint64), ("n_kmers", C.c_uint64), ("n_distinct", C.c_uint64), ("table_bytes", C.c_uint64),
                ("n_lookups", C.c_uint64), ("n_spill", C.c_uint64), ("ec_codes", C.c_uint64 * 8), ("n_launches", C.c_int)]


def pack_reads_quals(reads, quals=None):
    """(seqs u8, quals u8 | None, off int64) for lists of str/bytes; quals=None: no qualities."""
    s, off = pack_reads(reads)
    if quals is None:
        return s, None, off
    q, qoff = pack_reads(quals)
    assert np.array_equal(off, qoff), "sequence / quality length mismatch"
    return s, q, off


def unpack_reads(pool, off, lens=None):
    """Pool + offsets (+ new lengths) -> list[bytes]."""
    b = pool.tobytes()
    out = []
    for i in range(len(off) - 1):
        ln = int(off[i + 1] - off[i]) if lens is None else int(lens[i])
        out.append(b[int(off[i]):int(off[i]) + ln])
    return out


class UtgOvlp(C.Structure):
    """b200_utg_ovlp_t == fml_ovlp_t (fermi-lite/fml.h:34-37)."""
    _fields_ = [("len", C.c_uint32, 31), ("from_", C.c_uint32, 1), ("id", C.c_uint32, 31), ("to", C.c_uint32, 1)]


class Utg(C.Structure):
    """b200_utg_t == fml_utg_t (fermi-lite/fml.h:39-46)."""
    _fields_ = [("len", C.c_int32), ("nsr", C.c_int32), ("seq", C.c_void_p), ("cov", C.c_void_p),
                ("n_ovlp", C.c_int * 2), ("ovlp", C.POINTER(UtgOvlp))]


def utgs_to_py(n, arr):
    """fml_utg_t array -> list of dicts (seq, cov, nsr, ovlp[(len, from, id, to)])."""
    out = []
    for i in range(n):
        u = arr[i]
        no = u.n_ovlp[0] + u.n_ovlp[1]
        out.append(dict(seq=C.string_at(u.seq, u.len), cov=C.string_at(u.cov, u.len), nsr=u.nsr, n_ovlp=(u.n_ovlp[0], u.n_ovlp[1]),
                        ovlp=[(u.ovlp[j].len, u.ovlp[j].from_, u.ovlp[j].id, u.ovlp[j].to) for j in range(no)]))
    return out
