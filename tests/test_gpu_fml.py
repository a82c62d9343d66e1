"""GPU parity tests of the fermi-lite half: the CUDA path, called through the C ABI, against (a) the committed golden
vectors produced by the reference's own fermi-lite C and (b) the live reference library when oracle/_ref is present."""
import ctypes as C
import numpy as np
import pytest

import cases
import fmlcases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from seqlib_b200 import capi as c
    c.set_device(0)
    return c


def _ref():
    from oracle import pyref_fml
    if not pyref_fml.have_ref():
        pytest.skip("oracle/_ref not built")
    return pyref_fml


@pytest.mark.parametrize("name", fmlcases.FML_SETS)
def test_correct_and_fltuniq_vs_golden(capi, name):
    """b200_fml_correct_flat twice (correct, then unique filter on the corrected reads) == fml_correct + fml_fltuniq."""
    seqs, quals, off, z = fmlcases.load(name)
    got = fmlcases.pipeline(capi.fml_correct_flat, capi.fml_default_opt(), seqs, quals, off)
    assert fmlcases.compare(got, z) == []
    st = capi.fml_last_stats()
    assert st["n_launches"] > 0 and st["n_kmers"] > 0


def test_opt_init_matches_reference(capi):
    ref = _ref()
    assert bytes(ref.default_opt()) == bytes(capi.fml_default_opt())


@pytest.mark.parametrize("k,l_pre", [(11, 20), (21, 20), (31, 20), (32, 20), (33, 20), (41, 12), (63, 20)])
def test_count_table_vs_reference(capi, k, l_pre):
    """b200_fml_count + b200_kmer_table_hist/_size/_lookup == fml_count + bfc_ch_hist / bfc_ch_count / bfc_ch_kmer_occ."""
    ref = _ref()
    seqs, quals, off = cases.fml_reads(800, region=2500, seed=5)
    rc, rh, rmode, rnd = ref.count_hist(seqs, quals, off, k, 20, l_pre)
    tab = capi.KmerTable(seqs, quals, off, k, 20, l_pre)
    cnt, high, mode = tab.hist()
    assert tab.size() == rnd and mode == rmode
    assert np.array_equal(cnt, rc) and np.array_equal(high, rh)
    # probe: k-mers of the reads themselves (present), mutated ones (mostly absent), ones with N
    rng = np.random.default_rng(k)
    b = seqs.tobytes().upper()
    kmers = []
    for _ in range(400):
        i = int(rng.integers(0, len(off) - 1))
        ln = int(off[i + 1] - off[i])
        if ln < k:
            continue
        p = int(off[i]) + int(rng.integers(0, ln - k + 1))
        x = bytearray(b[p:p + k])
        u = rng.random()
        if u < 0.3:
            x[int(rng.integers(0, k))] = b"ACGT"[int(rng.integers(0, 4))]
        elif u < 0.35:
            x[int(rng.integers(0, k))] = ord("N")
        kmers.append(bytes(x))
    assert np.array_equal(tab.lookup(kmers), ref.kmer_occ(seqs, quals, off, k, kmers, 20, l_pre))
    tab.close()


def test_edge_cases_vs_reference(capi):
    ref = _ref()
    o = capi.fml_default_opt()
    off0 = np.zeros(1, dtype=np.int64)
    s, q, l, kcov = capi.fml_correct_flat(o, np.zeros(0, np.uint8), None, off0)
    assert kcov == 255.0 and len(l) == 0
    o.ec_k = 17
    s, q, l, kcov = capi.fml_correct_flat(o, np.zeros(0, np.uint8), None, off0)        # empty batch, real k
    assert len(l) == 0
    seqs, quals, off = cases.fml_reads(600, region=1500, seed=9)
    for ec_k in (13, 19, 27):
        o.ec_k = ec_k
        ro = ref.default_opt()
        ro.ec_k = ec_k
        for qq in (quals, None):
            for flt in (False, True):
                r = ref.correct_flat(ro, seqs, qq, off, flt_uniq=flt)
                e = capi.fml_correct_flat(o, seqs, qq, off, flt_uniq=flt)
                assert np.array_equal(r[2], e[2])
                assert r[3] == e[3]
                for i in range(len(off) - 1):
                    a, b = int(off[i]), int(off[i]) + int(r[2][i])
                    assert np.array_equal(r[0][a:b], e[0][a:b])
                    if qq is not None:
                        assert np.array_equal(r[1][a:b], e[1][a:b])
    rng = np.random.default_rng(3)
    n = 50
    seqs = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, n * 100)]
    off = np.arange(n + 1, dtype=np.int64) * 100
    o.ec_k = 21
    e = capi.fml_correct_flat(o, seqs, None, off)
    assert np.isnan(e[3]) and np.array_equal(e[0], seqs)


def test_long_and_scratch_heavy_reads(capi):
    """Reads longer than the main-pass scratch (512) and a low-complexity batch that inflates the search stack go through
    the spill pass and still match the reference."""
    ref = _ref()
    seqs, quals, off = cases.fml_reads(300, region=4000, read_len=700, seed=21, junk=0.0)
    o = capi.fml_default_opt()
    o.ec_k = 19
    ro = ref.default_opt()
    ro.ec_k = 19
    r = ref.correct_flat(ro, seqs, quals, off)
    e = capi.fml_correct_flat(o, seqs, quals, off)
    assert np.array_equal(r[0], e[0]) and np.array_equal(r[1], e[1]) and r[3] == e[3]
    assert capi.fml_last_stats()["n_spill"] > 0


def test_fseq_entry_points(capi):
    """b200_fml_correct / b200_fml_fltuniq on fseq1_t arrays: in-place rewrite, dropped reads freed and zeroed."""
    from seqlib_b200.abi import Fseq1
    ref = _ref()
    seqs, quals, off = cases.fml_reads(500, region=1500, seed=4)
    n = len(off) - 1
    libc = C.CDLL(None)
    libc.malloc.restype = C.c_void_p
    libc.malloc.argtypes = [C.c_size_t]
    libc.free.argtypes = [C.c_void_p]
    arr = (Fseq1 * n)()
    for i in range(n):
        ln = int(off[i + 1] - off[i])
        arr[i].l_seq = ln
        for name, pool in (("seq", seqs), ("qual", quals)):
            p = libc.malloc(ln + 1)
            C.memmove(p, pool[int(off[i]):int(off[i + 1])].tobytes() + b"\0", ln + 1)
            setattr(arr[i], name, p)
    o = capi.fml_default_opt()
    o.ec_k = 15
    ro = ref.default_opt()
    ro.ec_k = 15
    kcov = C.c_float(0)
    assert capi.lib().b200_fml_correct(C.byref(o), n, arr, C.byref(kcov)) == 0
    r = ref.correct_flat(ro, seqs, quals, off)
    assert kcov.value == r[3]
    for i in range(n):
        assert C.string_at(arr[i].seq) == r[0][int(off[i]):int(off[i + 1])].tobytes()
        assert C.string_at(arr[i].qual) == r[1][int(off[i]):int(off[i + 1])].tobytes()
    assert capi.lib().b200_fml_fltuniq(C.byref(o), n, arr, C.byref(kcov)) == 0
    r2 = ref.correct_flat(ro, r[0], r[1], off, flt_uniq=True)
    assert kcov.value == r2[3]
    for i in range(n):
        ln = int(r2[2][i])
        assert arr[i].l_seq == ln
        if ln == 0:
            assert not arr[i].seq and not arr[i].qual
        else:
            assert C.string_at(arr[i].seq) == r2[0][int(off[i]):int(off[i]) + ln].tobytes()
            libc.free(arr[i].seq)
            libc.free(arr[i].qual)
