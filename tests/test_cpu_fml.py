"""CPU-only suite of the fermi-lite half: golden vectors (made by the reference's own fermi-lite C) against the host build
of the per-read device code (tests/hostsim/fml_emul.cpp), and against the live reference library where oracle/_ref exists.
No CUDA compute."""
import ctypes as C
import numpy as np
import pytest

import cases
import fmlcases
import fmlsim
from seqlib_b200.abi import FmlOpt


def _emul_opt():
    """fml_opt_init values (fermi-lite/misc.c:31-41); the emulation has no init entry point of its own."""
    o = FmlOpt()
    o.n_threads, o.ec_k, o.min_cnt, o.max_cnt, o.min_asm_ovlp, o.min_merge_len = 1, 0, 4, 8, 33, 0
    return o


def _emul_correct(opt, seqs, quals, off, flt_uniq=False):
    return fmlsim.correct_flat(opt, seqs, quals, off, flt_uniq=flt_uniq)[:4]


@pytest.mark.parametrize("name", fmlcases.FML_SETS)
def test_bfc_device_code_on_cpu_vs_golden(name):
    """fml_correct then fml_fltuniq (the BFC stages of fml_assemble) reproduce the reference's output bit for bit:
    corrected bases (lower case = changed), recoded qualities, kcov, kept run per read."""
    seqs, quals, off, z = fmlcases.load(name)
    got = fmlcases.pipeline(_emul_correct, _emul_opt(), seqs, quals, off)
    assert fmlcases.compare(got, z) == []


@pytest.mark.parametrize("name", fmlcases.FML_SETS)
def test_c_restatement_vs_golden(name):
    """oracle/oracle_fml.c (plain-C restatement of fml_count / bfc_ec1 / fml_fltuniq, sorted-array count table) reproduces
    the reference's committed output: this pins the oracle that checks the CUDA path where oracle/_ref is absent."""
    from oracle import pyoracle
    seqs, quals, off, z = fmlcases.load(name)
    got = fmlcases.pipeline(lambda *a, **k: pyoracle.fml_correct_flat(*a, **k)[:4], pyoracle.fml_default_opt(), seqs, quals, off)
    assert fmlcases.compare(got, z) == []


def test_c_restatement_vs_live_reference():
    from oracle import pyoracle, pyref_fml
    if not pyref_fml.have_ref():
        pytest.skip("oracle/_ref not built")
    seqs, quals, off = cases.fml_reads(700, region=1800, seed=12)
    for ec_k in (15, 23, 33):
        o = pyoracle.fml_default_opt()
        o.ec_k = ec_k
        ro = pyref_fml.default_opt()
        ro.ec_k = ec_k
        for qq in (quals, None):
            for flt in (False, True):
                r = pyref_fml.correct_flat(ro, seqs, qq, off, flt_uniq=flt)
                e = pyoracle.fml_correct_flat(o, seqs, qq, off, flt_uniq=flt)
                assert np.array_equal(r[2], e[2]) and r[3] == e[3]
                for i in range(len(off) - 1):
                    a, b = int(off[i]), int(off[i]) + int(r[2][i])
                    assert np.array_equal(r[0][a:b], e[0][a:b])
                    if qq is not None:
                        assert np.array_equal(r[1][a:b], e[1][a:b])


def test_golden_matches_live_reference():
    """The committed fixtures are what the reference library produces here (pins the oracle)."""
    from oracle import pyref_fml
    if not pyref_fml.have_ref():
        pytest.skip("oracle/_ref not built")
    seqs, quals, off, z = fmlcases.load("fml_mt_2k")
    got = fmlcases.reference_pipeline(pyref_fml, seqs, quals, off)
    assert fmlcases.compare(got, z) == []
    o = pyref_fml.default_opt()
    e = _emul_opt()
    e.mag_opt = o.mag_opt
    assert bytes(o) == bytes(e)


@pytest.mark.parametrize("k,l_pre", [(11, 20), (21, 20), (31, 20), (32, 20), (33, 20), (41, 12), (63, 20)])
def test_count_histogram_vs_reference(k, l_pre):
    """worker_count + bfc_ch_insert + bfc_ch_hist for k on both sides of the 32-bit split of get_subhash
    (fermi-lite/htab.c:45-58): same number of distinct keys, same total / high-quality histograms, same mode."""
    from oracle import pyref_fml
    if not pyref_fml.have_ref():
        pytest.skip("oracle/_ref not built")
    seqs, quals, off = cases.fml_reads(800, region=2500, seed=5)
    rc, rh, rmode, rnd = pyref_fml.count_hist(seqs, quals, off, k, 20, l_pre)
    ec, eh, emode, end_ = fmlsim.count_hist(seqs, quals, off, k, 20, l_pre)
    assert rnd == end_ and rmode == emode
    assert np.array_equal(rc, ec) and np.array_equal(rh, eh)


def test_edge_cases_vs_reference():
    """Empty batch, k <= 0, explicit ec_k, reads without qualities, a batch where nothing is solid."""
    from oracle import pyref_fml
    if not pyref_fml.have_ref():
        pytest.skip("oracle/_ref not built")
    o = _emul_opt()
    off0 = np.zeros(1, dtype=np.int64)
    s, q, l, kcov = _emul_correct(o, np.zeros(0, np.uint8), None, off0)
    assert kcov == 255.0 and len(l) == 0
    seqs, quals, off = cases.fml_reads(600, region=1500, seed=9)
    for ec_k in (13, 19, 27):
        o.ec_k = ec_k
        ro = pyref_fml.default_opt()
        ro.ec_k = ec_k
        for qq in (quals, None):
            for flt in (False, True):
                r = pyref_fml.correct_flat(ro, seqs, qq, off, flt_uniq=flt)
                e = _emul_correct(o, seqs, qq, off, flt_uniq=flt)
                assert np.array_equal(r[2], e[2])
                assert r[3] == e[3]
                for i in range(len(off) - 1):
                    a, b = int(off[i]), int(off[i]) + int(r[2][i])
                    assert np.array_equal(r[0][a:b], e[0][a:b])
                    if qq is not None:
                        assert np.array_equal(r[1][a:b], e[1][a:b])
    # all reads unrelated: no k-mer reaches min_cnt, kcov is NaN on both sides and nothing changes
    rng = np.random.default_rng(3)
    n = 50
    seqs = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, n * 100)]
    off = np.arange(n + 1, dtype=np.int64) * 100
    o.ec_k = 21
    ro = pyref_fml.default_opt()
    ro.ec_k = 21
    r = pyref_fml.correct_flat(ro, seqs, None, off)
    e = _emul_correct(o, seqs, None, off)
    assert np.isnan(r[3]) and np.isnan(e[3])
    assert np.array_equal(r[0], e[0])


def test_product_fails_loudly_without_a_device():
    """No CPU fallback: on a box without a CUDA device the fermi entry points return B200_ERR_CUDA (the k <= 0 no-op, which does
    no device work at all, is the only call that succeeds)."""
    import os
    from seqlib_b200 import capi
    if not os.path.exists(capi.SO_PATH):
        pytest.skip("libseqlib_b200.so not built")
    if capi.lib().b200_device_count() > 0:
        pytest.skip("a CUDA device is present")
    seqs, quals, off = cases.fml_reads(50, region=500, seed=2)
    o = capi.fml_default_opt()
    s, q, l, kcov = capi.fml_correct_flat(o, seqs, quals, off)          # ec_k == 0: reference no-op
    assert kcov == 255.0 and np.array_equal(s, seqs)
    o.ec_k = 17
    with pytest.raises(capi.B200Error, match="-2"):
        capi.fml_correct_flat(o, seqs, quals, off)
    with pytest.raises(capi.B200Error, match="-2"):
        capi.fml_assemble_flat(capi.fml_default_opt(), seqs, quals, off)
    with pytest.raises(capi.B200Error, match="-2"):
        capi.Fmd(seqs, off)
    with pytest.raises(capi.B200Error, match="-2"):
        capi.fml_assemble_windows(capi.fml_default_opt(), seqs, quals, off, np.array([0, 25, 50], dtype=np.int64), 2)
    # argument checks come before any device work
    with pytest.raises(capi.B200Error, match="-1"):
        capi.fml_assemble_windows(capi.fml_default_opt(), seqs, quals, off, np.array([0, 30, 20], dtype=np.int64), 2)
