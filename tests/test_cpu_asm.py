"""CPU-only suite of the assembly half: FMD rank blocks, per-string overlap records, the seed-order walk and the graph
cleaning, compiled for the host (tests/hostsim/fmd_emul.cpp), against golden vectors made by the reference's own fermi-lite
and against the live reference library where oracle/_ref exists.  No CUDA compute."""
import hashlib
import numpy as np
import pytest

import cases
import fmdmodel
import fmdsim
import fmlcases
from seqlib_b200.abi import FmlOpt, unpack_reads


def _opt():
    """fml_opt_init + mag_init_opt values (fermi-lite/misc.c:31-41, mag.c:539-557)."""
    o = FmlOpt()
    o.n_threads, o.ec_k, o.min_cnt, o.max_cnt, o.min_asm_ovlp, o.min_merge_len = 1, 0, 4, 8, 33, 0
    m = o.mag_opt
    m.flag, m.trim_len, m.trim_depth, m.min_elen, m.min_ovlp, m.min_merge_len, m.min_ensr, m.min_insr = 0xc0, 0, 6, 300, 0, 0, 4, 3
    m.min_dratio1, m.max_bcov, m.max_bfrac, m.max_bvtx, m.max_bdist, m.max_bdiff = 0.7, 10., 0.15, 64, 512, 50
    return o


def _ref_bwt(fs, foff):
    """The reference's BWT of the filtered reads (needs oracle/_ref); the fixture pins its digest."""
    from oracle import pyref_fml
    if not pyref_fml.have_ref():
        pytest.skip("oracle/_ref not built")
    return pyref_fml.bwt(fs, foff)[0]


def test_bwt_model_matches_reference_digest():
    """The sort-order model of the BWT (tests/fmdmodel.py, what fmd.cu implements) equals the reference's BWT: checked on a
    slice small enough for the brute-force model, against the live library."""
    from oracle import pyref_fml
    if not pyref_fml.have_ref():
        pytest.skip("oracle/_ref not built")
    seqs, quals, off = cases.fml_reads(120, region=400, read_len=40, seed=100, junk=0.05)
    reads = unpack_reads(seqs, off) + [b"ACGTACGT", b"AATT", b"ACGT", b"ACGT", b"GGATCC", b"ACGNT"]
    from seqlib_b200.abi import pack_reads
    s2, o2 = pack_reads(reads)
    rb = pyref_fml.bwt(s2, o2)[0]
    assert np.array_equal(rb, fmdmodel.bwt(reads))


@pytest.mark.parametrize("two_phase", [False, True])
@pytest.mark.parametrize("name", fmlcases.FML_SETS)
def test_rank_graph_and_cleaning_on_cpu_vs_golden(name, two_phase, monkeypatch):
    """Over the reference's BWT of the filtered reads: rld_rank1a answers, the graph out of fml_fmi2mag, and the graph after
    fml_mag_clean are identical, as mag_g_print text, to the committed reference output."""
    # two_phase: the extension-then-consume loops the lane-cooperative kernel runs (one thread standing in for the group)
    if two_phase:
        monkeypatch.setenv("FMD_EMUL_TWO_PHASE", "1")
    else:
        monkeypatch.delenv("FMD_EMUL_TWO_PHASE", raising=False)
    seqs, quals, off, z = fmlcases.load(name)
    gold = fmlcases.load_asm(name)
    fs, foff = fmlcases.filtered_reads(z, off)
    bwt = _ref_bwt(fs, foff)
    assert hashlib.md5(bwt.tobytes()).hexdigest() == gold["bwt_md5"]
    rr, rs = fmdsim.rank(bwt, gold["rank_q"])
    assert np.array_equal(rr, gold["rank_r"]) and np.array_equal(rs, gold["rank_s"])
    kcov = float(z["flt_kcov"])
    for stage, key in ((0, "mag0"), (1, "mag1")):
        o = fmlcases.asm_opt_for(_opt(), int(foff[-1]), len(foff) - 1, kcov, clean=stage >= 1)
        txt, rdist, _ = fmdsim.mag_text(bwt, o, stage)
        assert txt == gold[key]
        assert np.float32(rdist) == np.float32(gold["rdist"]) or (np.isnan(rdist) and np.isnan(gold["rdist"]))


def test_small_set_with_model_bwt_no_reference_needed():
    """Same pipeline with the BWT from the model (no oracle/_ref needed at all): fixture built from 300 reads."""
    seqs, quals, off, z = fmlcases.load("fml_mixed_noqual")
    gold = fmlcases.load_asm("fml_mixed_noqual")
    fs, foff = fmlcases.filtered_reads(z, off)
    # the brute-force model is quadratic: only check the digest on a prefix-sized problem when it is small enough
    if int(foff[-1]) > 250000:
        pytest.skip("set too large for the brute-force model")
    bwt = fmdmodel.bwt(unpack_reads(fs, foff))
    assert hashlib.md5(bwt.tobytes()).hexdigest() == gold["bwt_md5"]


@pytest.mark.parametrize("flag", [0xc0, 0x40, 0x00, 0x20])
def test_bubble_passes_vs_live_reference(flag):
    """Graph cleaning with every combination of MAG_F_NO_SIMPL / MAG_F_POPOPEN / MAG_F_AGGRESSIVE on a diploid read set (real
    bubbles): mag_g_simplify_bubble, mag_g_pop_simple (incl. the fork's additive-gap SW "score") and mag_g_pop_open leave
    the same graph as the reference."""
    from oracle import pyref_fml
    if not pyref_fml.have_ref():
        pytest.skip("oracle/_ref not built")
    seqs, quals, off = cases.fml_diploid_reads(3000, 4000, 3)
    exp = fmlcases.reference_pipeline(pyref_fml, seqs, quals, off)
    fs, foff = fmlcases.filtered_reads(exp, off)
    kcov = float(exp["flt_kcov"])
    bwt = pyref_fml.bwt(fs, foff)[0]
    ro = pyref_fml.default_opt()
    ro.mag_opt.flag = flag
    want, _, _ = pyref_fml.mag_text(ro, 1, kcov, fs, foff)
    o = fmlcases.asm_opt_for(_opt(), int(foff[-1]), len(foff) - 1, kcov)
    o.mag_opt.flag = flag
    got, _, _ = fmdsim.mag_text(bwt, o, 1)
    assert got == want
