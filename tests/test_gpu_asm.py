"""GPU parity tests of the assembly half through the C ABI: FMD-index (BWT, counts, ranks), the unitig graph before and after
cleaning, and the unitigs of the full fml_assemble pipeline, against golden vectors made by the reference's own fermi-lite C
and against the live reference library when oracle/_ref is present."""
import ctypes as C
import hashlib
import numpy as np
import pytest

import cases
import fmdmodel
import fmlcases
from seqlib_b200.abi import unpack_reads, pack_reads

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from seqlib_b200 import capi as c
    c.set_device(0)
    return c


@pytest.mark.parametrize("name", fmlcases.FML_SETS)
def test_fmd_index_vs_golden(capi, name):
    """b200_fmd_build == fml_seq2fmi: same BWT (digest), same cumulative counts, same rld_rank1a answers."""
    seqs, quals, off, z = fmlcases.load(name)
    gold = fmlcases.load_asm(name)
    fs, foff = fmlcases.filtered_reads(z, off)
    f = capi.Fmd(fs, foff)
    cnt, mcnt = f.info()
    assert np.array_equal(cnt, gold["cnt"])
    assert hashlib.md5(f.bwt().tobytes()).hexdigest() == gold["bwt_md5"]
    rr, rs = f.rank1a(gold["rank_q"])
    assert np.array_equal(rr, gold["rank_r"]) and np.array_equal(rs, gold["rank_s"])
    f.close()


def test_fmd_edge_cases_vs_model(capi):
    """Palindromes (shortened by one), duplicates, reads with N (skipped), very short reads, ragged lengths, empty input."""
    reads = [b"ACGTACGT", b"AATT", b"ACGT", b"ACGT", b"GGATCC", b"ACGNT", b"A", b"TTTTTTTT", b"acgtacgtaa", b"GATTACA" * 9]
    seqs, quals, off = cases.fml_reads(150, region=500, read_len=50, seed=77, junk=0.05)
    reads += unpack_reads(seqs, off)
    s2, o2 = pack_reads(reads)
    f = capi.Fmd(s2, o2)
    assert np.array_equal(f.bwt(), fmdmodel.bwt(reads))
    f.close()
    e = capi.Fmd(np.zeros(0, np.uint8), np.zeros(1, np.int64))
    assert len(e.bwt()) == 0
    e.close()
    s3, o3 = pack_reads([b"ACNNT", b"NNNN"])
    e = capi.Fmd(s3, o3)
    assert len(e.bwt()) == 0
    e.close()


@pytest.mark.parametrize("name", fmlcases.FML_SETS)
def test_unitig_graph_and_cleaning_vs_golden(capi, name):
    """b200_fml_mag_text stage 0 / 1 == fml_seq2fmi + fml_fmi2mag (+ fml_mag_clean), as mag_g_print text."""
    seqs, quals, off, z = fmlcases.load(name)
    gold = fmlcases.load_asm(name)
    fs, foff = fmlcases.filtered_reads(z, off)
    kcov = float(z["flt_kcov"])
    for stage, key in ((0, "mag0"), (1, "mag1")):
        o = fmlcases.asm_opt_for(capi.fml_default_opt(), int(foff[-1]), len(foff) - 1, kcov, clean=stage >= 1)
        txt, rdist = capi.fml_mag_text(o, stage, fs, foff)
        assert txt == gold[key]
        assert np.float32(rdist) == np.float32(gold["rdist"]) or (np.isnan(rdist) and np.isnan(gold["rdist"]))


@pytest.mark.parametrize("name", fmlcases.FML_SETS)
def test_assemble_vs_golden(capi, name):
    """b200_fml_assemble_flat == fml_assemble: identical unitig sequences, coverage strings, read counts and overlaps."""
    seqs, quals, off, z = fmlcases.load(name)
    gold = fmlcases.load_asm(name)
    utgs = capi.fml_assemble_flat(capi.fml_default_opt(), seqs, quals, off)
    assert fmlcases.utg_text(utgs) == gold["utg"]
    st = capi.fml_last_stats()
    assert st["n_launches"] > 0 and st["fmd_symbols"] > 0 and st["n_utg"] == len(utgs)


def test_assemble_vs_live_reference(capi):
    """Fresh inputs (not in any fixture), including an explicit ec_k, ec_k < 0 (no correction) and a batch that filters to
    nothing, against the reference library run here."""
    from oracle import pyref_fml
    if not pyref_fml.have_ref():
        pytest.skip("oracle/_ref not built")
    for n, region, seed, ec_k in ((2500, 6000, 31, 0), (1800, 3000, 32, 19), (1500, 4000, 33, -1)):
        seqs, quals, off = cases.fml_reads(n, region=region, seed=seed)
        ro = pyref_fml.default_opt()
        ro.ec_k = ec_k
        exp, _ = pyref_fml.assemble(ro, seqs, quals, off)
        o = capi.fml_default_opt()
        o.ec_k = ec_k
        got = capi.fml_assemble_flat(o, seqs, quals, off)
        assert fmlcases.utg_text(got) == fmlcases.utg_text(exp)
    rng = np.random.default_rng(5)
    junk = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, 40 * 100)]
    joff = np.arange(41, dtype=np.int64) * 100
    assert capi.fml_assemble_flat(capi.fml_default_opt(), junk, None, joff) == []
    assert capi.fml_assemble_flat(capi.fml_default_opt(), np.zeros(0, np.uint8), None, np.zeros(1, np.int64)) == []


@pytest.mark.parametrize("n_threads", [1, 6])
def test_assemble_windows_equals_one_window_at_a_time(capi, n_threads):
    """b200_fml_assemble_windows: every window gets the unitigs b200_fml_assemble_flat gives for its reads alone (golden sets, an
    empty window, a window that filters to nothing), whatever the number of host threads / streams."""
    sets = [fmlcases.load(name) for name in fmlcases.FML_SETS if fmlcases.load(name)[1] is not None]
    rng = np.random.default_rng(9)
    junk = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, 30 * 100)]
    wins = []
    for rep in range(2):
        for seqs, quals, off, z in sets:
            wins.append((np.asarray(seqs, dtype=np.uint8), np.asarray(quals, dtype=np.uint8), np.asarray(off, dtype=np.int64)))
        wins.append((junk, np.full(len(junk), ord("I"), dtype=np.uint8), np.arange(31, dtype=np.int64) * 100))
        wins.append((np.zeros(0, np.uint8), np.zeros(0, np.uint8), np.zeros(1, np.int64)))
    seqs = np.concatenate([w[0] for w in wins]); quals = np.concatenate([w[1] for w in wins])
    off = [0]; win_off = [0]
    for s, q, o in wins:
        base = off[-1]
        off += [base + int(x) for x in o[1:]]
        win_off.append(len(off) - 1)
    opt = capi.fml_default_opt()
    got = capi.fml_assemble_windows(opt, seqs, quals, np.array(off, dtype=np.int64), np.array(win_off, dtype=np.int64), n_threads)
    assert len(got) == len(wins)
    for (s, q, o), g in zip(wins, got):
        exp = capi.fml_assemble_flat(opt, s, q, o) if len(o) > 1 else []
        assert fmlcases.utg_text(g) == fmlcases.utg_text(exp)
    assert capi.fml_last_stats()["n_launches"] >= 0


def test_direct_assemble_and_fseq_entry(capi):
    """b200_fml_seqs2utg_flat (FermiAssembler::DirectAssemble's path) and the fseq1_t form b200_fml_assemble."""
    from oracle import pyref_fml
    from seqlib_b200.abi import Fseq1, Utg, utgs_to_py
    if not pyref_fml.have_ref():
        pytest.skip("oracle/_ref not built")
    seqs, quals, off, z = fmlcases.load("fml_mt_2k")
    gold = fmlcases.load_asm("fml_mt_2k")
    fs, foff = fmlcases.filtered_reads(z, off)
    o = fmlcases.asm_opt_for(capi.fml_default_opt(), int(foff[-1]), len(foff) - 1, float(z["flt_kcov"]))
    got = capi.fml_seqs2utg_flat(o, fs, foff)
    assert fmlcases.utg_text(got) == gold["utg"]
    n = len(off) - 1
    arr = (Fseq1 * n)()
    keep = []
    for i in range(n):
        s = C.create_string_buffer(seqs[int(off[i]):int(off[i + 1])].tobytes())
        q = C.create_string_buffer(quals[int(off[i]):int(off[i + 1])].tobytes())
        keep += [s, q]
        arr[i].l_seq = int(off[i + 1] - off[i])
        arr[i].seq = C.cast(s, C.c_void_p)
        arr[i].qual = C.cast(q, C.c_void_p)
    nu = C.c_int(0)
    up = C.POINTER(Utg)()
    o2 = capi.fml_default_opt()
    assert capi.lib().b200_fml_assemble(C.byref(o2), n, arr, C.byref(nu), C.byref(up)) == 0
    assert fmlcases.utg_text(utgs_to_py(nu.value, up)) == gold["utg"]
    capi.lib().b200_fml_utg_destroy(nu.value, up)


def _cxx_exe(name):
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tests", "cxx", name)
    src = exe + ".cpp"
    lib = os.path.join(root, "seqlib_b200", "libSeqLibB200.so")
    subprocess.check_call(["make", "-s", "-C", os.path.join(root, "seqlib_b200", "cxx")])
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(src), os.path.getmtime(lib)):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-I" + os.path.join(root, "include"), "-o", exe, src,
                               "-L" + os.path.join(root, "seqlib_b200"), "-lSeqLibB200", "-lseqlib_b200",
                               "-Wl,-rpath," + os.path.join(root, "seqlib_b200")])
    return exe


def test_cxx_fermi_assembler_and_bfc(capi, tmp_path):
    """SeqLib::FermiAssembler::PerformAssembly through the C++ drop-in class gives the golden unitigs; the BFC ->
    DirectAssemble flow of the reference's "correct_and_assemble" test runs and agrees with the flat ABI."""
    import subprocess
    exe = _cxx_exe("test_fermi")
    seqs, quals, off, z = fmlcases.load("fml_mt_2k")
    gold = fmlcases.load_asm("fml_mt_2k")
    rs, rq = unpack_reads(seqs, off), unpack_reads(quals, off)
    tsv = tmp_path / "reads.tsv"
    with open(tsv, "w") as f:
        for i, (a, b) in enumerate(zip(rs, rq)):
            f.write("r%d\t%s\t%s\n" % (i, a.decode(), b.decode()))
    r = subprocess.run([exe, str(tsv), "assemble"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    want = [ln for ln in gold["utg"].split("\n")[1::4]]
    assert r.stdout.split() == want
    r = subprocess.run([exe, str(tsv), "bfc"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    contigs = r.stdout.split()
    assert contigs and max(len(c) for c in contigs) > 1000
    # the same flow on the flat ABI: count + correct with the trained k, then the assembly half with DirectAssemble's options
    kmer = int(r.stderr.split("kmer ")[1].split()[0])
    o = capi.fml_default_opt()
    o.ec_k = kmer
    cs, cq, _, kcov = capi.fml_correct_flat(o, seqs, quals, off)
    up = np.frombuffer(cs.tobytes().upper(), dtype=np.uint8)
    o2 = capi.fml_default_opt()
    me = o2.mag_opt.min_ensr if o2.mag_opt.min_ensr > kcov * .1 else int(kcov * .1 + .499)
    o2.mag_opt.min_ensr, o2.mag_opt.min_insr = me, me - 1
    flat = capi.fml_seqs2utg_flat(o2, up, off)
    assert [u["seq"].decode() for u in flat] == contigs


@pytest.mark.parametrize("flag", [0xc0, 0x40, 0x20])
def test_bubble_flags_vs_live_reference(capi, flag):
    """FermiAssembler::SetSimplifyBubble / SetAggressiveTrim paths (MAG_F_NO_SIMPL cleared, MAG_F_AGGRESSIVE set) on a diploid
    read set: the cleaned graph equals the reference's."""
    from oracle import pyref_fml
    if not pyref_fml.have_ref():
        pytest.skip("oracle/_ref not built")
    seqs, quals, off = cases.fml_diploid_reads(3000, 4000, 3)
    exp = fmlcases.reference_pipeline(pyref_fml, seqs, quals, off)
    fs, foff = fmlcases.filtered_reads(exp, off)
    kcov = float(exp["flt_kcov"])
    ro = pyref_fml.default_opt()
    ro.mag_opt.flag = flag
    want, _, _ = pyref_fml.mag_text(ro, 1, kcov, fs, foff)
    o = fmlcases.asm_opt_for(capi.fml_default_opt(), int(foff[-1]), len(foff) - 1, kcov)
    o.mag_opt.flag = flag
    got, _ = capi.fml_mag_text(o, 1, fs, foff)
    assert got == want
    # and the whole pipeline with that flag
    ro2 = pyref_fml.default_opt()
    ro2.mag_opt.flag = flag
    o2 = capi.fml_default_opt()
    o2.mag_opt.flag = flag
    e, _ = pyref_fml.assemble(ro2, seqs, quals, off)
    assert fmlcases.utg_text(capi.fml_assemble_flat(o2, seqs, quals, off)) == fmlcases.utg_text(e)
