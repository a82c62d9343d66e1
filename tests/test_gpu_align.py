"""GPU parity tests: the CUDA path, called through the C ABI, against (a) the committed golden vectors
produced by the reference's own bwa C and (b) the live reference library when oracle/_ref is present."""
import ctypes
import filecmp
import os
import numpy as np
import pytest

import cases
import goldenlib
import parity

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from seqlib_b200 import capi as c
    c.set_device(0)
    return c


def _libc():
    return ctypes.CDLL(None)


def test_kat_construct_and_align(capi):
    """seq_test/seq_test.cpp:848-911: 4 inline contigs (one with 100 Ns), two queries."""
    gold, z = goldenlib.load("kat")
    _libc().srand48(cases.KAT_SRAND)
    idx = capi.Index.construct(cases.KAT_NAMES, cases.KAT_SEQS)
    a = idx.arrays()
    assert a["seq_len"] == 646 and a["primary"] == int(z["primary"])
    assert list(a["L2"]) == list(z["L2"])
    assert np.array_equal(a["bwt"], z["bwt"])
    assert np.array_equal(a["sa"], z["sa"])
    assert np.array_equal(a["pac"][:len(z["pac"])], z["pac"])
    assert idx.n_seqs() == 4 and idx.seq_name(2) == "ref5" and idx.seq_name(4) is None
    got = capi.align(idx, cases.KAT_QUERIES, capi.default_opt(), z["ids"])
    assert parity.compare_results(got, gold) == []
    h = got.read_hits(0)
    assert len(h) == 2 and got.cigar_str(h[0]) == "38M"
    assert len(got.read_hits(1)) == 2


def test_config1_10kb(capi):
    """BASELINE config 1: 1k x 150 bp synthetic reads vs a 10 kb in-memory ConstructIndex, bit-exact."""
    gold, z = goldenlib.load("c1")
    pac, ctg, ref_ascii = cases.c1_reference()
    idx = capi.Index.construct(["ref10k"], [ref_ascii])
    a = idx.arrays()
    assert a["primary"] == int(z["primary"]) and list(a["L2"]) == list(z["L2"])
    assert np.array_equal(a["bwt"], z["bwt"]) and np.array_equal(a["sa"], z["sa"])
    seqs, off = cases.c1_reads(pac, ctg)
    got = capi.align(idx, (seqs, off), capi.default_opt(), cases.ids_for(len(off) - 1))
    assert parity.compare_results(got, gold) == []
    st = capi.last_stats()
    assert st["n_launches"] > 0 and st["occ_blocks"] > 0


@pytest.mark.parametrize("name", ["sim1_5k", "bcr_2k"])
def test_tiny_fa_reads(capi, name):
    """The reference's shipped tiny.fa index + its wgsim reads (repeats, split reads, soft clips)."""
    gold, z = goldenlib.load(name)
    idx = capi.Index.load(goldenlib.path("tiny", "tiny.fa"))
    reads = cases.read_lines(goldenlib.path(name + ".txt"))
    ioff, intv = capi.collect_intv(idx, reads, capi.default_opt())
    assert np.array_equal(ioff, z["intv_off"])
    assert np.array_equal(intv, z["intv"])
    got = capi.align(idx, reads, capi.default_opt(), cases.ids_for(len(reads)))
    assert parity.compare_results(got, gold) == []


def test_small_chunks_and_spill(capi, monkeypatch):
    """Chunked execution (B200_CHUNK) must not change results."""
    gold, z = goldenlib.load("bcr_2k")
    idx = capi.Index.load(goldenlib.path("tiny", "tiny.fa"))
    reads = cases.read_lines(goldenlib.path("bcr_2k.txt"))
    monkeypatch.setenv("B200_CHUNK", "1024")
    got = capi.align(idx, reads, capi.default_opt(), cases.ids_for(len(reads)))
    assert parity.compare_results(got, gold) == []


def test_index_write_load_roundtrip(capi, tmp_path):
    """WriteIndex/LoadIndex on-disk format: loading tiny.fa and writing it back is byte-identical."""
    idx = capi.Index.load(goldenlib.path("tiny", "tiny.fa"))
    out = str(tmp_path / "rt")
    idx.write(out)
    for ext in ("bwt", "sa", "pac", "amb"):
        assert filecmp.cmp(out + "." + ext, goldenlib.path("tiny", "tiny.fa." + ext), shallow=False), ext
    # bns_restore_core turns the " (null)" annotation into "" (bwa/bntseq.c:124-126) and bns_dump then omits it (:76-78),
    # so the reference's own Load->Write round trip drops it too
    exp_ann = open(goldenlib.path("tiny", "tiny.fa.ann")).read().replace(" (null)", "")
    assert open(out + ".ann").read() == exp_ann
    from oracle import pyref
    if pyref.have_ref():
        ref_out = str(tmp_path / "ref")
        pyref.RefIndex.load(goldenlib.path("tiny", "tiny.fa")).write(ref_out)
        for ext in ("bwt", "sa", "pac", "ann", "amb"):
            assert filecmp.cmp(out + "." + ext, ref_out + "." + ext, shallow=False), ext
    idx2 = capi.Index.load(out)
    assert idx2.n_seqs() == idx.n_seqs() and idx2.l_pac() == idx.l_pac()


def test_gpu_builder_reproduces_tiny_index(capi):
    """The GPU suffix sort must reproduce the reference's shipped tiny.fa.bwt/.sa from the pac alone (real genome, repeats)."""
    idx = capi.Index.load(goldenlib.path("tiny", "tiny.fa"))
    a = idx.arrays()
    built = capi.Index.construct_pac(a["pac"], a["l_pac"], a["contigs"], keep_host=True)
    b = built.arrays()
    assert b["primary"] == a["primary"] and list(b["L2"]) == list(a["L2"])
    assert np.array_equal(b["bwt"], a["bwt"])
    assert np.array_equal(b["sa"], a["sa"])


@pytest.mark.parametrize("G", [0, 4, 8, 16, 32])
def test_ksw_extend2_batch(capi, monkeypatch, G):
    """ksw_extend2 tuples: scalar kernel (G=0) and the lane-cooperative kernels vs the reference's outputs."""
    monkeypatch.setenv("B200_KSW_G", str(G))
    z = np.load(goldenlib.path("ksw_c3.npz"))
    jobs, qp, tp = cases.c3_tuples(4000)
    opt = capi.default_opt()
    mat = np.array(list(opt.mat), dtype=np.int8)
    out, cells, ms = capi.ksw_extend2_batch(jobs, qp, tp, mat)
    for f in out.dtype.names:
        assert np.array_equal(out[f], z["out"][f]), f
    assert cells > 0
    # a harsher mix (short/unrelated/N-containing queries, narrow and wide bands, z-drop on/off) against the live reference
    from oracle import pyref
    if pyref.have_ref():
        jobs2, qp2, tp2 = cases.mixed_tuples(30000)
        exp, _ = pyref.ksw_extend2_batch(jobs2, qp2, tp2, mat, n_threads=8)
        got, cells2, _ = capi.ksw_extend2_batch(jobs2, qp2, tp2, mat)
        for f in got.dtype.names:
            assert np.array_equal(got[f], exp[f]), f


def test_against_live_reference_with_indels(capi):
    """Live oracle/_ref (the reference's bwa compiled in place): 200 kb random reference, reads with subs + indels + Ns."""
    from oracle import pyref
    if not pyref.have_ref():
        pytest.skip("oracle/_ref not built")
    from seqlib_b200 import synth
    l_pac = 200000
    pac = synth.reference(l_pac, seed=77)
    ctg = synth.contigs_for(l_pac, 3, "c")
    names = [c[0] for c in ctg]
    seqs = [synth.ascii_of(pac, c[1], c[1] + c[2]) for c in ctg]
    idx = capi.Index.construct(names, seqs)
    ridx = pyref.RefIndex.construct(names, seqs)
    a, b = idx.arrays(), ridx.arrays()
    assert np.array_equal(a["bwt"], b["bwt"]) and np.array_equal(a["sa"], b["sa"]) and a["primary"] == b["primary"]
    r, off, _, _ = synth.reads(pac, l_pac, ctg, 20000, 150, 0.02, 2e-3, seed=99)
    r = r.copy()
    r[::997] = ord("N")                      # sprinkle ambiguous bases
    ids = cases.ids_for(20000)
    opt = capi.default_opt()
    got = capi.align(idx, (r, off), opt, ids)
    exp, _ = pyref.align(ridx, (r, off), opt, ids, n_threads=4)
    assert parity.compare_results(got, exp) == []


def test_scale_properties(capi):
    """Size-independent properties on a larger run: idempotence, truth recovery, CIGAR/NM consistency."""
    from seqlib_b200 import synth
    l_pac = 4_000_000
    pac = synth.reference(l_pac, seed=5)
    ctg = synth.contigs_for(l_pac, 4)
    idx = capi.Index.construct_pac(pac, l_pac, ctg)
    n = 200000
    r, off, pos, strand = synth.reads(pac, l_pac, ctg, n, 150, 0.01, 0.0, seed=6)
    ids = cases.ids_for(n)
    opt = capi.default_opt()
    a = capi.align(idx, (r, off), opt, ids)
    b = capi.align(idx, (r, off), opt, ids)
    assert np.array_equal(a.hit_off, b.hit_off) and a.hits.tobytes() == b.hits.tobytes() and np.array_equal(a.cigar, b.cigar)
    first = a.hits[a.hit_off[:-1][np.diff(a.hit_off) > 0]]
    coff = np.array([c[1] for c in ctg], dtype=np.int64)
    has = np.diff(a.hit_off) > 0
    assert has.mean() > 0.999
    gpos = coff[first["rid"]] + first["pos"]
    ok = (np.abs(gpos - pos[has]) <= 8) & (first["is_rev"] == strand[has])
    assert ok.mean() > 0.995
    # query span of every hit equals the M/I/S(op3) lengths of its CIGAR
    ops = a.cigar & 0xf
    lens = (a.cigar >> 4).astype(np.int64)
    qlen = np.where((ops == 0) | (ops == 1) | (ops == 3), lens, 0)
    csum = np.concatenate([[0], np.cumsum(qlen)])
    per_hit = csum[a.hits["cigar_off"] + a.hits["n_cigar"]] - csum[a.hits["cigar_off"]]
    assert np.all(per_hit == 150)


def test_long_reads_vs_live_reference(capi):
    """Contig-like queries (0.8-4 kb, chimeras, indels): mem_flt_chained_seeds / mem_seed_sw (bwa/bwamem.c:597-641) is active,
    every read goes through the large-slot spill pass; hits, CIGARs, MD and MAPQ equal the reference library's."""
    from oracle import pyref
    if not pyref.have_ref():
        pytest.skip("oracle/_ref not built")
    tidx = pyref.RefIndex.load(goldenlib.path("tiny", "tiny.fa"))
    a = tidx.arrays()
    reads = cases.long_reads(a["pac"], int(a["l_pac"]), n=60) + cases.read_lines(goldenlib.path("bcr_2k.txt"))[:200]
    ids = cases.ids_for(len(reads))
    exp, _ = pyref.align(tidx, reads, pyref.default_opt(), ids)
    idx = capi.Index.load(goldenlib.path("tiny", "tiny.fa"))
    got = capi.align(idx, reads, capi.default_opt(), ids)
    assert parity.compare_results(got, exp) == []


@pytest.mark.parametrize("name", ["sim1_5k", "bcr_2k"])
def test_row_synchronous_extension_kernel_alone(name):
    """B200_WAVE_G=0 switches the wavefront extension kernel off, so every read goes through the row-synchronous kernel that
    otherwise only sees what the wavefront hands back (gap events, N bases); the choice is latched at the first call, so the run
    happens in a fresh process.  Same golden hits, CIGARs and MAPQs."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import cases, goldenlib, parity\n"
        "from seqlib_b200 import capi\n"
        "capi.set_device(0)\n"
        "gold, z = goldenlib.load(%r)\n"
        "idx = capi.Index.load(goldenlib.path('tiny', 'tiny.fa'))\n"
        "reads = cases.read_lines(goldenlib.path(%r + '.txt'))\n"
        "got = capi.align(idx, reads, capi.default_opt(), cases.ids_for(len(reads)))\n"
        "bad = parity.compare_results(got, gold)\n"
        "assert bad == [], bad[:5]\n"
        "print('lane ok', len(got.hits))\n") % (root, os.path.join(root, "tests"), name, name)
    env = dict(os.environ, B200_WAVE_G="0")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0 and "lane ok" in r.stdout, r.stdout[-1500:] + r.stderr[-1500:]


def test_cxx_dropin_kat(capi, tmp_path):
    """The reference's bwa_wrapper Boost test (seq_test/seq_test.cpp:793-915) against the C++ drop-in classes."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tests", "cxx", "test_bwa_wrapper")
    lib = os.path.join(root, "seqlib_b200", "libSeqLibB200.so")
    if not os.path.exists(exe) or os.path.getmtime(exe) < os.path.getmtime(exe + ".cpp") or (os.path.exists(lib) and os.path.getmtime(exe) < os.path.getmtime(lib)):
        subprocess.check_call(["make", "-s", "-C", os.path.join(root, "seqlib_b200", "cxx")])
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-I" + os.path.join(root, "include"), "-o", exe, exe + ".cpp",
                               "-L" + os.path.join(root, "seqlib_b200"), "-lSeqLibB200", "-lseqlib_b200",
                               "-Wl,-rpath," + os.path.join(root, "seqlib_b200")])
    r = subprocess.run([exe, str(tmp_path / "kat")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr


def test_tandem_repeat_reads_do_not_fail_the_batch(capi):
    """Reads from a 700-copy tandem repeat: every SMEM interval holds more than max_occ = 500 positions, so a read makes
    tens of thousands of seeds (bwa/bwamem.c:300-313) -- far beyond the main-pass slots.  They must go through the
    spill pass (sized from opt.max_occ), give the reference's hits, and never fail the unique reads next to them."""
    from oracle import pyref
    if not pyref.have_ref():
        pytest.skip("oracle/_ref not built")
    rng = np.random.Generator(np.random.PCG64(77))
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    unit = acgt[rng.integers(0, 4, 300)]
    left = acgt[rng.integers(0, 4, 20000)]
    right = acgt[rng.integers(0, 4, 20000)]
    ref = np.concatenate([left, np.tile(unit, 700), right])
    ref_ascii = ref.tobytes().decode()
    reads = []
    for k in range(24):                                   # inside the repeat, 2 % substitutions
        p = 20000 + int(rng.integers(0, 300 * 699))
        r = ref[p:p + 150].copy()
        for q in np.nonzero(rng.random(150) < 0.02)[0]:
            r[q] = acgt[(int(np.nonzero(acgt == r[q])[0][0]) + 1 + int(rng.integers(0, 3))) & 3]
        reads.append(r.tobytes().decode())
    for k in range(24):                                   # unique flanks and repeat boundaries
        p = int(rng.choice([rng.integers(0, 19800), 19900 + rng.integers(0, 120), 20000 + 210000 - 80 + rng.integers(0, 60), 230000 + rng.integers(100, 19800)]))
        reads.append(ref[p:p + 150].tobytes().decode())
    ids = cases.ids_for(len(reads))
    ridx = pyref.RefIndex.construct(["rep"], [ref_ascii])
    exp, _ = pyref.align(ridx, reads, pyref.default_opt(), ids)
    idx = capi.Index.construct(["rep"], [ref_ascii])
    got = capi.align(idx, reads, capi.default_opt(), ids)
    assert parity.compare_results(got, exp) == []
    st = capi.last_stats()
    assert st["n_overflow"] >= 20 and st["n_failed"] == 0


def test_reads_with_n_take_the_main_pass(capi):
    """Reads holding an N are not taken by the seeding machine; they go through the reference-shaped kernel inside the main pass
    (not the few-thread spill pass) and through the row-synchronous extension kernel.  Hits equal the reference library's."""
    from oracle import pyref
    from seqlib_b200 import synth
    if not pyref.have_ref():
        pytest.skip("oracle/_ref not built")
    l_pac = 300000
    pac = synth.reference(l_pac, seed=91)
    ctg = synth.contigs_for(l_pac, 3, "n")
    names = [c[0] for c in ctg]
    seqs = [synth.ascii_of(pac, c[1], c[1] + c[2]) for c in ctg]
    r, off, _, _ = synth.reads(pac, l_pac, ctg, 6000, 150, 0.02, 1e-3, seed=92)
    r = r.copy()
    r[7::211] = ord("N")                         # ~70 % of the reads get at least one N
    ids = cases.ids_for(6000)
    ridx = pyref.RefIndex.construct(names, seqs)
    exp, _ = pyref.align(ridx, (r, off), pyref.default_opt(), ids)
    idx = capi.Index.construct(names, seqs)
    got = capi.align(idx, (r, off), capi.default_opt(), ids)
    assert parity.compare_results(got, exp) == []
    st = capi.last_stats()
    assert st["n_overflow"] < 100 and st["ext_fallback"] > 1000 and st["n_failed"] == 0


def test_coordinates_beyond_2_32_vs_live_reference(capi):
    """A 2.2 Gb reference (forward + reverse text = 4.4 * 10^9 symbols > 2^32): interval coordinates need the high nibbles of the
    packed list entries / table entries (seed2.cuh), the suffix sorter runs its super-bucket paths, contigs straddle 2^32.  The
    reference library aligns 60 000 of the reads on the SAME index arrays (host view of the GPU-built index); everything must be
    bit-identical, and the primaries of all reads must sit at their simulated origin."""
    from oracle import pyref
    from seqlib_b200 import synth
    if not pyref.have_ref():
        pytest.skip("oracle/_ref not built")
    l_pac = 2_200_000_000
    pac = synth.reference(l_pac, seed=41)
    ctg = synth.contigs_for(l_pac, 19)
    idx = capi.Index.construct_pac(pac, l_pac, ctg, keep_host=True)
    n = 300_000
    r, off, pos, strand = synth.reads(pac, l_pac, ctg, n, 150, 0.015, 2e-4, seed=42)
    ids = cases.ids_for(n)
    opt = capi.default_opt()
    got = capi.align(idx, (r, off), opt, ids)
    rec, mapped = parity.truth_recovery(got, pos, strand, ctg, tol=12)
    assert mapped > 0.999 and rec > 0.995
    assert int(got.hits["rb"].max()) > (1 << 32)               # the reverse strand lives beyond 2^32
    m = 60_000
    ridx = pyref.RefIndex.from_view(idx.view(), keep=idx)
    exp, _ = pyref.align(ridx, (r[:m * 150], off[:m + 1]), opt, ids[:m], n_threads=os.cpu_count() or 1)
    bad, msgs = parity.compare_prefix(got, exp, m)
    assert bad == 0, msgs


@pytest.mark.parametrize("over", [dict(min_seed_len=12), dict(min_seed_len=8), dict(min_seed_len=27, split_factor=1.1), dict(split_width=3),
                                  dict(split_width=40, split_factor=1.0), dict(max_mem_intv=0), dict(max_mem_intv=120),
                                  dict(split_width=300), dict(max_mem_intv=1000, min_seed_len=15)])
def test_seeding_options_vs_live_reference(capi, over):
    """Seeding options decide the thresholds of the seeding machine and whether its chain table may be used (seed2.cuh): hits, and
    the interval lists of the debug entry point, equal the live reference's for every set."""
    from oracle import pyref
    if not pyref.have_ref():
        pytest.skip("oracle/_ref not built")
    idx = capi.Index.load(goldenlib.path("tiny", "tiny.fa"))
    tidx = pyref.RefIndex.load(goldenlib.path("tiny", "tiny.fa"))
    reads = cases.read_lines(goldenlib.path("sim1_5k.txt"))[:2500] + cases.read_lines(goldenlib.path("bcr_2k.txt"))[:1000]
    opt, ropt = capi.default_opt(), pyref.default_opt()
    for k, v in over.items():
        setattr(opt, k, v); setattr(ropt, k, v)
    ids = cases.ids_for(len(reads))
    exp, _ = pyref.align(tidx, reads, ropt, ids, n_threads=os.cpu_count() or 1)
    got = capi.align(idx, reads, opt, ids)
    assert parity.compare_results(got, exp) == []
    eoff, eintv = pyref.collect_intv(tidx, reads, ropt)
    ioff, intv = capi.collect_intv(idx, reads, opt)
    assert np.array_equal(ioff, eoff) and np.array_equal(intv, eintv)


def test_reference_with_ambiguous_bases_keeps_off_the_text_path(capi):
    """SeqLib's ConstructIndex randomises N bases twice (src/BWAIndex.cpp:102-125): the BWT is built over another text than the
    one bns_get_seq returns.  The engine's proof BWT[k] == text[SA[k]-1] fails for such an index, the seeding machine keeps to the
    Occ blocks and the chain table (no text-path requests), and the hits equal the live reference's, also around the N runs."""
    from oracle import pyref
    if not pyref.have_ref():
        pytest.skip("oracle/_ref not built")
    rng = np.random.Generator(np.random.PCG64(77))
    L = 200_000
    ref = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, L)].copy()
    for p, ln in ((20_000, 60), (90_000, 400), (150_000, 7)):
        ref[p:p + ln] = ord("N")
    seq = ref.tobytes().decode()
    _libc().srand48(3)
    idx = capi.Index.construct(["c1"], [seq])
    pyref.srand48(3)
    ridx = pyref.RefIndex.construct(["c1"], [seq])
    starts = np.concatenate([rng.integers(0, L - 150, 2500), np.array([19_900, 19_960, 89_950, 90_300, 149_900, 149_990])])
    reads = []
    for s in starts:
        r = ref[s:s + 150].copy()
        m = rng.random(150) < 0.01
        r[m] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, int(m.sum()))]
        reads.append(r.tobytes().decode())
    ids = cases.ids_for(len(reads))
    opt = capi.default_opt()
    got = capi.align(idx, reads, opt, ids)
    st = capi.last_stats()
    exp, _ = pyref.align(ridx, reads, pyref.default_opt(), ids, n_threads=os.cpu_count() or 1)
    assert parity.compare_results(got, exp) == []
    assert st["tab_lookups_lo"] == 0 and st["tab_lookups_hi"] > 0
    # the same sequence without N: the text is the BWT's text and the text path is taken
    clean = seq.replace("N", "A")
    idx2 = capi.Index.construct(["c1"], [clean])
    capi.align(idx2, reads[:500], opt, ids[:500])
    assert capi.last_stats()["tab_lookups_lo"] > 0


def test_high_copy_repeat_family_vs_live_reference(capi):
    """150 000 copies of a 300-bp element (3 % divergence) in a 60 Mb reference: 15-mers of the element occur ~95 000 times, so the
    seeding machine's ring holds entries whose interval sizes exceed 16 bits (22-bit size field for texts below 2^33 symbols).  Hits
    equal the live reference's, and the reads stay on the fast path (no spill pass)."""
    from oracle import pyref
    from seqlib_b200 import synth
    if not pyref.have_ref():
        pytest.skip("oracle/_ref not built")
    rng = np.random.Generator(np.random.PCG64(20240))
    n_copy, el, sp = 150_000, 300, 100
    elem = rng.integers(0, 4, el, dtype=np.uint8)
    ref = rng.integers(0, 4, n_copy * (el + sp), dtype=np.uint8)
    cop = np.tile(elem, n_copy).reshape(n_copy, el)
    mut = rng.random((n_copy, el)) < 0.03
    cop[mut] = (cop[mut] + rng.integers(1, 4, int(mut.sum()), dtype=np.uint8)) & 3
    ref.reshape(n_copy, el + sp)[:, :el] = cop
    l_pac = len(ref)
    pac = np.zeros((l_pac + 3) // 4 + 1, np.uint8)
    r4 = np.concatenate([ref, np.zeros((-l_pac) % 4, np.uint8)]).reshape(-1, 4)
    pac[: len(r4)] = (r4[:, 0] << 6 | r4[:, 1] << 4 | r4[:, 2] << 2 | r4[:, 3]).astype(np.uint8)
    ctg = synth.contigs_for(l_pac, 3)
    idx = capi.Index.construct_pac(pac, l_pac, ctg, keep_host=True)
    n = 1500
    starts = rng.integers(0, l_pac - 150, n)
    letters = np.frombuffer(b"ACGT", dtype=np.uint8)
    reads = []
    for s in starts:
        r = ref[s:s + 150].copy()
        m = rng.random(150) < 0.01
        r[m] = (r[m] + rng.integers(1, 4, int(m.sum()), dtype=np.uint8)) & 3
        reads.append(letters[r].tobytes().decode())
    ids = cases.ids_for(n)
    opt = capi.default_opt()
    got = capi.align(idx, reads, opt, ids)
    st = capi.last_stats()
    ridx = pyref.RefIndex.from_view(idx.view(), keep=idx)
    exp, _ = pyref.align(ridx, reads, pyref.default_opt(), ids, n_threads=os.cpu_count() or 1)
    assert parity.compare_results(got, exp) == []
    assert st["n_failed"] == 0
