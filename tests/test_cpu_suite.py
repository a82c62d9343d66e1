"""CPU-only suite: golden vectors vs the oracle(s), host logic, C ABI surface.  No CUDA compute."""
import ctypes as C
import os
import re
import numpy as np
import pytest

import cases
import goldenlib
import parity

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abi_exports_match_header():
    """Every function include/seqlib_b200.h declares is exported by libseqlib_b200.so (no compute calls)."""
    from seqlib_b200 import capi
    if not os.path.exists(capi.SO_PATH):
        pytest.skip("libseqlib_b200.so not built")
    hdr = open(os.path.join(ROOT, "include", "seqlib_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", hdr)))
    L = C.CDLL(capi.SO_PATH)
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert set(capi.EXPORTS) <= set(names)


def test_opt_defaults_match_reference():
    from seqlib_b200 import capi
    if not os.path.exists(capi.SO_PATH):
        pytest.skip("libseqlib_b200.so not built")
    o = capi.default_opt()
    assert (o.a, o.b, o.o_del, o.e_del, o.w, o.T, o.zdrop, o.min_seed_len, o.max_occ, o.mapQ_coef_fac) == (1, 4, 6, 1, 100, 30, 100, 19, 500, 3)
    assert o.flag & 0x200
    assert list(o.mat)[:6] == [1, -4, -4, -4, -1, -4]
    from oracle import pyref
    if pyref.have_ref():
        r = pyref.default_opt()
        assert bytes(r) == bytes(o)


def _sim_index_for_tiny():
    from oracle import pyref
    import simlib
    tidx = pyref.RefIndex.load(goldenlib.path("tiny", "tiny.fa"))
    return simlib.SimIndex(tidx.view(), keep=tidx), tidx


@pytest.mark.parametrize("name,small,seed_v2,tab_k,text", [("sim1_5k", False, 0, 0, 0), ("bcr_2k", False, 0, 0, 0), ("bcr_2k", True, 0, 0, 0),
                                                           ("sim1_5k", False, 32, 0, 0), ("bcr_2k", False, 16, 0, 0), ("bcr_2k", False, 8, 0, 0),
                                                           ("sim1_5k", False, 16, 8, 0), ("bcr_2k", False, 16, 8, 0), ("bcr_2k", False, 16, 4, 0),
                                                           ("sim1_5k", False, 16, 5, 0), ("bcr_2k", False, 4, 6, 0),
                                                           ("sim1_5k", False, 16, 8, 1), ("bcr_2k", False, 16, 9, 1), ("sim1_5k", False, 32, 0, 1),
                                                           ("bcr_2k", False, 2, 6, 1), ("sim1_5k", False, 8, 7, 1)])
def test_stage_functions_on_cpu_vs_golden(name, small, seed_v2, tab_k, text, monkeypatch):
    """The per-read device functions, compiled for the host (tests/hostsim), reproduce the golden vectors;
    `small` forces every read through the spill path; seed_v2 = work-list capacity of the single-extension-site
    seeding machine (seed2.cuh) the GPU kernel runs (0: the reference-shaped loops; small ones: reads overflow the list
    and fall back, as in the kernel's spill pass); tab_k = depth of the prefix-chain table that answers everything
    bwt_smem1a asks about strings of at most tab_k bases (interval lists must stay identical); text = the machine
    follows size-one intervals through the text (same hits; such an interval carries its text position instead of x0)."""
    if seed_v2:
        monkeypatch.setenv("HOSTSIM_SEED_V2", str(seed_v2))
    else:
        monkeypatch.delenv("HOSTSIM_SEED_V2", raising=False)
    monkeypatch.setenv("HOSTSIM_SEED_TAB_K", str(tab_k))
    monkeypatch.setenv("HOSTSIM_SEED_TEXT", str(text))
    from oracle import pyref
    if not pyref.have_ref():
        pytest.skip("needs oracle/_ref to parse the bwa index for the host harness")
    import simlib
    gold, z = goldenlib.load(name)
    sidx, tidx = _sim_index_for_tiny()
    reads = cases.read_lines(goldenlib.path(name + ".txt"))
    if small:
        reads = reads[:600]
    got = simlib.align(sidx, reads, pyref.default_opt(), cases.ids_for(len(reads)), small)
    if small:
        n = len(reads)
        assert np.array_equal(got.hit_off, gold.hit_off[:n + 1])
        k = int(gold.hit_off[n])
        for f in parity.REG_FIELDS + parity.ALN_FIELDS:
            assert np.array_equal(got.hits[f], gold.hits[f][:k]), f
    else:
        assert parity.compare_results(got, gold) == []
        assert np.array_equal(got.intv_off, z["intv_off"])
        gi, zi = got.intv, z["intv"]
        assert np.array_equal(gi["x2"], zi["x2"]) and np.array_equal(gi["info"], zi["info"])
        flagged = (gi["x0"] >> np.uint64(63)) != 0
        assert np.array_equal(gi["x0"][~flagged], zi["x0"][~flagged])
        if not text:
            assert not flagged.any() and np.array_equal(gi["x1"], zi["x1"])
        else:
            # a flagged interval: size one, and the text at its position spells the read's substring
            assert (flagged.mean() > 0.1 or seed_v2 < 8) and (gi["x2"][flagged] == 1).all()
            a = tidx.arrays()
            l_pac = int(a["l_pac"])
            fwd = np.unpackbits(np.frombuffer(a["pac"], dtype=np.uint8)[: (l_pac + 3) // 4]).reshape(-1, 2)
            fwd = (fwd[:, 0] * 2 + fwd[:, 1])[:l_pac].astype(np.uint8)
            txt = np.concatenate([fwd, 3 - fwd[::-1]])
            code = np.full(256, 4, np.uint8)
            for i, ch in enumerate("ACGT"):
                code[ord(ch)] = i; code[ord(ch.lower())] = i
            rid = np.searchsorted(z["intv_off"], np.arange(len(gi)), side="right") - 1
            idxs = np.nonzero(flagged)[0]
            for t in idxs[:: max(1, len(idxs) // 3000)]:
                pos = int(gi["x0"][t] & np.uint64((1 << 63) - 1))
                st, en = int(gi["info"][t] >> np.uint64(32)), int(gi["info"][t] & np.uint64(0xffffffff))
                q = code[np.frombuffer(reads[rid[t]].encode(), dtype=np.uint8)][st:en]
                assert np.array_equal(txt[pos:pos + en - st], q), (t, pos, st, en)


SEED_OPTION_SETS = [dict(min_seed_len=12), dict(min_seed_len=8), dict(min_seed_len=27, split_factor=1.1), dict(split_width=3),
                    dict(split_width=40, split_factor=1.0), dict(max_mem_intv=0), dict(max_mem_intv=6), dict(max_mem_intv=120),
                    dict(split_width=300), dict(max_mem_intv=1000, min_seed_len=15)]


@pytest.mark.parametrize("over", SEED_OPTION_SETS)
def test_seeding_machine_options_on_cpu_vs_reference(over, monkeypatch):
    """The seeding options reach the machine's thresholds (min_intv of the re-seeding pass = split_width + 1 at most, max_mem_intv of
    the third pass, the report filter min_seed_len) and decide whether the chain table may be used at all (min_seed_len > K,
    thresholds <= 255): interval lists and hits of the host build (table K = 8, text path) equal the live reference's."""
    from oracle import pyref
    if not pyref.have_ref():
        pytest.skip("oracle/_ref not built")
    import simlib
    monkeypatch.setenv("HOSTSIM_SEED_V2", "16")
    monkeypatch.setenv("HOSTSIM_SEED_TAB_K", "8")
    monkeypatch.setenv("HOSTSIM_SEED_TEXT", "1")
    sidx, tidx = _sim_index_for_tiny()
    reads = cases.read_lines(goldenlib.path("sim1_5k.txt"))[:1200] + cases.read_lines(goldenlib.path("bcr_2k.txt"))[:600]
    opt = pyref.default_opt()
    for k, v in over.items():
        setattr(opt, k, v)
    ids = cases.ids_for(len(reads))
    exp, _ = pyref.align(tidx, reads, opt, ids)
    eoff, eintv = pyref.collect_intv(tidx, reads, opt)
    got = simlib.align(sidx, reads, opt, ids, False)
    assert parity.compare_results(got, exp) == []
    assert np.array_equal(got.intv_off, eoff)
    assert np.array_equal(got.intv["x2"], eintv["x2"]) and np.array_equal(got.intv["info"], eintv["info"])
    flagged = (got.intv["x0"] >> np.uint64(63)) != 0
    assert np.array_equal(got.intv["x0"][~flagged], eintv["x0"][~flagged])


def test_long_reads_seed_sw_filter_on_cpu_vs_reference(monkeypatch):
    """Contig-like queries (0.8-4 kb) activate mem_flt_chained_seeds / mem_seed_sw -> ksw_i16 (bwa/bwamem.c:597-641): the host
    build of the stage functions (seedsw.cuh replays the striped kernel) gives the reference's hits, CIGARs and MAPQs; with
    the filter switched off the same inputs differ, so the inputs do exercise it."""
    from oracle import pyref
    if not pyref.have_ref():
        pytest.skip("oracle/_ref not built")
    import simlib
    sidx, tidx = _sim_index_for_tiny()
    a = tidx.arrays()
    reads = cases.long_reads(a["pac"], int(a["l_pac"]), n=24)
    opt = pyref.default_opt()
    ids = cases.ids_for(len(reads))
    exp, _ = pyref.align(tidx, reads, opt, ids)
    monkeypatch.delenv("HOSTSIM_SEED_V2", raising=False)
    got = simlib.align(sidx, reads, opt, ids, False)
    assert parity.compare_results(got, exp) == []
    monkeypatch.setenv("HOSTSIM_NO_SEEDSW", "1")
    off = simlib.align(sidx, reads, opt, ids, False)
    assert parity.compare_results(off, exp) != []


def test_reference_library_reproduces_golden():
    """oracle/_ref (when present) still produces the committed vectors: guards the fixtures themselves."""
    from oracle import pyref
    if not pyref.have_ref():
        pytest.skip("oracle/_ref not built")
    gold, z = goldenlib.load("kat")
    pyref.srand48(cases.KAT_SRAND)
    kidx = pyref.RefIndex.construct(cases.KAT_NAMES, cases.KAT_SEQS)
    a = kidx.arrays()
    assert np.array_equal(a["bwt"], z["bwt"]) and np.array_equal(a["sa"], z["sa"])
    res, _ = pyref.align(kidx, cases.KAT_QUERIES, pyref.default_opt(), z["ids"])
    assert parity.compare_results(res, gold) == []
    # survey goldens: seq_len 646, primary 28 on the default lrand48 stream is process-state dependent;
    # structure that is not: two 38M hits of score 38 for query 0, rid {2,0}
    h = res.read_hits(0)
    assert sorted(h["rid"].tolist()) == [0, 2] and all(res.cigar_str(x) == "38M" for x in h) and set(h["score"]) == {38}


def test_synth_is_deterministic():
    from seqlib_b200 import synth
    p1 = synth.reference(10000)
    p2 = synth.reference(10000)
    assert np.array_equal(p1, p2)
    ctg = synth.contigs_for(10000, 1, "ref10k")
    a = synth.reads(p1, 10000, ctg, 100)
    b = synth.reads(p1, 10000, ctg, 100)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    assert synth.ascii_of(p1, 0, 20) == "AGTAAATGTTCCTCAGACTG"


def _opt_default():
    """mem_opt_init + MEM_F_SOFTCLIP without needing any shared library."""
    from seqlib_b200.abi import MemOpt
    o = MemOpt()
    o.a, o.b, o.o_del, o.e_del, o.o_ins, o.e_ins = 1, 4, 6, 1, 6, 1
    o.pen_unpaired, o.pen_clip5, o.pen_clip3, o.w, o.zdrop = 17, 5, 5, 100, 100
    o.max_mem_intv, o.T, o.flag, o.min_seed_len, o.min_chain_weight, o.max_chain_extend = 20, 30, 0x200, 19, 0, 1 << 30
    o.split_factor, o.split_width, o.max_occ, o.max_chain_gap, o.n_threads, o.chunk_size = 1.5, 10, 500, 10000, 1, 10000000
    o.mask_level, o.drop_ratio, o.XA_drop_ratio, o.mask_level_redun = 0.5, 0.5, 0.8, 0.95
    o.mapQ_coef_len, o.mapQ_coef_fac, o.max_ins, o.max_matesw, o.max_XA_hits, o.max_XA_hits_alt = 50, 3, 10000, 50, 5, 200
    m = []
    for i in range(4):
        m += [1 if i == j else -4 for j in range(4)] + [-1]
    m += [-1] * 5
    for i, x in enumerate(m):
        o.mat[i] = x
    return o


@pytest.mark.parametrize("name", ["sim1_5k", "bcr_2k"])
def test_c_restatement_vs_golden_tiny(name):
    """oracle/oracle_bwa.c (plain-C restatement) reproduces the golden vectors made by the reference's own bwa."""
    from oracle import pyoracle
    if not pyoracle.have():
        pytest.skip("oracle/liboracle.so not built")
    gold, z = goldenlib.load(name)
    view = pyoracle.load_bwa_index(goldenlib.path("tiny", "tiny.fa"))
    reads = cases.read_lines(goldenlib.path(name + ".txt"))
    got = pyoracle.align(view, reads, _opt_default(), cases.ids_for(len(reads)))
    assert parity.compare_results(got, gold) == []


def test_c_restatement_vs_golden_kat_and_c1():
    from oracle import pyoracle
    if not pyoracle.have():
        pytest.skip("oracle/liboracle.so not built")
    gold, z = goldenlib.load("kat")
    v = pyoracle.View(z["primary"], z["L2"], z["bwt"], z["sa"], 32, 323, z["pac"], [(n, sum(len(s) for s in cases.KAT_SEQS[:i]), len(cases.KAT_SEQS[i]))
                                                                                       for i, n in enumerate(cases.KAT_NAMES)])
    got = pyoracle.align(v, cases.KAT_QUERIES, _opt_default(), z["ids"])
    assert parity.compare_results(got, gold) == []
    gold, z = goldenlib.load("c1")
    pac, ctg, _ = cases.c1_reference()
    v = pyoracle.View(z["primary"], z["L2"], z["bwt"], z["sa"], 32, 10000, pac, ctg)
    seqs, off = cases.c1_reads(pac, ctg)
    got = pyoracle.align(v, (seqs, off), _opt_default(), cases.ids_for(len(off) - 1))
    assert parity.compare_results(got, gold) == []


def test_c_restatement_ksw_vs_golden():
    from oracle import pyoracle
    if not pyoracle.have():
        pytest.skip("oracle/liboracle.so not built")
    z = np.load(goldenlib.path("ksw_c3.npz"))
    jobs, qp, tp = cases.c3_tuples(4000)
    out = pyoracle.ksw_extend2_batch(jobs, qp, tp, np.array(list(_opt_default().mat), dtype=np.int8))
    for f in out.dtype.names:
        assert np.array_equal(out[f], z["out"][f]), f


def test_c_restatement_vs_live_reference_random():
    """Random reference + reads with indels and Ns: restatement == the reference's compiled bwa."""
    from oracle import pyoracle, pyref
    if not (pyoracle.have() and pyref.have_ref()):
        pytest.skip("needs liboracle.so and oracle/_ref")
    from seqlib_b200 import synth
    l_pac = 50000
    pac = synth.reference(l_pac, seed=123)
    ctg = synth.contigs_for(l_pac, 2, "k")
    names = [c[0] for c in ctg]
    seqs = [synth.ascii_of(pac, c[1], c[1] + c[2]) for c in ctg]
    ridx = pyref.RefIndex.construct(names, seqs)
    a = ridx.arrays()
    v = pyoracle.View(a["primary"], a["L2"], a["bwt"], a["sa"], 32, l_pac, a["pac"], ctg)
    r, off, _, _ = synth.reads(pac, l_pac, ctg, 3000, 150, 0.03, 3e-3, seed=321)
    r = r.copy()
    r[::613] = ord("N")
    ids = cases.ids_for(3000)
    opt = pyref.default_opt()
    exp, _ = pyref.align(ridx, (r, off), opt, ids)
    got = pyoracle.align(v, (r, off), opt, ids)
    assert parity.compare_results(got, exp) == []


def _emul_lib():
    """tests/hostsim/{group,reg}_emul.cpp: lock-step CPU emulations of the lane-cooperative DP kernels."""
    import ctypes as C
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    so = os.path.join(here, "hostsim", "libemul.so")
    srcs = [os.path.join(here, "hostsim", f) for f in ("group_emul.cpp", "reg_emul.cpp", "wave_emul.cpp")]
    csrc = os.path.join(here, "..", "seqlib_b200", "csrc")
    deps = srcs + [os.path.join(csrc, f) for f in ("ksw.cuh", "ksw_reg.cuh", "ksw_wave.cuh", "common.cuh", "fmindex.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(so) < os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-o", so] + srcs)
    return C.CDLL(so)


@pytest.mark.parametrize("fn,G", [("reg_emul_check", 8), ("reg_emul_check", 4), ("group_emul_check", 8)])
def test_lane_cooperative_extend_equals_scalar(fn, G):
    """The lane-cooperative ksw_extend2 schemes (registers: ksw_reg.cuh, shared memory: ksw_group.cuh), emulated in
    lock step on the CPU with the kernels' own phase functions, give the scalar recurrence's six outputs and cell
    count on config-3 shaped pairs and on mixed shapes (short queries, unrelated targets, narrow bands, z-drop)."""
    import ctypes as C
    L = _emul_lib()
    mat = np.array(list(_opt_default().mat), dtype=np.int8)
    sets = []
    j, q, t = cases.c3_tuples_fast(6000)
    sets.append((j, q, t, 100))
    j, q, t = cases.mixed_tuples(30000)
    q = np.minimum(q, 3).astype(np.uint8)        # the register scheme is used on the 2-bit reference text and N-free reads only
    for zd in (100, 0, 20):
        sel = j[j["zdrop"] == zd]
        sets.append((sel, q, t, zd))
    for jobs, qp, tp, zd in sets:
        n = len(jobs)
        ql = np.ascontiguousarray(jobs["qlen"], dtype=np.int32); tl = np.ascontiguousarray(jobs["tlen"], dtype=np.int32)
        qo = np.ascontiguousarray(jobs["q_off"], dtype=np.int64); to = np.ascontiguousarray(jobs["t_off"], dtype=np.int64)
        ws = np.ascontiguousarray(jobs["w"], dtype=np.int32); h0 = np.ascontiguousarray(jobs["h0"], dtype=np.int32)
        first = C.c_long(-1)
        bad = getattr(L, fn)(C.c_int(G), C.c_long(n), ql.ctypes.data_as(C.c_void_p), tl.ctypes.data_as(C.c_void_p), qo.ctypes.data_as(C.c_void_p),
                             to.ctypes.data_as(C.c_void_p), qp.ctypes.data_as(C.c_void_p), tp.ctypes.data_as(C.c_void_p), ws.ctypes.data_as(C.c_void_p),
                             h0.ctypes.data_as(C.c_void_p), mat.ctypes.data_as(C.c_void_p), C.c_int(6), C.c_int(1), C.c_int(6), C.c_int(1), C.c_int(5),
                             C.c_int(zd), C.byref(first))
        assert bad == 0, (fn, G, zd, first.value)


@pytest.mark.parametrize("G", [4, 8, 2])
def test_wavefront_extend_equals_scalar(G):
    """ksw_wave.cuh (packed 16-bit anti-diagonal wavefront: two target rows per lane, the column state streamed lane to
    lane) emulated in lock step on the CPU with the kernel's own lane-step and commit code: every job it accepts gives the
    scalar recurrence's six outputs; the jobs it hands back to the row-synchronous kernel (gap events) stay rare."""
    import ctypes as C
    L = _emul_lib()
    mat = np.array(list(_opt_default().mat), dtype=np.int8)
    sets = []
    j, q, t = cases.c3_tuples_fast(4000)
    sets.append((j, q, t, 100))
    j, q, t = cases.mixed_tuples(30000)
    q = np.minimum(q, 3).astype(np.uint8)
    t = np.minimum(t, 3).astype(np.uint8)
    for zd in (100, 0, 20):
        sets.append((j[j["zdrop"] == zd], q, t, zd))
    tot_run = tot_gap = 0
    for jobs, qp, tp, zd in sets:
        n = len(jobs)
        ql = np.ascontiguousarray(jobs["qlen"], dtype=np.int32); tl = np.ascontiguousarray(jobs["tlen"], dtype=np.int32)
        qo = np.ascontiguousarray(jobs["q_off"], dtype=np.int64); to = np.ascontiguousarray(jobs["t_off"], dtype=np.int64)
        ws = np.ascontiguousarray(jobs["w"], dtype=np.int32); h0 = np.ascontiguousarray(jobs["h0"], dtype=np.int32)
        first = C.c_long(-1); ngap = C.c_long(0); nrun = C.c_long(0); ratio = C.c_double(0)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        bad = L.wave_emul_check(C.c_int(G), C.c_long(n), p(ql), p(tl), p(qo), p(to), p(qp), p(tp), p(ws), p(h0), p(mat),
                                6, 1, 6, 1, 5, int(zd), C.byref(first), C.byref(ngap), C.byref(nrun), C.byref(ratio))
        assert bad == 0, "first wrong job %d (zdrop %d)" % (first.value, zd)
        assert 1.0 <= ratio.value < 1.25          # cells streamed vs the reference's band-trimmed cells
        tot_run += nrun.value; tot_gap += ngap.value
    assert tot_run > 30000 and tot_gap <= tot_run // 500


def test_compare_prefix_detects_changes():
    """parity.compare_prefix (the vectorised checker bench.py runs inside the headline measurement) agrees with the field-by-field one."""
    import copy
    gold, z = goldenlib.load("bcr_2k")
    n = len(gold.hit_off) - 1
    assert parity.compare_prefix(gold, gold, n) == (0, [])
    g2 = copy.deepcopy(gold)
    g2.cigar = g2.cigar.copy(); g2.cigar[5] ^= 16
    bad, msgs = parity.compare_prefix(g2, gold, n)
    assert bad == 1 and "cigar" in msgs[0]
    g3 = copy.deepcopy(gold)
    g3.hits = g3.hits.copy(); g3.hits["mapq"][7] += 1
    bad, msgs = parity.compare_prefix(g3, gold, n)
    assert bad == 1 and "mapq" in msgs[0]


def test_aligner_half_fails_loudly_without_a_device():
    """No CPU fallback in the product: without a CUDA device the aligner entry points (index construction, batch alignment, the
    ksw batch, page-locked host memory) return B200_ERR_CUDA / B200_ERR_NOMEM instead of computing anything on the host."""
    import ctypes as C
    from seqlib_b200 import capi
    if not os.path.exists(capi.SO_PATH):
        pytest.skip("libseqlib_b200.so not built")
    L = capi.lib()
    if L.b200_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(capi.B200Error):
        capi.Index.construct(["c"], ["ACGTACGTTTGACCAGT" * 20])
    L.b200_host_alloc.argtypes = [C.c_size_t, C.POINTER(C.c_void_p)]
    L.b200_host_alloc.restype = C.c_int
    p = C.c_void_p()
    assert L.b200_host_alloc(1 << 20, C.byref(p)) != 0 and not p.value
    L.b200_host_free.argtypes = [C.c_void_p]
    L.b200_host_free(None)                      # harmless
