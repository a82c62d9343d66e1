"""FASTA/FASTQ ingest: b200_fastq_open / next_batch against the reference's kseq parser (oracle/_ref,
compiled from bwa/kseq.h in place) and against the committed golden outputs."""
import gzip
import json
import os

import pytest

import fastqcases
import goldenlib


def _ours(path, batch):
    from seqlib_b200 import fastq
    r = fastq.FastqReader(path=path)
    recs, last, seen = [], 0, []
    while True:
        b = r.next_batch(batch)
        recs += b.records()
        if b.status != 0:
            last = -1 if b.status == 1 else b.status
            break
        assert b.n == batch
    seen = r.buffers_seen()
    r.close()
    return recs, last, seen


def _ours_has(path, batch):
    from seqlib_b200 import fastq
    r = fastq.FastqReader(path=path)
    has = []
    while True:
        b = r.next_batch(batch)
        has += [int(x) for x in b.has]
        if b.status != 0:
            break
    r.close()
    return has


def _write(tmp_path, name, text, gz):
    p = os.path.join(str(tmp_path), name + (".fq.gz" if gz else ".fq"))
    if gz:
        with gzip.open(p, "wb") as f:
            f.write(text)
    else:
        with open(p, "wb") as f:
            f.write(text)
    return p


@pytest.mark.parametrize("name", sorted(fastqcases.CASES))
def test_stream_parser_vs_golden(name, tmp_path):
    """every case against the committed outputs of the reference's parser (tests/golden/fastq_cases.json)"""
    gold = json.load(open(goldenlib.path("fastq_cases.json")))[name]
    for gz in (False, True):
        for batch in (1, 7, 1 << 20):
            recs, last, seen = _ours(_write(tmp_path, name, fastqcases.CASES[name], gz), batch)
            assert [[f.decode("latin1") for f in r] for r in recs] == gold["records"], (name, gz, batch)
            assert last == gold["last"]
            assert seen == (gold["has"][-1] if gold["has"] and last == -1 else seen)
            assert _ours_has(_write(tmp_path, name, fastqcases.CASES[name], gz), batch) == gold["has"]


def test_stream_parser_vs_live_reference(tmp_path):
    from oracle import pyref_kseq
    if not pyref_kseq.have_ref():
        pytest.skip("oracle/_ref/libseqref_kseq.so not built")
    for name, text in fastqcases.CASES.items():
        p = _write(tmp_path, name, text, False)
        exp, has, last = pyref_kseq.parse(p)
        recs, mylast, seen = _ours(p, 13)
        assert recs == exp, name
        assert mylast == last, name
    # a larger ragged file through a 1 MiB buffer boundary, gz
    big = fastqcases.strict_fastq(40000, seed=3, read_len=0)
    p = _write(tmp_path, "big", big, True)
    exp, has, last = pyref_kseq.parse(p)
    recs, mylast, seen = _ours(p, 4096)
    assert recs == exp and mylast == last and len(recs) == 40000


def test_reference_fixture_files():
    """the reference's own FASTQ fixtures (tests/data/*.fq) when the mount is present"""
    from oracle import pyref_kseq
    d = "/root/reference/tests/data"
    if not (pyref_kseq.have_ref() and os.path.isdir(d)):
        pytest.skip("reference mount absent")
    n = 0
    for f in sorted(os.listdir(d)):
        if f.endswith((".fq", ".fa", ".fq.gz", ".fa.gz", ".fasta")):
            p = os.path.join(d, f)
            exp, has, last = pyref_kseq.parse(p)
            recs, mylast, seen = _ours(p, 1000)
            assert recs == exp and mylast == last, f
            n += 1
    assert n > 0


def test_open_errors():
    from seqlib_b200 import fastq
    from seqlib_b200.capi import B200Error
    with pytest.raises(B200Error):
        fastq.FastqReader(path="/nonexistent/file.fq")


def test_device_parser_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from seqlib_b200 import fastq
    from seqlib_b200.capi import B200Error
    r = fastq.FastqReader(text=b"")
    with pytest.raises(B200Error):
        r.parse_device(fastqcases.CASES["strict"])


def test_cxx_fastq_reader_class(tmp_path):
    """SeqLib::FastqReader (include/SeqLib/FastqReader.h) used like the reference's class: Open + GetNextSequence loop on a
    reused UnalignedSequence; Com / Qual keep the caller's value until the parser has seen a comment / a '+' line."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tests", "cxx", "test_fastq_reader")
    subprocess.check_call(["make", "-s", "-C", os.path.join(root, "seqlib_b200", "cxx")])
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I" + os.path.join(root, "include"), os.path.join(root, "tests", "cxx", "test_fastq_reader.cpp"),
                           "-o", exe, "-L" + os.path.join(root, "seqlib_b200"), "-lSeqLibB200", "-lseqlib_b200",
                           "-Wl,-rpath," + os.path.join(root, "seqlib_b200")])
    gold = json.load(open(goldenlib.path("fastq_cases.json")))
    for name, text in fastqcases.CASES.items():
        if b"\x01" in text:
            continue
        p = _write(tmp_path, name, text, name.startswith("strict"))
        out = subprocess.run([exe, p], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout
        lines = out.split(b"\n")[:-1] if out else []
        g = gold[name]
        exp = []
        for rec, has in zip(g["records"], g["has"]):
            nm, com, sq, ql = [f.encode("latin1") for f in rec]
            exp.append(b"\x01".join([nm, com if has & 1 else b"KEEP", sq, ql if has & 2 else b"KEEPQ"]))
        if name in ("cr_only_comment",):          # a bare CR inside a field would break the line-based comparison
            continue
        assert lines == exp, name


def test_truncated_gzip_is_an_error_not_end_of_input(tmp_path):
    """A gzip stream cut in the middle: gzread fails; the batch call must report B200_ERR_IO instead of a clean end of input
    with silently fewer records (kseq itself stops quietly there; the batch API has an error channel, so it uses it)."""
    import gzip
    import random
    from seqlib_b200 import fastq as fq, capi
    rnd = random.Random(3)
    text = "".join("@r%d\n%s\n+\n%s\n" % (i, "".join(rnd.choice("ACGT") for _ in range(100)), "I" * 100) for i in range(3000)).encode()
    p = tmp_path / "cut.fq.gz"
    blob = gzip.compress(text)
    p.write_bytes(blob[:len(blob) // 2])
    r = fq.FastqReader(path=str(p))
    with pytest.raises(capi.B200Error) as e:
        while True:
            b = r.next_batch(1000)
            if b.status != 0:
                break
    assert "b200 error -4" in str(e.value)
    r.close()
