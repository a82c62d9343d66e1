"""GPU line parser for strict four-line FASTQ (b200_fastq_parse_device) against the reference parser's golden outputs and
the stream parser, and its refusal of everything that is not strict four-line FASTQ."""
import json

import numpy as np
import pytest

import fastqcases
import goldenlib

pytestmark = pytest.mark.gpu

STRICT = {"strict", "strict_ragged", "strict_crlf", "strict_no_final_newline", "qual_starts_with_at", "empty_name",
          "single_char_cr", "cr_only_comment", "vt_ff_delims", "empty"}


@pytest.mark.parametrize("name", sorted(fastqcases.CASES))
def test_device_parser_vs_golden(name):
    from seqlib_b200 import fastq
    from seqlib_b200.capi import B200Error
    gold = json.load(open(goldenlib.path("fastq_cases.json")))[name]
    r = fastq.FastqReader(text=b"")
    text = fastqcases.CASES[name]
    if name in STRICT:
        b = r.parse_device(text)
        assert b.parsed_on_device == 1 and b.status == 1
        assert [[f.decode("latin1") for f in rec] for rec in b.records()] == gold["records"]
        assert [int(x) for x in b.has] == gold["has"]
    else:
        with pytest.raises(B200Error):
            r.parse_device(text)
    r.close()


def test_device_parser_large_equals_stream_parser_and_feeds_the_aligner():
    """200k ragged records: the device parser and the stream parser give the same flat buffers, and the (seq, seq_off) pair goes
    into b200_mem_align_batch as it is."""
    from seqlib_b200 import fastq
    text = fastqcases.strict_fastq(200000, seed=11, read_len=0)
    r = fastq.FastqReader(text=text)
    a = r.next_batch(1 << 40)
    d = fastq.FastqReader(text=b"")
    b = d.parse_device(text)
    assert a.n == b.n == 200000
    for f in ("seq", "qual", "name", "comment"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f
        assert np.array_equal(getattr(a, f + "_off"), getattr(b, f + "_off")), f
    assert np.array_equal(a.has, b.has)
    r.close(); d.close()
