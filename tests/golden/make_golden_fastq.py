"""Generates tests/golden/fastq_cases.json: the output of the reference's kseq parser (oracle/_ref/libseqref_kseq.so, built
from /root/reference/bwa/kseq.h by oracle/Makefile) on every text of tests/fastqcases.py.  Run from the repo root:
    python tests/golden/make_golden_fastq.py"""
import json
import os
import sys
import tempfile

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import fastqcases
from oracle import pyref_kseq

out = {}
with tempfile.TemporaryDirectory() as d:
    for name, text in fastqcases.CASES.items():
        p = os.path.join(d, name + ".fq")
        open(p, "wb").write(text)
        recs, has, last = pyref_kseq.parse(p)
        out[name] = {"records": [[f.decode("latin1") for f in r] for r in recs], "has": has, "last": last}
json.dump(out, open(os.path.join(os.path.dirname(__file__), "fastq_cases.json"), "w"), indent=0)
print({k: (len(v["records"]), v["last"]) for k, v in out.items()})
