#!/usr/bin/env python
"""Ingest throughput: strict 150-bp FASTQ text -> flat (bytes, offsets) batches.
  reference : the reference's kseq parser (oracle/_ref/libseqref_kseq.so, one thread, the only way it runs)
  stream    : b200_fastq_next_batch on the same file / on the in-memory text
  device    : b200_fastq_parse_device on the in-memory text (H2D of the text and D2H of the fields inside the call)
Prints one JSON object."""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def make_text(n, read_len=150, seed=5):
    rng = np.random.default_rng(seed)
    rec = 1 + 9 + 1 + read_len + 3 + read_len + 1
    a = np.empty((n, rec), dtype=np.uint8)
    a[:, 0] = ord("@"); a[:, 1] = ord("r")
    idx = np.arange(n)
    for d in range(8):
        a[:, 2 + d] = 48 + (idx // 10 ** (7 - d)) % 10
    a[:, 10] = 10
    a[:, 11:11 + read_len] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=(n, read_len))]
    a[:, 11 + read_len] = 10; a[:, 12 + read_len] = ord("+"); a[:, 13 + read_len] = 10
    a[:, 14 + read_len:14 + 2 * read_len] = rng.integers(33, 74, size=(n, read_len), dtype=np.uint8)
    a[:, 14 + 2 * read_len] = 10
    return a.tobytes()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--records", type=int, default=2000000)
    ap.add_argument("--no-device", action="store_true")
    args = ap.parse_args()
    from seqlib_b200 import fastq
    from oracle import pyref_kseq
    text = make_text(args.records)
    mb = len(text) / 1e6
    out = {"records": args.records, "text_mb": mb}
    d = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
    path = os.path.join(d, "b200_bench_%d.fq" % os.getpid())
    open(path, "wb").write(text)
    try:
        if pyref_kseq.have_ref():
            t = time.perf_counter(); n = pyref_kseq.lib()  # load
            import ctypes as C
            fields = (C.c_void_p * 4)(); offs = (C.c_void_p * 4)(); has = C.c_void_p(); last = C.c_int()
            t = time.perf_counter()
            n = n.refdrv_kseq_parse(path.encode(), 1 << 62, fields, offs, C.byref(has), C.byref(last))
            dt = time.perf_counter() - t
            out["reference_kseq"] = {"mb_s": mb / dt, "reads_s": n / dt, "seconds": dt, "cores": 1}
        r = fastq.FastqReader(path=path)
        t = time.perf_counter(); n = 0
        import ctypes as C
        from seqlib_b200.fastq import FastqBatch, _bind
        L = _bind(); b = FastqBatch()
        while True:
            L.b200_fastq_next_batch(r.h, 1 << 20, C.byref(b)); n += b.n
            if b.status != 0:
                break
        dt = time.perf_counter() - t
        out["stream_file"] = {"mb_s": mb / dt, "reads_s": n / dt, "seconds": dt, "cores": 1}
        r.close()
        r = fastq.FastqReader(text=text)
        t = time.perf_counter()
        L.b200_fastq_next_batch(r.h, 1 << 40, C.byref(b))
        dt = time.perf_counter() - t
        out["stream_mem"] = {"mb_s": mb / dt, "reads_s": b.n / dt, "seconds": dt, "cores": 1}
        r.close()
        if not args.no_device:
            r = fastq.FastqReader(text=b"")
            a = np.frombuffer(text, dtype=np.uint8)
            best = None
            for it in range(4):
                t = time.perf_counter()
                rc = L.b200_fastq_parse_device(r.h, a.ctypes.data_as(C.c_void_p), len(a), C.byref(b))
                dt = time.perf_counter() - t
                assert rc == 0 and b.n == args.records
                if it and (best is None or dt < best):
                    best = dt
            out["device_e2e"] = {"mb_s": mb / best, "reads_s": args.records / best, "seconds": best,
                                 "note": "pageable host text in, host fields out; H2D + 6 kernels/cub passes + D2H inside"}
            r.close()
    finally:
        os.unlink(path)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
