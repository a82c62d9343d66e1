#!/usr/bin/env python
"""The whole chain a `bwa mem reads.fq` user sees, on the B200 path: FASTQ text -> records (b200_fastq_parse_device) -> alignments
(b200_mem_align_batch) -> SAM text (b200_results_to_sam), N synthetic 150-bp reads against the 3 Gb random reference; next to it the
reference's mem_process_seqs (which ends in mem_reg2sam) on all host cores for a bounded sample of the same reads.
Prints one JSON object; --out writes the SAM text."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def fastq_text(seqs, n, read_len):
    rec = 1 + 9 + 1 + read_len + 3 + read_len + 1
    a = np.empty((n, rec), dtype=np.uint8)
    a[:, 0] = ord("@"); a[:, 1] = ord("r")
    idx = np.arange(n)
    for d in range(8):
        a[:, 2 + d] = 48 + (idx // 10 ** (7 - d)) % 10
    a[:, 10] = 10
    a[:, 11:11 + read_len] = np.asarray(seqs[:n * read_len]).reshape(n, read_len)
    a[:, 11 + read_len] = 10; a[:, 12 + read_len] = ord("+"); a[:, 13 + read_len] = 10
    a[:, 14 + read_len:14 + 2 * read_len] = ord("I")
    a[:, 14 + 2 * read_len] = 10
    return a.tobytes()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=4000000)
    ap.add_argument("--ref-len", type=int, default=3000000000)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--out", default=None)
    ap.add_argument("--chunk", type=int, default=1000000)
    args = ap.parse_args()
    import ctypes as C
    from seqlib_b200 import capi, synth, fastq, sam
    from seqlib_b200.fastq import FastqBatch, _bind
    capi.set_device(0)
    l_pac = args.ref_len
    pac = synth.reference(l_pac)
    ctg = synth.contigs_for(l_pac, 24 if l_pac >= 24000 else 1)
    seqs, off, _, _ = synth.reads(pac, l_pac, ctg, args.reads, 150, 0.01, 0.0)
    text = fastq_text(seqs, args.reads, 150)
    t0 = time.perf_counter()
    idx = capi.Index.construct_pac(pac, l_pac, ctg, keep_host=True)
    t_index = time.perf_counter() - t0
    rnames = [idx.seq_name(i) for i in range(idx.n_seqs())]
    opt = capi.default_opt()
    ids = np.arange(args.reads, dtype=np.int64)
    out = {"reads": args.reads, "fastq_mb": len(text) / 1e6, "index_build_s": t_index}
    rd = fastq.FastqReader(text=b"")
    L = _bind()
    a = np.frombuffer(text, dtype=np.uint8)
    best = None
    for it in range(2):             # first pass warms the pools
        t0 = time.perf_counter()
        b = FastqBatch()
        assert L.b200_fastq_parse_device(rd.h, a.ctypes.data_as(C.c_void_p), len(a), C.byref(b)) == 0 and b.n == args.reads
        t1 = time.perf_counter()
        h = C.c_void_p()
        assert capi.lib().b200_mem_align_batch(idx.h, C.byref(opt), b.n, b.seq, b.seq_off, ids.ctypes.data_as(C.c_void_p), C.byref(h)) == 0
        t2 = time.perf_counter()
        from seqlib_b200.abi import ResultsView
        v = ResultsView()
        capi.lib().b200_results_view(h, C.byref(v))
        rn = (C.c_char_p * len(rnames))(*[r.encode() for r in rnames])
        sam.results_to_sam_flat  # binds argtypes lazily below
        txt = C.c_void_p(); ln = C.c_int64()
        Ls = capi.lib()
        Ls.b200_results_to_sam.argtypes = [C.POINTER(ResultsView), C.c_void_p, C.POINTER(C.c_char_p), C.c_int] + [C.c_void_p] * 8 + \
            [C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
        assert Ls.b200_results_to_sam(C.byref(v), C.byref(opt), rn, len(rnames), b.seq, b.seq_off, b.qual, b.qual_off, b.name, b.name_off,
                                      b.comment, b.comment_off, C.byref(txt), C.byref(ln)) == 0
        t3 = time.perf_counter()
        if args.out and it == 1:
            open(args.out, "wb").write(C.string_at(txt, ln.value))
        C.CDLL(None).free(txt)
        capi.lib().b200_results_free(h)
        best = {"parse_s": t1 - t0, "align_s": t2 - t1, "sam_s": t3 - t2, "total_s": t3 - t0, "sam_mb": ln.value / 1e6}
    out["b200"] = dict(best, reads_per_s=args.reads / best["total_s"])
    # the same chain with the stages overlapped: chunks of --chunk records, parse + align of chunk i+1 on this thread (GPU) while a
    # second thread formats chunk i (host cores); the synthetic records all have the same size, so chunk boundaries are exact
    import queue
    import threading
    from seqlib_b200.abi import ResultsView
    rec_bytes = len(text) // args.reads
    assert rec_bytes * args.reads == len(text)
    Ls = capi.lib()
    rn = (C.c_char_p * len(rnames))(*[r.encode() for r in rnames])
    for it in range(2):
        q = queue.Queue(maxsize=2)
        sam_bytes = [0]

        def format_worker():
            while True:
                item = q.get()
                if item is None:
                    return
                rd_i, b_i, h_i = item
                v = ResultsView()
                Ls.b200_results_view(h_i, C.byref(v))
                txt = C.c_void_p(); ln = C.c_int64()
                assert Ls.b200_results_to_sam(C.byref(v), C.byref(opt), rn, len(rnames), b_i.seq, b_i.seq_off, b_i.qual, b_i.qual_off,
                                              b_i.name, b_i.name_off, b_i.comment, b_i.comment_off, C.byref(txt), C.byref(ln)) == 0
                sam_bytes[0] += ln.value
                C.CDLL(None).free(txt)
                Ls.b200_results_free(h_i)
                rd_i.close()
        th = threading.Thread(target=format_worker)
        t0 = time.perf_counter()
        th.start()
        for r0 in range(0, args.reads, args.chunk):
            nrec = min(args.chunk, args.reads - r0)
            rd_i = fastq.FastqReader(text=b"")
            b_i = FastqBatch()
            sub = a[r0 * rec_bytes:(r0 + nrec) * rec_bytes]
            assert L.b200_fastq_parse_device(rd_i.h, sub.ctypes.data_as(C.c_void_p), len(sub), C.byref(b_i)) == 0 and b_i.n == nrec
            h_i = C.c_void_p()
            assert Ls.b200_mem_align_batch(idx.h, C.byref(opt), b_i.n, b_i.seq, b_i.seq_off, ids[r0:r0 + nrec].ctypes.data_as(C.c_void_p), C.byref(h_i)) == 0
            q.put((rd_i, b_i, h_i))
        q.put(None)
        th.join()
        dt = time.perf_counter() - t0
        out["b200_overlapped"] = {"chunk_records": args.chunk, "total_s": dt, "reads_per_s": args.reads / dt, "sam_mb": sam_bytes[0] / 1e6}
    from oracle import pyref
    if pyref.have_ref():
        cores = os.cpu_count() or 1
        ridx = pyref.RefIndex.from_view(idx.view(), keep=idx)
        ropt = pyref.default_opt()
        probe = 4000
        t = pyref.process_seqs(ridx, (seqs[:probe * 150], off[:probe + 1]), ropt, cores)
        n = int(min(args.reads, max(probe, probe / t * args.cpu_seconds)))
        t = pyref.process_seqs(ridx, (seqs[:n * 150], off[:n + 1]), ropt, cores)
        out["reference"] = {"reads_per_s": n / t, "sample_reads": n, "seconds": t, "cores": cores,
                            "what": "mem_process_seqs (bwa/bwamem.c:1235-1264: align + mem_reg2sam) on reads already in memory; FASTQ parsing not included"}
        out["speedup"] = out["b200"]["reads_per_s"] / out["reference"]["reads_per_s"]
        out["speedup_overlapped"] = out["b200_overlapped"]["reads_per_s"] / out["reference"]["reads_per_s"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
