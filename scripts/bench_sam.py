#!/usr/bin/env python
"""Result materialisation rate: b200_results_to_sam over the hits of N synthetic 150-bp reads aligned to a
100 Mb random reference (all host threads).  Prints one JSON object."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000000
    from seqlib_b200 import capi, synth, sam
    capi.set_device(0)
    l_pac = 100000000
    pac = synth.reference(l_pac)
    ctg = synth.contigs_for(l_pac, 4, "chr")
    idx = capi.Index.construct_pac(pac, l_pac, ctg)
    seqs, off, _, _ = synth.reads(pac, l_pac, ctg, n, 150, 0.01, 5e-4)
    opt = capi.default_opt()
    ids = np.arange(n, dtype=np.int64)
    t = time.perf_counter()
    res = capi.align(idx, (seqs, off), opt, ids)
    t_align = time.perf_counter() - t
    digits = 8
    nm = np.empty((n, 1 + digits), dtype=np.uint8); nm[:, 0] = ord("r")
    ii = np.arange(n)
    for d in range(digits):
        nm[:, 1 + d] = 48 + (ii // 10 ** (digits - 1 - d)) % 10
    names = (nm.reshape(-1), np.arange(n + 1, dtype=np.int64) * (1 + digits))
    quals = (np.full(int(off[-1]), ord("I"), dtype=np.uint8), np.ascontiguousarray(off, dtype=np.int64))
    rnames = [idx.seq_name(i) for i in range(idx.n_seqs())]
    tm = {}
    text = sam.results_to_sam_flat(res, opt, rnames, seqs, off, names, quals, None, timing=tm)
    print(json.dumps({"reads": n, "sam_bytes": len(text), "records": text.count(b"\n"), "align_s_incl_numpy": t_align,
                      "to_sam_s": tm["seconds"], "reads_per_s": n / tm["seconds"], "gb_per_s": len(text) / tm["seconds"] / 1e9,
                      "cores": os.cpu_count()}))


if __name__ == "__main__":
    main()
