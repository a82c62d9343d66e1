"""Do two chunks in flight (two host threads, each with its own engine / stream) overlap the memory-bound seeding of one
chunk with the ALU-bound extension of the other?  Aggregate device-resident throughput of T threads x READS/T reads."""
import os, sys, time, threading
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from seqlib_b200 import capi, synth
capi.set_device(0)
L = int(os.environ.get("REF", 400_000_000)); n = int(os.environ.get("READS", 4_000_000)); T = int(os.environ.get("THREADS", 2))
pac = synth.reference(L); ctg = synth.contigs_for(L, 4)
seqs, off, _, _ = synth.reads(pac, L, ctg, n, 150, 0.01, 0.0)
idx = capi.Index.construct_pac(pac, L, ctg, keep_host=False)
opt = capi.default_opt()
ids = np.arange(n, dtype=np.int64) * 7919 + 13
per = n // T
def worker(t, res):
    capi.set_device(0)
    b = capi.Batch(idx, (seqs[t * per * 150:(t + 1) * per * 150], off[:per + 1]), opt, ids[t * per:(t + 1) * per])
    b.run()
    bar.wait()
    t0 = time.perf_counter()
    for _ in range(3):
        b.run()
    res[t] = (time.perf_counter() - t0) / 3
    bar.wait()
    b.close()
bar = threading.Barrier(T)
res = [0.0] * T
th = [threading.Thread(target=worker, args=(t, res)) for t in range(T)]
for x in th: x.start()
for x in th: x.join()
print("threads %d: %.1f ms per pass of %d reads -> %.2f M reads/s aggregate" % (T, 1e3 * max(res), n, n / max(res) / 1e6), flush=True)
