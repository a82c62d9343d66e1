import os, sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from seqlib_b200 import capi
import cases, goldenlib
capi.set_device(0)
idx = capi.Index.load(goldenlib.path("tiny", "tiny.fa"))
reads = cases.read_lines(goldenlib.path("sim1_5k.txt"))
opt = capi.default_opt()
for bs in (8, 64, 1024):
    for i in range(3):
        capi.align(idx, reads[i*bs:(i+1)*bs], opt, np.arange(bs, dtype=np.int64))
    st = capi.last_stats()
    print(bs, {k: round(v, 3) if isinstance(v, float) else v for k, v in st.items()}, flush=True)
