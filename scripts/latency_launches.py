"""One 1024-read call (after two warm-up calls) for an ncu launch list of the small-batch pipeline."""
import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from seqlib_b200 import capi
import cases, goldenlib
capi.set_device(0)
idx = capi.Index.load(goldenlib.path("tiny", "tiny.fa"))
reads = cases.read_lines(goldenlib.path("sim1_5k.txt"))
opt = capi.default_opt()
bs = int(os.environ.get("BS", 1024))
for i in range(3):
    capi.align(idx, reads[i * bs:(i + 1) * bs], opt, np.arange(bs, dtype=np.int64))
print(capi.last_stats())
