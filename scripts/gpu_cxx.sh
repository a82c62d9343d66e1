# the C++ drop-in tests + the alignSequences throughput on the 3 Gb index
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "cxx or wrapper or kat" 2>&1 | tail -3
python - <<'PY'
import json, sys, types
sys.argv = ["bench.py"]
import bench
a = types.SimpleNamespace(ref_len=3_000_000_000)
print(json.dumps(bench.cxx_extra(a)))
PY
