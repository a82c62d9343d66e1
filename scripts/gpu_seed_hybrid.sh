set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_align.py -x -q 2>&1 | tail -5
REF=400000000 READS=8000000 python scripts/e2e_probe.py 2>&1 | grep -E "iter [23]|device-resident" | tail -4
python bench.py --steps 3 --warmup 2 --cpu-seconds 3 > gpurun_out/r02_bench_hybrid.json 2> gpurun_out/r02_bench_hybrid.err
tail -3 gpurun_out/r02_bench_hybrid.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench_hybrid.json"))
print({k: d[k] for k in ["value", "ms_per_step", "stage_ms_per_step", "e2e", "parity", "occ_blocks_per_read", "seed_table_lookups_per_read"]})
print(d["roofline"])
print(json.dumps(d.get("extra"))[:3000])
PY
