set -x
mkdir -p gpurun_out
TAG=${TAG:-r02_bench_c2_v3}
python bench.py --steps 10 --warmup 3 > gpurun_out/$TAG.json 2> gpurun_out/$TAG.err
tail -3 gpurun_out/$TAG.err
python - <<PY
import json
d = json.load(open("gpurun_out/$TAG.json"))
print({k: d.get(k) for k in ["value", "ms_per_step", "stage_ms_per_step", "e2e", "parity", "seed_gathers_per_read", "cpu_baseline", "clocks", "gpu_launches"]})
print(d["roofline"]); print(d["roofline_hbm"])
print(json.dumps(d.get("extra"))[:3000])
PY
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_reference_arm.json 2> gpurun_out/${TAG}_reference_arm.err; cut -c1-400 gpurun_out/${TAG}_reference_arm.json
