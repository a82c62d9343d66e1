set -x
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_c2_v1.json 2> gpurun_out/r02_bench_c2_v1.err
tail -3 gpurun_out/r02_bench_c2_v1.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench_c2_v1.json"))
print({k: d[k] for k in ["value", "ms_per_step", "stage_ms_per_step", "e2e", "parity", "occ_blocks_per_read", "seed_table_lookups_per_read", "cpu_baseline", "clocks", "gpu_launches"]})
print(d["roofline"])
print(json.dumps(d.get("extra"))[:2500])
PY
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err; cut -c1-400 gpurun_out/r02_bench_reference_arm.json
