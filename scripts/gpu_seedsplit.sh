set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_align.py -x -q 2>&1 | tail -3
export B200_BENCH_READS=4000000
python bench.py --steps 2 --warmup 1 --no-extra --no-cpu-baseline > gpurun_out/r02_bench_seedsplit.json 2> gpurun_out/r02_bench_seedsplit.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench_seedsplit.json"))
print(d["value"], d["stage_ms_per_step"], d["spill_reads_per_step"], d["occ_blocks_per_read"])
PY
