# GPU parity tests, then a short bench (4M reads vs the 3 Gb index)
set -x
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu.log
B200_BENCH_READS=${READS:-4000000} python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_quick.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_quick.json').read())
print("value %.3fM/s e2e %.3fM/s" % (d['value']/1e6, d['e2e']['value']/1e6), d['stage_ms_per_step'], "roofline", d['roofline']['achieved'], d['roofline']['frac'], "spill", d['spill_reads_per_step'])
PY
