# wavefront extension kernel: parity tests, then A/B of lanes per read
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_align.py -x -q 2>&1 | tail -5
python bench.py --steps 2 --warmup 2 --no-extra --cpu-seconds 3 > gpurun_out/r02_bench_wave_g4.json 2> gpurun_out/r02_bench_wave_g4.err
tail -3 gpurun_out/r02_bench_wave_g4.err
export B200_BENCH_READS=4000000
for g in 8 2 0; do
  B200_WAVE_G=$g python bench.py --steps 2 --warmup 1 --no-extra --no-cpu-baseline > gpurun_out/r02_bench_wave_g$g.json 2> gpurun_out/r02_bench_wave_g$g.err
done
python - <<'PY'
import json
for g in (4, 8, 2, 0):
    try:
        d = json.load(open("gpurun_out/r02_bench_wave_g%d.json" % g))
        print(g, d["value"], d["stage_ms_per_step"], d.get("parity"), d["sw_cells_per_step"], d["config"]["reads_per_gpu_per_step"])
    except Exception as e:
        print(g, "failed", e)
PY
for w in 4 8 0; do B200_KSW_WAVE=$w python bench.py --workload ksw --steps 3 --warmup 2 --no-cpu-baseline; done
