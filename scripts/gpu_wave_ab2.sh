set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_align.py -x -q 2>&1 | tail -3
export B200_BENCH_READS=4000000
for g in 8 4; do
  B200_WAVE_G=$g python bench.py --steps 2 --warmup 1 --no-extra --no-cpu-baseline > gpurun_out/r02_bench_wave2_g$g.json 2> gpurun_out/r02_bench_wave2_g$g.err
done
python - <<'PY'
import json
for g in (8, 4):
    try:
        d = json.load(open("gpurun_out/r02_bench_wave2_g%d.json" % g))
        print(g, d["value"], d["stage_ms_per_step"], d["sw_cells_per_step"])
    except Exception as e:
        print(g, "failed", e)
PY
for w in 4 8; do B200_KSW_WAVE=$w python bench.py --workload ksw --steps 3 --warmup 2 --no-cpu-baseline | cut -c1-200; done
