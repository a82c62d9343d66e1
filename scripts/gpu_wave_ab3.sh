set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_align.py -x -q 2>&1 | tail -3
export B200_BENCH_READS=4000000
for g in 4 8; do
  B200_WAVE_G=$g python bench.py --steps 2 --warmup 1 --no-extra --no-cpu-baseline > gpurun_out/r02_bench_wave3_g$g.json 2> gpurun_out/r02_bench_wave3_g$g.err
done
python - <<'PY'
import json
for g in (4, 8):
    try:
        d = json.load(open("gpurun_out/r02_bench_wave3_g%d.json" % g))
        print(g, d["value"], d["stage_ms_per_step"], d["sw_cells_per_step"])
    except Exception as e:
        print(g, "failed", e)
PY
B200_BENCH_READS=1000000 timeout 600 ncu --metrics smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,gpu__time_duration.sum,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_extend_wave -s 1 -c 1 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra 2>&1 | grep -E "thread_inst|inst_executed.sum|gpu__time|pipe_alu"
python bench.py --workload ksw --steps 3 --warmup 2 --no-cpu-baseline | cut -c1-160
