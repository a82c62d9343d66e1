set -x
mkdir -p gpurun_out
export B200_BENCH_READS=2000000
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_2M.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/r02_launches_2M.log 2>&1
python scripts/ncu_summary.py launches gpurun_out/r02_launches_2M.csv | head -40
