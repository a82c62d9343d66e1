# launch list of the default pipeline (2 M reads per step) + ncu --set full of the two seeding kernels
set -x
mkdir -p gpurun_out
export B200_BENCH_READS=2000000
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_2M.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra --parity-reads 1000 > gpurun_out/${TAG}_launches_2M.log 2>&1
python scripts/ncu_summary.py launches gpurun_out/${TAG}_launches_2M.csv | head -24
export B200_BENCH_READS=1000000
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_seed2 -s 1 -c 1 -o gpurun_out/${TAG}_seed2 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra --parity-reads 1000 > gpurun_out/${TAG}_seed2.log 2>&1; echo "ncu rc=$?"
