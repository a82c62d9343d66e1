# ncu launch list of the bench command (reduced to one chunk of 1M reads per step) + --set full of the seeding and extension kernels
set -x
mkdir -p gpurun_out
export B200_BENCH_READS=1000000
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/ncu_bench_launch.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:^k_seed2$" -s 1 -c 1 -o gpurun_out/prof_seed2_r1b python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/prof_seed2_r1b.log 2>&1; echo "seed rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:^k_extend_group$" -s 1 -c 1 -o gpurun_out/prof_ext_r1b python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/prof_ext_r1b.log 2>&1; echo "ext rc=$?"
ls -la gpurun_out/*r1b*
