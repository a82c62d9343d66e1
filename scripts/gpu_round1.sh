set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/gpu.txt
(time python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
(time python bench.py) > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench rc=$?"
cat gpurun_out/bench_c2.json
tail -5 gpurun_out/bench_c2.err
export B200_BENCH_READS=1000000
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "ncu1 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_stage -s 4 -c 2 -o gpurun_out/prof_seed_r1 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu2 rc=$?"
tail -3 gpurun_out/ncu_full.log
