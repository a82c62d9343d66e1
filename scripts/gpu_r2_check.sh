set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_align.py -x -q 2>&1 | tail -15
./scripts/microbench/int_peak > gpurun_out/r02_int_peak.json 2>&1; cat gpurun_out/r02_int_peak.json
python bench.py --workload ksw --steps 3 --warmup 2 > gpurun_out/r02_bench_ksw.json 2> gpurun_out/r02_bench_ksw.err; cat gpurun_out/r02_bench_ksw.json; tail -3 gpurun_out/r02_bench_ksw.err
