# GPU parity tests of the fermi-lite half, then the config-4 shaped BFC measurement
set -x
mkdir -p gpurun_out
(time python -m pytest tests/test_gpu_fml.py -m gpu -x -q) > gpurun_out/pytest_gpu_fml.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/pytest_gpu_fml.log
python scripts/bench_fml.py --reads ${READS:-200000} --region ${REGION:-200000} --steps 2 --warmup 1 --check 1 > gpurun_out/bench_fml_small.json 2> gpurun_out/bench_fml_small.err; echo "small rc=$?"
tail -3 gpurun_out/bench_fml_small.err; cat gpurun_out/bench_fml_small.json
timeout 600 python scripts/bench_fml.py --reads 1000000 --region 1000000 --steps 2 --warmup 1 --cpu-sample 20000 > gpurun_out/bench_fml_c4.json 2> gpurun_out/bench_fml_c4.err; echo "c4 rc=$?"
tail -3 gpurun_out/bench_fml_c4.err; cat gpurun_out/bench_fml_c4.json
