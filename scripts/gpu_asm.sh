# GPU parity tests of the fermi-lite half (BFC + assembly), then the config-4 shaped assembly measurement
set -x
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_fml.py tests/test_gpu_asm.py -m gpu -x -q) > gpurun_out/pytest_gpu_asm.log 2>&1; echo "pytest rc=$?"
tail -40 gpurun_out/pytest_gpu_asm.log
timeout 600 python scripts/bench_asm.py --reads 100000 --steps 1 --warmup 1 --ref-reads 100000 --check > gpurun_out/bench_asm_100k.json 2> gpurun_out/bench_asm_100k.err; echo "100k rc=$?"
tail -3 gpurun_out/bench_asm_100k.err; cat gpurun_out/bench_asm_100k.json
timeout 900 python scripts/bench_asm.py --reads 1000000 --steps 1 --warmup 1 > gpurun_out/bench_asm_c4.json 2> gpurun_out/bench_asm_c4.err; echo "c4 rc=$?"
tail -3 gpurun_out/bench_asm_c4.err; cat gpurun_out/bench_asm_c4.json
