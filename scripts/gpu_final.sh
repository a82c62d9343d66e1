# full GPU suite + smoke + the default bench line + the reference arm (what the driver runs at round end)
set -x
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu_all.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu_all.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
(time python bench.py) > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
cat gpurun_out/bench_default.json | cut -c1-900
tail -3 gpurun_out/bench_default.err
python scripts/bench_sam.py 2000000 > gpurun_out/bench_sam.json 2> gpurun_out/bench_sam.err; cat gpurun_out/bench_sam.json
(time python bench.py --impl reference --steps 2 --warmup 1) > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_ref.json | cut -c1-600
