# compute-sanitizer evidence for the round-2 kernels: memcheck on the alignment parity tests that touch every new path
# (prefix-interval tables, k_seed2/k_seed3, wavefront extension + hand-backs, two-tier spill), racecheck on the wavefront's
# shared-memory column stream (ksw batch test).
set -x
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_align.py -x -q -k "config1 or kat or ksw_extend2 or small_chunks or reads_with_n" > gpurun_out/r02_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -5 gpurun_out/r02_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_align.py -x -q -k "config1" > gpurun_out/r02_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"
tail -5 gpurun_out/r02_sanitizer_racecheck.log
