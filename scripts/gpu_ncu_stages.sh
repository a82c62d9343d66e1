# ncu --set full of the thread-per-read stages (chain, finalize) and the CIGAR kernel: one launch each
set -x
mkdir -p gpurun_out
export B200_BENCH_READS=1000000
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_stage|k_finalize_dp" -s 4 -c 4 -o gpurun_out/${TAG}_stages python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra --parity-reads 1000 > gpurun_out/${TAG}_stages.log 2>&1; echo "ncu rc=$?"
