# ncu --set full of one kernel: KERNEL=regex SKIP=n OUT=name
set -x
mkdir -p gpurun_out
export B200_BENCH_READS=1000000
timeout 900 ncu --set full --clock-control none --import-source on -k regex:${KERNEL} -s ${SKIP:-1} -c 1 -o gpurun_out/${OUT} python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${OUT}.log 2>&1; echo "ncu rc=$?"
