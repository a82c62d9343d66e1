set -x
mkdir -p gpurun_out
export B200_BENCH_READS=2000000
for cfg in "B200_SEED_WINDOW=10 B200_TRACE=1" "B200_SEED_WINDOW=11" "B200_SEED_WINDOW=9"; do
  echo "== $cfg"
  env $cfg timeout 600 ncu --metrics dram__sectors_read.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:k_seed2 -s 1 -c 1 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra 2>&1 | grep -E "dram__sectors|gpu__time|lts__t|access policy"
done
echo "== gather_bw under ncu"
timeout 300 ncu --metrics dram__sectors_read.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,smsp__inst_executed.sum --clock-control none -k regex:k_gather -s 10 -c 6 ./scripts/microbench/gather_bw 2>&1 | grep -E "k_gather|dram__sectors|gpu__time|lts__t|inst_exec|bytes_per_gather" | head -60
