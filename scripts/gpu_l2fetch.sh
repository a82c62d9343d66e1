# A/B: L2 fetch granularity hint and seed-table depth, seeding stage time + DRAM sectors of k_seed2
set -x
mkdir -p gpurun_out
export B200_BENCH_READS=2000000
for cfg in "default" "B200_L2_FETCH=32" "B200_L2_FETCH=128" "B200_SEED_TAB_K=13" "B200_SEED_TAB_K=12" "B200_SEED_TAB_K=10"; do
  echo "== $cfg"
  if [ "$cfg" = default ]; then E=""; else E="$cfg"; fi
  env $E timeout 600 ncu --metrics dram__sectors_read.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:k_seed2 -s 2 -c 1 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra 2>&1 | grep -E "dram__sectors|gpu__time|lts__t|ms_seed" | sed 's/.*stage_ms_per_step/stage/' | cut -c1-200
done
