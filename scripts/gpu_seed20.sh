set -x
export B200_BENCH_READS=4000000
for cap in 20 16; do
B200_SEED_CAP=$cap python bench.py --steps 2 --warmup 1 --no-extra --no-cpu-baseline > gpurun_out/r02_bench_seedcap$cap.json 2> gpurun_out/r02_bench_seedcap$cap.err
python - <<PY
import json
d = json.load(open("gpurun_out/r02_bench_seedcap$cap.json"))
print($cap, d["value"], d["stage_ms_per_step"], d["spill_reads_per_step"], d["occ_blocks_per_read"])
PY
done
