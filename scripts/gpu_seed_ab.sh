# A/B of seeding kernel variants: parity tests, then short bench runs (stage times per 10 M reads)
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_align.py -x -q 2>&1 | tail -4
for v in 5 6; do
  B200_SEED_MINB=$v python bench.py --no-extra --no-cpu-baseline --steps 2 --warmup 2 --parity-reads 200000 > gpurun_out/${TAG}_minb$v.json 2> gpurun_out/${TAG}_minb$v.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_minb$v.json").read().strip().splitlines()[-1])
print("MINB=$v value %.2f M e2e %.2f M" % (d["value"]/1e6, d["e2e"]["value"]/1e6), d["stage_ms_per_step"], d["parity"]["mismatches"], d.get("occ_blocks_per_read"), d.get("seed_table_lookups_per_read"))
PY
done
