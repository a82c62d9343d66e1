# A/B of environment-selected kernel variants: short bench runs, stage times per 10 M reads.  VARIANTS="A=1 B=2|C=3|..."
set -x
mkdir -p gpurun_out
IFS='|' read -ra VS <<< "$VARIANTS"
i=0
for v in "${VS[@]}"; do
  env $v python bench.py --no-extra --no-cpu-baseline --steps 2 --warmup 2 --parity-reads 1000 > gpurun_out/${TAG}_$i.json 2> gpurun_out/${TAG}_$i.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_$i.json").read().strip().splitlines()[-1])
    print("RESULT [$v] value %.2f M e2e %.2f M" % (d["value"]/1e6, d["e2e"]["value"]/1e6), {k: round(x,1) for k,x in d["stage_ms_per_step"].items()}, "truth", d["parity"]["truth_within_8bp"], "hits", d["hits_per_step"], "spill", d.get("spill_reads_per_step"))
except Exception as e:
    print("RESULT [$v] failed", e)
PY
  i=$((i+1))
done
