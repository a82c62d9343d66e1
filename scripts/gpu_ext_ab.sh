run() { echo "== $*"; env "$@" B200_BENCH_READS=2000000 python bench.py --steps 2 --warmup 1 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.3fM/s' % (d['value']/1e6), d['stage_ms_per_step'], 'spill', d['spill_reads_per_step'])"; }
run X=1
run B200_EXTEND_SMEM=1
