# A/B of the one-lane-per-read extension kernel (extend_lane.cuh) against the group kernel, then parity tests
set -x
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" python bench.py --reads 2000000 --steps 2 --warmup 1 --no-cpu-baseline --no-extra 2>gpurun_out/lane_ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.3fM/s' % (d['value']/1e6), d['stage_ms_per_step'], 'spill', d['spill_reads_per_step'], 'cells', d['sw_cells_per_step'], 'hits', d['hits_per_step'])"; tail -2 gpurun_out/lane_ab.err; }
(time timeout 900 python -m pytest tests/test_gpu_align.py -m gpu -x -q) > gpurun_out/pytest_gpu_lane.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/pytest_gpu_lane.log
run X=1
run B200_LANE_CBATCH=1
run B200_LANE_CBATCH=8
run B200_LANE_BATCH=8
run B200_LANE_BATCH=2
run B200_EXTEND_GROUP=1
