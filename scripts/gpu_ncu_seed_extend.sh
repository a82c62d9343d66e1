# one ncu --set full capture of the seeding kernel and of the extension kernel (1M reads vs the 3 Gb index)
set -x
mkdir -p gpurun_out
export B200_BENCH_READS=1000000
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_stage -s 3 -c 1 -o gpurun_out/prof_seed python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_seed.log 2>&1; echo "ncu seed rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_extend_group -s 1 -c 1 -o gpurun_out/prof_extend python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_extend.log 2>&1; echo "ncu extend rc=$?"
B200_SEED_FSM=1 B200_BENCH_READS=4000000 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_fsm.json 2> gpurun_out/bench_fsm.err; cat gpurun_out/bench_fsm.json
