# ncu of the fermi-lite half: launch list of one assembly + --set full of the correction and overlap-record kernels
set -x
mkdir -p gpurun_out
READS=${READS:-300000}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_asm.csv python scripts/bench_asm.py --reads $READS --steps 1 --warmup 0 > gpurun_out/ncu_asm_launch.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ec -c 1 -o gpurun_out/prof_k_ec python scripts/bench_asm.py --reads $READS --steps 1 --warmup 0 > gpurun_out/prof_k_ec.log 2>&1; echo "k_ec rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_utg_nodes -c 1 -o gpurun_out/prof_k_utg python scripts/bench_asm.py --reads $READS --steps 1 --warmup 0 > gpurun_out/prof_k_utg.log 2>&1; echo "k_utg rc=$?"
ls -la gpurun_out/*.ncu-rep
