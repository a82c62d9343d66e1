# ncu --set full of the BFC correction kernel proper (k_ec<5>, not k_ec_difficulty)
set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:^k_ec$" -c 1 -o gpurun_out/prof_k_ec python scripts/bench_asm.py --reads ${READS:-300000} --steps 1 --warmup 0 > gpurun_out/prof_k_ec.log 2>&1; echo "k_ec rc=$?"
ls -la gpurun_out/prof_k_ec.ncu-rep
