# config 5: the driver's 8-GPU launch of bench.py (weak scaling, 10 M reads per GPU)
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus_n8.txt
(time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 3 --warmup 3 --no-extra) > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "bench rc=$?"
cat gpurun_out/bench_n8.json | cut -c1-1200
tail -5 gpurun_out/bench_n8.err
