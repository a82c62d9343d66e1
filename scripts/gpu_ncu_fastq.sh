# ncu launch list of the ingest path (device FASTQ parser): evidence for the kernels behind b200_fastq_parse_device
set -x
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_fastq.csv python scripts/bench_fastq.py --records 1000000 > gpurun_out/ncu_fastq.log 2>&1; echo "ncu rc=$?"
tail -2 gpurun_out/ncu_fastq.log
