set -x
mkdir -p gpurun_out
for cfg in "B200_CHUNK_TAPER=0 B200_CHUNK=1048576" "B200_CHUNK_TAPER=0 B200_CHUNK=2097152" "B200_CHUNK_TAPER=1 B200_CHUNK=2097152" "B200_CHUNK_TAPER=0 B200_CHUNK=4194304" "B200_CHUNK_TAPER=1 B200_CHUNK=4194304"; do
  echo "== $cfg"
  env $cfg REF=400000000 READS=8000000 python scripts/e2e_probe.py 2>&1 | grep -E "iter [23]|device-resident" | tail -5
done
