#!/bin/bash
# config 4 at 1024 spp with 2^24 / 2^25 / 2^26 paths in flight (CORNELIS_POOL_PATHS); usage: tools/run_c4_pools.sh <subdir>
cd /root/repo; out=gpurun_out/${1:-pools}; mkdir -p $out
for p in 16777216 33554432 67108864; do
  CORNELIS_POOL_PATHS=$p python tools/bench_config4.py --no-cpu --no-exhaustive --no-persistent --spp 1024 2>>$out/err.log | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); w=j['wavefront']
print('pool $p', round(w['msamples_per_s'],1), 'Msamples/s', round(w['gpu_ms'],1), 'ms intersect', round(w['stage_ms_per_pass']['intersect_ms'],3), 'ms/pass shade', round(w['stage_ms_per_pass']['shade_ms'],3), 'launches', w['kernel_launches'])" | tee -a $out/pools.log
done
