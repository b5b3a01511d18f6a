#!/bin/bash
# config 4 (10 000 spheres, wavefront + k_walk) once per build in cornelis_b200/lib/variants and per environment override
cd /root/repo; out=gpurun_out/${1:-r2o}; mkdir -p $out
run() { name=$1; so=$2; shift 2
  env "$@" CORNELIS_CUDA_LIB=$PWD/$so python tools/bench_config4.py --no-cpu --no-exhaustive --spp 256 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); w=j['wavefront']
print('$name', round(w['msamples_per_s'],1), 'Msamples/s intersect', round(w['stage_ms_per_pass']['intersect_ms'],3), 'ms/pass shade', round(w['stage_ms_per_pass']['shade_ms'],3), 'persistent', round(j['persistent']['msamples_per_s'],1))" | tee -a $out/c4.log
}
for so in cornelis_b200/lib/variants/*.so; do
  run "$(basename $so .so)_sharedranges" $so CORNELIS_WALK_SHARED_RANGES=1
  run "$(basename $so .so)_globalranges" $so CORNELIS_WALK_SHARED_RANGES=0
done
