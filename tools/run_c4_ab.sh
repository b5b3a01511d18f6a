#!/bin/bash
# config 4 (10 000 spheres, wavefront + k_walk) once per A/B build in cornelis_b200/lib/variants (tools/variants.py build);
# usage on the GPU box: tools/run_c4_ab.sh <gpurun_out subdirectory> [spp]
cd /root/repo; out=gpurun_out/${1:-ab}; spp=${2:-256}; mkdir -p $out
for so in cornelis_b200/lib/variants/*.so; do
  name=$(basename $so .so)
  CORNELIS_CUDA_LIB=$PWD/$so python tools/bench_config4.py --no-cpu --no-exhaustive --spp $spp 2>>$out/err.log | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); w=j['wavefront']
print('$name', round(w['msamples_per_s'],1), 'Msamples/s intersect', round(w['stage_ms_per_pass']['intersect_ms'],3), 'ms/pass shade', round(w['stage_ms_per_pass']['shade_ms'],3), 'persistent', round(j['persistent']['msamples_per_s'],1))" | tee -a $out/c4.log
done
