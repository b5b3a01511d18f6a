#!/bin/bash
# ncu --set full on every stage kernel of ONE steady-state wavefront pass (Cornell and config 4) and on the batch kernel.
mkdir -p gpurun_out/r2x
python tools/profile_run.py 64 --pipeline wavefront > gpurun_out/r2x/plain_cornell.log 2>&1 || exit 1
python tools/profile_run.py 64 --scene config4 --pipeline wavefront > gpurun_out/r2x/plain_config4.log 2>&1 || exit 1
# pass 8 of the second render: launches are plan, raygen, intersect, shade, accumulate (x passes); skip well into the run
ncu --set full --clock-control none -k regex:'k_raygen|k_intersect|k_shade|k_accumulate|k_resolve' -s 20 -c 5 -o gpurun_out/r2x/wave_cornell -f python tools/profile_run.py 64 --pipeline wavefront > gpurun_out/r2x/ncu1.log 2>&1
ncu --set full --clock-control none -k regex:'k_raygen|k_walk|k_compact_hits|k_shade|k_accumulate' -s 20 -c 5 -o gpurun_out/r2x/wave_config4 -f python tools/profile_run.py 64 --scene config4 --pipeline wavefront > gpurun_out/r2x/ncu2.log 2>&1
ncu --set full --clock-control none -k regex:k_intersect_batch -s 3 -c 1 -o gpurun_out/r2x/batch -f python tools/microbench_intersect.py 22 > gpurun_out/r2x/ncu3.log 2>&1
for f in wave_cornell wave_config4 batch; do ncu -i gpurun_out/r2x/$f.ncu-rep --page raw --csv > gpurun_out/r2x/${f}_raw.csv 2>/dev/null; done
cat gpurun_out/r2x/plain_cornell.log gpurun_out/r2x/plain_config4.log
