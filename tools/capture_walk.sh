#!/bin/bash
# ncu --set full on ONE steady-state k_walk launch of config 4 (camera-ray packets and pooled bounce rays in the same pass):
# usage on the GPU box: tools/capture_walk.sh <gpurun_out subdirectory>
out=gpurun_out/${1:-walk}; mkdir -p $out
python tools/profile_run.py 64 --scene config4 --pipeline wavefront > $out/plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_walk -s 12 -c 1 -o $out/walk -f \
    python tools/profile_run.py 64 --scene config4 --pipeline wavefront > $out/ncu.log 2>&1
ncu -i $out/walk.ncu-rep --page raw --csv > $out/walk_raw.csv 2>/dev/null
ncu -i $out/walk.ncu-rep --page source --csv --print-source cuda,sass > $out/walk_cs.csv 2>/dev/null
cat $out/plain.log
