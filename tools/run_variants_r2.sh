#!/bin/bash
# runs each variant .so with optional env overrides; prints Msamples/s
cd /root/repo
out=gpurun_out/r2b; mkdir -p $out
run() { # name so env...
  name=$1; so=$2; shift 2
  r=$(env "$@" CORNELIS_CUDA_LIB=$so python bench.py --no-cpu-baseline --no-configs --steps 2 --warmup 1 2>>$out/variants.err | tail -1)
  echo "$name $(echo "$r" | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "sm_mhz", d["clocks"]["sm_mhz"])' 2>&1 | tail -1)" | tee -a $out/variants.log
}
V=cornelis_b200/lib/variants
run base $V/base.so X=1
run mb7_auto $V/mb7.so X=1
run mb7_7persm $V/mb7.so CORNELIS_PERSISTENT_BLOCKS_PER_SM=7
run mb7_6persm $V/mb7.so CORNELIS_PERSISTENT_BLOCKS_PER_SM=6
run mb7_philox10 $V/mb7_philox10.so X=1
run mb7_unroll2 $V/mb7_unroll2.so X=1
