#!/bin/bash
# A/B of the builds in cornelis_b200/lib/variants (tools/variants.py build ...) with optional environment overrides:
# one short bench.py run each, Msamples/s appended to gpurun_out/$1/variants.log
cd /root/repo
out=gpurun_out/${1:-r2}; mkdir -p $out
run() { # name so env...
  name=$1; so=$2; shift 2
  r=$(env "$@" CORNELIS_CUDA_LIB=$so python bench.py --no-cpu-baseline --no-configs --steps 4 --warmup 2 2>>$out/variants.err | tail -1)
  echo "$name $(echo "$r" | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "sm_mhz", d["clocks"]["sm_mhz"])' 2>&1 | tail -1)" | tee -a $out/variants.log
}
V=cornelis_b200/lib/variants
for so in $V/*.so; do run $(basename $so .so) $so X=1; done
