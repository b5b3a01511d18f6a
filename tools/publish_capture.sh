#!/bin/bash
# Condenses gpurun_out/r2l (tools/capture_persistent.sh) into the committed files under profiles/r2_ncu.
set -e
cd "$(dirname "$0")/.."
rays=$(python -c "import ast,sys; print(ast.literal_eval(open('gpurun_out/r2l/plain.log').read().split('} Msamples')[0]+'}')['rays'])")
python tools/ncu_json.py gpurun_out/r2l/persist_raw.csv k_persistent_queued --units $rays --command "ncu --set full --clock-control none --import-source on -k regex:k_persistent -s 1 -c 1 python tools/profile_run.py 512 (Cornell 1920x1080, 512 spp, one launch; tools/capture_persistent.sh)" > profiles/r2_ncu/k_persistent_queued.json
python tools/ncu_summary.py gpurun_out/r2l/persist_raw.csv k_persistent > profiles/r2_ncu/ncu_full_summary_k_persistent_queued_cornell_512spp.txt
python tools/ncu_sass_mix.py gpurun_out/r2l/persist_sass.csv $rays > profiles/r2_ncu/sass_mix_k_persistent_queued_cornell_512spp.txt
python tools/ncu_lines.py gpurun_out/r2l/persist_cs.csv 60 > profiles/r2_ncu/hot_lines_k_persistent_queued_cornell_512spp.txt
cp gpurun_out/r2l/launches_bench.csv profiles/r2_ncu/launches_bench.csv
cp gpurun_out/r2l/bench_plain.json profiles/r2_ncu/bench_plain.json
python tools/sass_histogram.py > profiles/r2_sass/opcodes.txt
python - <<'PY'
import json
d=json.load(open('profiles/r2_ncu/k_persistent_queued.json'))
print(d['git_head'][:10], d['csrc_sha256'][:12], d['launch'], d['issue']['warp_instructions_per_unit'], d['issue']['busy_pct'])
b=json.load(open('profiles/r2_ncu/bench_plain.json')); print(b['value'], b['roofline']['frac'], b['roofline'].get('issue_slots',{}).get('stale'))
PY
