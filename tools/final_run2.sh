#!/bin/bash
# Shorter end-of-session pass after a change of csrc: the ncu capture first (bench.py marks its figures stale otherwise),
# then the parity suite, the wavefront line and the default line.
out=gpurun_out/${1:-final2}; mkdir -p $out
bash tools/capture_persistent.sh > $out/capture_persistent.log 2>&1; tail -1 $out/capture_persistent.log
python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; tail -1 $out/pytest.log
python bench.py --pipeline wavefront --no-configs --no-cpu-baseline > $out/bench_wavefront.json 2> $out/bench_wavefront.err; tail -c 200 $out/bench_wavefront.json
python bench.py > $out/bench_default.json 2> $out/bench_default.err; tail -c 300 $out/bench_default.json
