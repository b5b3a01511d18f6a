#!/bin/bash
# End-of-session validation on the GPU box: parity suite, smoke, the default bench line, the ncu capture the bench's
# roofline block reads (tools/capture_persistent.sh -> tools/publish_capture.sh), the wavefront bench line, one k_walk capture.
out=gpurun_out/${1:-final}; mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; tail -2 $out/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; tail -1 $out/smoke.log
python bench.py > $out/bench_default.json 2> $out/bench_default.err; tail -c 600 $out/bench_default.json
bash tools/capture_persistent.sh > $out/capture_persistent.log 2>&1; tail -3 $out/capture_persistent.log
python bench.py --pipeline wavefront --no-configs --no-cpu-baseline > $out/bench_wavefront.json 2> $out/bench_wavefront.err; tail -c 300 $out/bench_wavefront.json
bash tools/capture_walk.sh $(basename $out)/walk > $out/capture_walk.log 2>&1; tail -1 $out/capture_walk.log
