#!/bin/bash
# End-of-session validation on the GPU box, most important first (a call that runs out of its time limit keeps what is done):
# the ncu capture that bench.py's roofline block reads (tools/capture_persistent.sh -> tools/publish_capture.sh; bench.py
# marks those figures stale when cornelis_b200/csrc has changed since), the parity suite, smoke, the default bench line,
# the wavefront bench line, one k_walk capture.   usage: tools/final_run.sh <gpurun_out subdirectory>
out=gpurun_out/${1:-final}; mkdir -p $out
bash tools/capture_persistent.sh > $out/capture_persistent.log 2>&1; tail -1 $out/capture_persistent.log
python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; tail -1 $out/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; tail -1 $out/smoke.log
python bench.py > $out/bench_default.json 2> $out/bench_default.err; tail -c 600 $out/bench_default.json
python bench.py --pipeline wavefront --no-configs --no-cpu-baseline > $out/bench_wavefront.json 2> $out/bench_wavefront.err; tail -c 300 $out/bench_wavefront.json
bash tools/capture_walk.sh $(basename $out)/walk > $out/capture_walk.log 2>&1; tail -1 $out/capture_walk.log
