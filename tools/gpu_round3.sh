#!/bin/bash
tag=${1:-run}; out=gpurun_out/$tag; mkdir -p $out; export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_backward.py -q -rA --timeout 300 --timeout-method thread > $out/pytest_bwd.log 2>&1; echo "pytest bwd rc=$?"
grep -E "passed|failed|rel err|FAILED|^\[f16\]" $out/pytest_bwd.log | tail -70
for w in relation learn_nms; do
  timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $out/launches_bwd_$w.csv python tools/bwd_one.py $w f16 > $out/bwd_$w.log 2>&1; echo "ncu $w rc=$?"
done
timeout 200 python tools/timeline.py > $out/timeline.log 2>&1; cp gpurun_out/timeline.json $out/timeline.json; tail -2 $out/timeline.log
