#!/bin/bash
tag=${1:-run}; out=gpurun_out/$tag; mkdir -p $out; export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_backward.py -q -rA --timeout 300 --timeout-method thread > $out/pytest_bwd.log 2>&1; echo "pytest bwd rc=$?"
grep -E "passed|failed|FAILED|AssertionError|^\[f16|grad engine f16" $out/pytest_bwd.log | tail -80
for w in relation learn_nms; do
  timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $out/launches_bwd_$w.csv python tools/bwd_one.py $w f16 > $out/bwd_$w.log 2>&1; echo "ncu $w rc=$?"
done
timeout 300 python tools/backward_bench.py 2>/dev/null | head -2 | tee $out/backward_bench.jsonl
