#!/bin/bash
# second batch: tf32 GEMM + backward + pipeline tests, training-side / deformable micro-benches, then the wide-head fused tests
tag=${1:-run}; out=gpurun_out/$tag; mkdir -p $out; export PYTHONUNBUFFERED=1
timeout 700 python -m pytest tests/test_gpu_backward.py tests/test_gpu_pipelines.py -q -rA --timeout 300 --timeout-method thread > $out/pytest_bwd_pipe.log 2>&1; echo "pytest bwd/pipelines rc=$?"
grep -E "passed|failed|rel err|FAILED|^\[|fpn|faster|deformable" $out/pytest_bwd_pipe.log | tail -60
timeout 300 python tools/backward_bench.py > $out/backward_bench.jsonl 2> $out/backward_bench.err; echo "bwd bench rc=$?"; cat $out/backward_bench.jsonl; tail -3 $out/backward_bench.err
timeout 200 python tools/deform_bench.py > $out/deform_bench.jsonl 2> $out/deform_bench.err; cat $out/deform_bench.jsonl
timeout 200 python tools/psroi_bench.py > $out/psroi_bench.jsonl 2> $out/psroi_bench.err; cat $out/psroi_bench.jsonl
timeout 400 python -m pytest tests/test_gpu_fused.py -q -rA --timeout 120 --timeout-method thread > $out/pytest_fused.log 2>&1; echo "pytest fused rc=$?"
grep -E "passed|failed|fused|FAILED|Error" $out/pytest_fused.log | tail -40
