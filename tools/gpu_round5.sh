#!/bin/bash
tag=${1:-run}; out=gpurun_out/$tag; mkdir -p $out; export PYTHONUNBUFFERED=1
timeout 200 python tools/bwd_probe.py 70 256 4 50 1 2>&1 | tail -14
timeout 200 python tools/bwd_probe.py 131 256 16 131 1 2>&1 | tail -12
timeout 400 compute-sanitizer --tool memcheck --print-limit 20 python tools/bwd_probe.py 70 256 4 50 1 > $out/memcheck_bwd.log 2>&1; grep -E "ERROR SUMMARY|Invalid|at .*\+0x|by thread" $out/memcheck_bwd.log | head -20
timeout 600 python -m pytest tests/test_gpu_backward.py -q -rA --timeout 300 --timeout-method thread -k "relation_backward or learn_nms_backward or small_head or bucket or gemm" > $out/pytest_bwd.log 2>&1; echo "pytest bwd rc=$?"
grep -E "passed|failed|FAILED|AssertionError" $out/pytest_bwd.log | tail -20
timeout 300 python tools/backward_bench.py 2>/dev/null | head -2
