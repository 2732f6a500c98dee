#!/bin/bash
tag=${1:-run}; out=gpurun_out/$tag; mkdir -p $out; export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_backward.py tests/test_gpu_parity.py -q -rA --timeout 300 --timeout-method thread -k "tf32 or relation_backward or learn_nms_backward or small_head or bucket" > $out/pytest_bwd.log 2>&1; echo "pytest bwd rc=$?"
grep -E "passed|failed|rel err|FAILED|^\[f16\]|tf32 forward|AssertionError" $out/pytest_bwd.log | tail -80
