#!/bin/bash
tag=${1:-run}; out=gpurun_out/$tag; mkdir -p $out; export PYTHONUNBUFFERED=1
timeout 500 python -m pytest tests/test_gpu_refpin.py tests/test_gpu_parity.py tests/test_gpu_pipelines.py tests/test_gpu_trunk.py tests/test_gpu_backward.py -q -rA --timeout 300 --timeout-method thread -k "deform or dcn or linear or gemm" > $out/pytest_deform.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|FAILED" $out/pytest_deform.log | tail
timeout 200 python tools/deform_bench.py 2>/dev/null | tee $out/deform_bench.jsonl
