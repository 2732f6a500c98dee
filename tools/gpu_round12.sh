#!/bin/bash
tag=${1:-run}; out=gpurun_out/$tag; mkdir -p $out; export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_fused.py tests/test_gpu_parity.py tests/test_gpu_pipelines.py tests/test_gpu_compat.py -q -rA --timeout 200 --timeout-method thread > $out/pytest_fwd.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|FAILED" $out/pytest_fwd.log | tail -8
timeout 300 python tools/fused_bench.py 300,1024,16 1000,1024,16 3000,1024,16 3000,1024,4 2>/dev/null | tee $out/fused_bench.jsonl
timeout 600 python bench.py --no-sweep --no-configs > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('$out/bench.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], d['hot_path'])
print('train', d['train']['ms_per_step'], d['train']['images_per_sec'], d['train']['launch'][:40])
PY
