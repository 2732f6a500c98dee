#!/bin/bash
tag=${1:-run}; out=gpurun_out/$tag; mkdir -p $out; export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py tests/test_gpu_compat.py tests/test_gpu_pipelines.py tests/test_gpu_refpin.py -q -rA --timeout 200 --timeout-method thread -k "nms or proposal or pipeline or edge or compat" > $out/pytest_nms.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|FAILED" $out/pytest_nms.log | tail -8
timeout 200 python tools/timeline.py > $out/timeline.log 2>&1; cp gpurun_out/timeline.json $out/timeline.json
python - <<PY
import json
ev=json.load(open('$out/timeline.json'))
for e in ev:
    if any(k in e['name'] for k in ('nms_','proposal_','rpn_finish','roi_pool')): print(e['name'][:50], e['start_us'], e['dur_us'])
print('span', ev[-1]['start_us']+ev[-1]['dur_us'])
PY
timeout 600 python bench.py --no-sweep --no-configs --no-train 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value', d['value'], 'e2e', d['e2e']['value'], d['hot_path'])"
