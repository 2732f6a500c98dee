#!/bin/bash
# full GPU suite + bench (N=1) + micro-benches after the rpn-head / PS-ROI / backward changes
tag=${1:-run}; out=gpurun_out/$tag; mkdir -p $out; export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -q -rA --timeout 300 --timeout-method thread > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|FAILED" $out/pytest_gpu.log | tail -20
timeout 900 python bench.py > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"; tail -5 $out/bench.err
python - <<PY
import json
d=json.load(open('$out/bench.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], 'hot', d['hot_path'])
print('roofline', d['roofline']['frac'], d['roofline']['duration_us'])
print('train', d['train'])
print('configs', json.dumps(d['configs'])[:1500])
for r in d['sweep']: print(r['N'], r['d'], r['H'], r.get('nm_us'), r.get('module_us'), r['path'][:30])
PY
timeout 300 python tools/backward_bench.py 2>/dev/null | tee $out/backward_bench.jsonl
timeout 200 python tools/psroi_bench.py 2>/dev/null | tee $out/psroi_bench.jsonl
timeout 200 python tools/timeline.py > $out/timeline.log 2>&1; cp gpurun_out/timeline.json $out/timeline.json; tail -1 $out/timeline.log
