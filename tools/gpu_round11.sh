#!/bin/bash
tag=${1:-run}; out=gpurun_out/$tag; mkdir -p $out; export PYTHONUNBUFFERED=1
timeout 900 python bench.py --no-sweep > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"; tail -5 $out/bench.err
python - <<PY
import json
d=json.load(open('$out/bench.json'))
print('value', d['value'], 'e2e', d['e2e']['value'])
print('train', json.dumps(d['train']))
print('cfg3 train', json.dumps(d['configs']['3_fpn']['train']))
print('cfg2', d['configs']['2_deformable_faster']['images_per_sec'], 'cfg3 test', d['configs']['3_fpn']['test'])
PY
