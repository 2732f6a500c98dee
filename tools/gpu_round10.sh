#!/bin/bash
tag=${1:-run}; out=gpurun_out/$tag; mkdir -p $out; export PYTHONUNBUFFERED=1
timeout 200 python tools/deform_bench.py 2>/dev/null | tee $out/deform_bench.jsonl
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $out/launches_deform.csv python tools/deform_bench.py > $out/deform.log 2>&1
python - <<PY
import csv
rows=list(csv.reader(open('$out/launches_deform.csv')))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: h=r; start=i; break
ix={k:i for i,k in enumerate(h)}
for r in rows[start+2:][-8:]:
    if len(r)>=len(h): print(r[ix['Kernel Name']].split('(')[0][-50:], float(r[ix['Metric Value']])/1000, r[ix['Grid Size']])
PY
