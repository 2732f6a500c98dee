#!/bin/bash
# One gpurun batch: GPU tests, the bench line, the ncu launch list of one step, ncu --set full of the fused relation
# kernel at N=300 and N=3000.  Everything lands in gpurun_out/ (copied to profiles/ by hand after reading).
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh <tag>'
tag=${1:-run}
out=gpurun_out/$tag
mkdir -p $out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 500 > $out/clocks.csv &
SMI=$!
if [ -z "$SKIP_TESTS" ]; then
  timeout 900 python -m pytest tests -m gpu -q -rA > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/pytest_gpu.log
  tail -3 $out/pytest_gpu.log
fi
if [ -z "$SKIP_BENCH" ]; then
  timeout 1200 python bench.py > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
  head -c 1500 $out/bench.json; echo
fi
if [ -z "$SKIP_NCU" ]; then
  timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
      --log-file $out/launches_step.csv python tools/profile_step.py --what step > $out/launches_step.log 2>&1
  for n in 300 3000; do
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:relation_fused -s 2 -c 1 -f \
        -o $out/ncu_fused_n$n python tools/fused_one.py $n 1024 16 1 4 > $out/ncu_fused_n$n.log 2>&1
    echo "ncu n=$n rc=$?"
  done
fi
kill $SMI
ls -la $out
