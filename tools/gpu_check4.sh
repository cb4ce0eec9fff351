#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/tests.log; tail -3 gpurun_out/tests.log
for w in c2 c4; do
  echo "== $w"
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2> gpurun_out/ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print(d['ms_per_step'], d['value'], d['stage_ms'], d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['config']['rows_joined'])"
  tail -2 gpurun_out/ab.err
done 2>&1 | tee gpurun_out/ab4.txt
KERNELS="select_tma:2 bloom_filter:1 plan_compact:1" bash tools/gpu_profile.sh > gpurun_out/profile_run.log 2>&1; tail -2 gpurun_out/profile_run.log
