#!/bin/bash
# GPU round trip: parity tests, the default bench line, the c3 / c4 workloads.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/tests.log
tail -4 gpurun_out/tests.log
for v in "SMJ_PDL=1" "SMJ_STAGE_EVENTS=0" "SMJ_NO_GRAPH=1"; do
  echo "== $v"
  env $v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e 2> gpurun_out/ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print(d['ms_per_step'], d['value'], d['stage_ms'], d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['config']['rows_joined'])"
  tail -2 gpurun_out/ab.err
done 2>&1 | tee gpurun_out/ab2.txt
for w in c4 c3; do
  timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  echo "bench $w exit $?"; cat gpurun_out/bench_$w.json; tail -3 gpurun_out/bench_$w.err
done
