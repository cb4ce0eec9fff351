#!/bin/bash
# A/B bench lines over environment knobs, one short GPU round trip:
#   WORKLOADS="c2 c4" tools/gpu_ab.sh "SMJ_SEMIJOIN=1" "SMJ_SEMIJOIN=0" "SMJ_PDL=0 SMJ_NO_GRAPH=1"
# (SMJ_LIB=path/to/variant/libsmj.so in a variant compares kernel builds, e.g. -DSMJ_SEL_CTAS=3)
mkdir -p gpurun_out
for w in ${WORKLOADS:-c2}; do
for v in "$@"; do
  echo "== $w $v"
  env $v timeout 300 python bench.py --workload $w --steps ${STEPS:-10} --warmup 3 --no-cpu-baseline --no-e2e 2> gpurun_out/ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); r=d['roofline']
print(d['ms_per_step'], d['value'], d['stage_ms'], r['avg_launch_ms'], round(r['frac'], 4), d['config']['rows_selected'], d['config']['rows_joined'])"
  tail -2 gpurun_out/ab.err
done; done 2>&1 | tee gpurun_out/ab.txt
