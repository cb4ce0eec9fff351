#!/bin/bash
# GPU round trip: A/B bench lines over the launch knobs.
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e"
for v in "SMJ_PDL=1" "SMJ_PDL=0" "SMJ_PDL=1 SMJ_STAGE_EVENTS=0" "SMJ_PDL=0 SMJ_STAGE_EVENTS=0" "SMJ_PDL=1 SMJ_NO_GRAPH=1" "SMJ_PDL=0 SMJ_NO_GRAPH=1"; do
  echo "== $v"
  env $v timeout 300 $B 2> gpurun_out/ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print(d['ms_per_step'], d['value'], d['stage_ms'], d['roofline']['avg_launch_ms'], d['config']['rows_joined'])"
  tail -2 gpurun_out/ab.err
done 2>&1 | tee gpurun_out/ab_pdl.txt
