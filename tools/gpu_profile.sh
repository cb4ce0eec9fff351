#!/bin/bash
# Full ncu captures of the hot kernels of one bench step (after the same command exited 0 without ncu).
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 300 $CMD > gpurun_out/prof_plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/prof_plain.log; exit 1; }
cat gpurun_out/prof_plain.log
for k in radix_pass select_pairs join_match join_materialize join_partition; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 8 -c 2 -f -o gpurun_out/prof_$k $CMD > gpurun_out/ncu_$k.log 2>&1
  echo "ncu $k exit $?"
done
ls -la gpurun_out
