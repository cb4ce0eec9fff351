#!/bin/bash
# Launch list + full ncu captures of the hot kernels of one bench step (after the same command exited 0 without ncu).
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e ${BENCH_ARGS}"
timeout 300 $CMD > gpurun_out/prof_plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/prof_plain.log; exit 1; }
cat gpurun_out/prof_plain.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?"
for spec in ${KERNELS:-radix_pass:8 select_tma:2 join_match:1 join_materialize:1 join_partition:1}; do
  k=${spec%%:*}; s=${spec##*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c 2 -f -o gpurun_out/prof_$k $CMD > gpurun_out/ncu_$k.log 2>&1
  echo "ncu $k exit $?"
done
ls -la gpurun_out
