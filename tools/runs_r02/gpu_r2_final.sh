#!/bin/bash
# Round 2 final 1-GPU capture: full parity suite, smoke, the bench lines (own arm, reference arm, c3 / c4), launch list with
# DRAM bytes, full ncu captures of the top kernels.
mkdir -p gpurun_out
T=gpurun_out/r2f
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > ${T}_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider --durations=8 > ${T}_tests.log 2>&1
echo "pytest exit $?" | tee -a ${T}_tests.log; tail -14 ${T}_tests.log | cut -c1-200
timeout 300 python __graft_entry__.py smoke > ${T}_smoke.log 2>&1; echo "smoke exit $?" >> ${T}_smoke.log; tail -2 ${T}_smoke.log
timeout 600 python bench.py > ${T}_bench_n1.json 2> ${T}_bench_n1.err; echo "bench exit $?"; cut -c1-500 ${T}_bench_n1.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > ${T}_bench_ref.json 2> ${T}_bench_ref.err; echo "ref exit $?"; cut -c1-300 ${T}_bench_ref.json
for w in c4 c3; do
  timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > ${T}_bench_$w.json 2> ${T}_bench_$w.err
  echo "bench $w exit $?"; cut -c1-300 ${T}_bench_$w.json
done
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
SMJ_BENCH_NO_EAGER=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file ${T}_launches_c2.csv $CMD > ${T}_ncu_c2.log 2>&1
echo "ncu c2 exit $?"; python tools/ncu_summary.py step_bytes ${T}_launches_c2.csv 5 ${T}_c2_dram_bytes.json "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none; $CMD (5 identical steps incl. the checksum step)" > ${T}_launches_c2_summary.txt 2>&1; tail -3 ${T}_launches_c2_summary.txt
export SMJ_BENCH_NO_EAGER=1
for spec in radix_pass:8 select_tma:2 join_materialize:1; do
  k=${spec%%:*}; s=${spec##*:}
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c 2 -f -o ${T}_prof_$k $CMD > ${T}_ncu_$k.log 2>&1
  echo "ncu $k exit $?"
done
