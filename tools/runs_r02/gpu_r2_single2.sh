#!/bin/bash
# Round 2, second 1-GPU round trip: parity suite, row-store A/B at C2, merge lab, full ncu of the partition / exchange
# kernels (world-1 fabric path), full-size tests (c3 Zipf, c4 pinned to the C port), the C1 leg.
mkdir -p gpurun_out
T=gpurun_out/r2t
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -p no:cacheprovider --ignore=tests/test_multi_gpu.py --deselect tests/test_gpu_full_size.py -k "not (end_to_end and (2] or 4]))" > ${T}_tests.log 2>&1
echo "pytest exit $?" | tee -a ${T}_tests.log; tail -4 ${T}_tests.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > ${T}_bench_c2_rowstore.json 2> ${T}_bench_c2_rowstore.err; echo "bench c2 rowstore exit $?"; cut -c1-260 ${T}_bench_c2_rowstore.json
SMJ_ROWSTORE_MB=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > ${T}_bench_c2_norowstore.json 2> ${T}_bench_c2_norowstore.err; echo "bench c2 no rowstore exit $?"; cut -c1-260 ${T}_bench_c2_norowstore.json
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
SMJ_BENCH_NO_EAGER=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file ${T}_launches_c2.csv $CMD > ${T}_ncu_c2.log 2>&1
echo "ncu c2 exit $?"; python tools/ncu_summary.py step_bytes ${T}_launches_c2.csv 4 ${T}_c2_dram_bytes.json "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none; $CMD (4 identical steps)"
timeout 200 tools/bin/merge_lab 100000000 > ${T}_merge_lab.txt 2>&1; echo "merge_lab exit $?"; cat ${T}_merge_lab.txt
timeout 200 tools/bin/merge_lab 5000000 >> ${T}_merge_lab.txt 2>&1; tail -2 ${T}_merge_lab.txt
for spec in select_partition:2 partition_exchange:2 rowstore:1 join_materialize:1; do
  k=${spec%%:*}; s=${spec##*:}
  if [ $k = rowstore ] || [ $k = join_materialize ]; then PCMD="$CMD"; export SMJ_BENCH_NO_EAGER=1; else PCMD="python tools/dist1.py 3"; fi
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c 1 -f -o ${T}_prof_$k $PCMD > ${T}_ncu_$k.log 2>&1
  echo "ncu $k exit $?"
done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file ${T}_launches_dist1.csv python tools/dist1.py 4 > ${T}_ncu_dist1.log 2>&1
echo "ncu dist1 exit $?"; python tools/ncu_summary.py step_bytes ${T}_launches_dist1.csv 4
timeout 1500 python -m pytest tests/test_gpu_full_size.py -m gpu -v --timeout 1400 -p no:cacheprovider --durations=5 > ${T}_tests_full.log 2>&1
echo "full-size pytest exit $?" | tee -a ${T}_tests_full.log; grep -E "PASSED|FAILED|passed|failed|s call" ${T}_tests_full.log | tail -12
timeout 600 python tools/bench_c1.py > ${T}_bench_c1.json 2> ${T}_bench_c1.err; echo "bench c1 exit $?"; cut -c1-1200 ${T}_bench_c1.json; tail -3 ${T}_bench_c1.err
