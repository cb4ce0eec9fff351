#!/bin/bash
# Round 2, fifth 1-GPU round trip (short): launch list of the world-1 fabric path (partition v3 / exchange kernels), full
# captures of both, radix lab at 512 and 256 threads per CTA, C2 / c3 with the 256-thread variant.
mkdir -p gpurun_out
T=gpurun_out/r2w
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file ${T}_launches_dist1.csv python tools/dist1.py 4 > ${T}_ncu_dist1.log 2>&1
echo "ncu dist1 exit $?"; python tools/ncu_summary.py step_bytes ${T}_launches_dist1.csv 4
for spec in select_partition:2 partition_exchange:2; do
  k=${spec%%:*}; s=${spec##*:}
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c 1 -f -o ${T}_prof_$k python tools/dist1.py 3 > ${T}_ncu_$k.log 2>&1
  echo "ncu $k exit $?"
done
for n in 5000000 50000000 400000000; do
  for v in radix_lab radix_lab_256; do echo "== $v $n"; timeout 200 tools/bin/$v $n 31 2>&1 | tail -3; done
done > ${T}_radix_lab.txt 2>&1; cat ${T}_radix_lab.txt
for w in c2 c3; do
  SMJ_LIB=$PWD/tools/bin/libsmj_rs256.so timeout 400 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > ${T}_bench_${w}_rs256.json 2> ${T}_bench_${w}_rs256.err
  echo "bench $w rs256 exit $?"; python -c "import json; d=json.loads(open('${T}_bench_${w}_rs256.json').read()); print(round(d['ms_per_step'],4), d['stage_ms'], round(d['roofline']['frac'],3), d['config']['rows_joined'])"
done
