#!/bin/bash
# Round 2, 1-GPU round trip: parity suite, C2 bench with A/B knobs, launch lists with DRAM bytes (single-GPU step and the
# world-1 fabric path), c3 (Zipf) / c4 bench lines.
mkdir -p gpurun_out
T=gpurun_out/r2s
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -p no:cacheprovider --ignore=tests/test_multi_gpu.py --deselect tests/test_gpu_full_size.py -k "not (end_to_end and (2 or 4))" > ${T}_tests.log 2>&1
echo "pytest exit $?" | tee -a ${T}_tests.log; tail -4 ${T}_tests.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > ${T}_bench_c2.json 2> ${T}_bench_c2.err; echo "bench c2 exit $?"; cut -c1-400 ${T}_bench_c2.json
SMJ_RADIX_DYN_TILES=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > ${T}_bench_c2_fixedtiles.json 2> ${T}_bench_c2_fixedtiles.err; echo "bench c2 fixed tiles exit $?"; cut -c1-260 ${T}_bench_c2_fixedtiles.json
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
SMJ_BENCH_NO_EAGER=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file ${T}_launches_c2.csv $CMD > ${T}_ncu_c2.log 2>&1
echo "ncu c2 exit $?"; python tools/ncu_summary.py step_bytes ${T}_launches_c2.csv 4 ${T}_c2_dram_bytes.json "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none; $CMD (4 identical steps)"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file ${T}_launches_dist1.csv python tools/dist1.py 4 > ${T}_ncu_dist1.log 2>&1
echo "ncu dist1 exit $?"; python tools/ncu_summary.py step_bytes ${T}_launches_dist1.csv 4
for w in c3 c4; do
  timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > ${T}_bench_$w.json 2> ${T}_bench_$w.err
  echo "bench $w exit $?"; cut -c1-330 ${T}_bench_$w.json; tail -2 ${T}_bench_$w.err
done
