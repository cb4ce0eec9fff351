#!/bin/bash
# Round 2, fourth 1-GPU round trip: parity incl. the all-ranks-on-one-GPU variants, the fabric path at G = 8 on one GPU
# (checked against the oracle, then a launch list), merge lab, C2 bench.
mkdir -p gpurun_out
T=gpurun_out/r2v
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -p no:cacheprovider --deselect tests/test_gpu_full_size.py -k "not (end_to_end and (2] or 4]))" > ${T}_tests.log 2>&1
echo "pytest exit $?" | tee -a ${T}_tests.log; tail -4 ${T}_tests.log | cut -c1-300
timeout 300 python tools/dist_onegpu.py 8 2000000 2 1 > ${T}_onegpu_check.txt 2>&1; echo "onegpu check exit $?"; tail -2 ${T}_onegpu_check.txt | cut -c1-400
SMJ_DIST_STAGE_MIN_G=99 timeout 300 python tools/dist_onegpu.py 8 2000000 2 1 > ${T}_onegpu_check_direct.txt 2>&1; echo "onegpu check (direct stores) exit $?"; tail -1 ${T}_onegpu_check_direct.txt | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1500 --csv --log-file ${T}_launches_onegpu8.csv python tools/dist_onegpu.py 8 10000000 2 > ${T}_ncu_onegpu8.log 2>&1
echo "ncu onegpu8 exit $?"; tail -2 ${T}_ncu_onegpu8.log | cut -c1-300; python tools/ncu_summary.py step_bytes ${T}_launches_onegpu8.csv 16
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > ${T}_bench_c2.json 2> ${T}_bench_c2.err; echo "bench c2 exit $?"; cut -c1-260 ${T}_bench_c2.json
timeout 200 tools/bin/merge_lab 100000000 > ${T}_merge_lab.txt 2>&1; echo "merge_lab exit $?"; cat ${T}_merge_lab.txt
