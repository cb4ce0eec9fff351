#!/bin/bash
# Round 2, sixth 1-GPU round trip: parity (incl. all-ranks-on-one-GPU variants) with the two-pass partition, G = 8 on one GPU
# checked against the oracle (staged and direct stores), launch list of the world-1 fabric path.
mkdir -p gpurun_out
T=gpurun_out/r2y
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -p no:cacheprovider --deselect tests/test_gpu_full_size.py -k "not (end_to_end and (2] or 4]))" > ${T}_tests.log 2>&1
echo "pytest exit $?" | tee -a ${T}_tests.log; tail -4 ${T}_tests.log | cut -c1-300
timeout 300 python tools/dist_onegpu.py 8 2000000 2 1 > ${T}_onegpu_check.txt 2>&1; echo "onegpu check exit $?"; tail -2 ${T}_onegpu_check.txt | cut -c1-400
SMJ_DIST_STAGE_MIN_G=99 timeout 300 python tools/dist_onegpu.py 8 2000000 2 1 > ${T}_onegpu_check_direct.txt 2>&1; echo "onegpu check (direct stores) exit $?"; tail -1 ${T}_onegpu_check_direct.txt | cut -c1-200
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file ${T}_launches_dist1.csv python tools/dist1.py 4 > ${T}_ncu_dist1.log 2>&1
echo "ncu dist1 exit $?"; python tools/ncu_summary.py step_bytes ${T}_launches_dist1.csv 4 | head -9
timeout 300 ncu --set full --clock-control none --import-source on -k regex:partition_route -s 2 -c 1 -f -o ${T}_prof_partition_pass python tools/dist1.py 3 > ${T}_ncu_partition_pass.log 2>&1
echo "ncu partition_pass exit $?"
