#!/bin/bash
# GPU round trip: parity tests, then A/B bench lines (device sort plans vs forced four passes) and the CUB comparator.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/tests.log
tail -8 gpurun_out/tests.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_plan.json 2> gpurun_out/bench_plan.err
echo "bench(plan) exit $?"; cat gpurun_out/bench_plan.json; tail -3 gpurun_out/bench_plan.err
SMJ_FULL_PASSES=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
echo "bench(full passes) exit $?"; cat gpurun_out/bench_full.json; tail -3 gpurun_out/bench_full.err
timeout 300 tools/bin/cub_compare > gpurun_out/cub_compare.txt 2>&1; echo "cub exit $?"; cat gpurun_out/cub_compare.txt
for a in "5000000 25" "10000000 24" "50000000 32" "200000000 32"; do timeout 120 tools/bin/radix_lab $a; done > gpurun_out/radix_lab_r1b.txt 2>&1
cat gpurun_out/radix_lab_r1b.txt
