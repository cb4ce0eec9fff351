#!/bin/bash
# GPU round trip used while iterating: parity tests, then a short bench line.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 -p no:cacheprovider > gpurun_out/tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/tests.log
tail -15 gpurun_out/tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
echo "bench exit $?"; cat gpurun_out/bench_quick.json; tail -5 gpurun_out/bench_quick.err
