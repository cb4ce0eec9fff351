#!/bin/bash
# Round 2: parity run for the int64-layout adapters (smj_table_from_i64 / to_i64) and the many-CTA scan of the stage entry
# points; whole -m gpu suite with -x as the driver runs it.
mkdir -p gpurun_out
T=gpurun_out/r2i
timeout 700 python -m pytest tests -m gpu -x -q --timeout 600 -p no:cacheprovider --durations=5 > ${T}_tests.log 2>&1
echo "pytest exit $?" | tee -a ${T}_tests.log; tail -12 ${T}_tests.log | cut -c1-300
