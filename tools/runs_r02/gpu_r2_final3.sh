#!/bin/bash
# Round 2, parity run after the merge fix (unsorted runs: skipped tiles are zero-filled and reported): whole -m gpu suite,
# smoke entry, one C2 bench line.
mkdir -p gpurun_out
T=gpurun_out/r2h
timeout 700 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider --durations=8 > ${T}_tests.log 2>&1
echo "pytest exit $?" | tee -a ${T}_tests.log; tail -14 ${T}_tests.log | cut -c1-200
timeout 200 python __graft_entry__.py smoke > ${T}_smoke.log 2>&1; echo "smoke exit $?" >> ${T}_smoke.log; tail -2 ${T}_smoke.log
timeout 300 python bench.py > ${T}_bench_n1.json 2> ${T}_bench_n1.err; echo "bench exit $?"; python -c "import json; d=json.loads(open('${T}_bench_n1.json').read()); r=d['roofline']; print(round(d['ms_per_step'],4), d['fresh_tables_ms_per_step'], d['eager_ms_per_step'], d['e2e']['ms_per_step'], r['frac'], r['traffic'], r['pipeline_dram_frac_of_peak'], d['gpu_launches'])"
