#!/bin/bash
# Round 2 final parity run on one B200: the whole -m gpu suite (no -x), then the smoke entry and a C2 bench line.
mkdir -p gpurun_out
T=gpurun_out/r2g
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider --durations=8 > ${T}_tests.log 2>&1
echo "pytest exit $?" | tee -a ${T}_tests.log; tail -14 ${T}_tests.log | cut -c1-200
timeout 300 python __graft_entry__.py smoke > ${T}_smoke.log 2>&1; echo "smoke exit $?" >> ${T}_smoke.log; tail -2 ${T}_smoke.log
timeout 600 python bench.py > ${T}_bench_n1.json 2> ${T}_bench_n1.err; echo "bench exit $?"; python -c "import json; d=json.loads(open('${T}_bench_n1.json').read()); r=d['roofline']; print(round(d['ms_per_step'],4), d['fresh_tables_ms_per_step'], d['eager_ms_per_step'], d['e2e']['ms_per_step'], r['frac'], r['traffic'], r['pipeline_dram_frac_of_peak'], d['gpu_launches'])"
