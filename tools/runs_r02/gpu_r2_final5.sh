#!/bin/bash
# Round 2, last 1-GPU capture: the whole -m gpu suite as the driver runs it (-x), the smoke entry, the default bench line, and
# C1 through the C driver beside the reference's own cpu_app at -O2 (the -O0 build, 34 s, was timed earlier in the round).
mkdir -p gpurun_out
T=gpurun_out/r2k
timeout 700 python -m pytest tests -m gpu -x -q --timeout 600 -p no:cacheprovider --durations=5 > ${T}_tests.log 2>&1
echo "pytest exit $?" | tee -a ${T}_tests.log; tail -5 ${T}_tests.log | cut -c1-200
timeout 200 python __graft_entry__.py smoke > ${T}_smoke.log 2>&1; echo "smoke exit $?" >> ${T}_smoke.log; tail -2 ${T}_smoke.log
timeout 300 python bench.py > ${T}_bench_n1.json 2> ${T}_bench_n1.err; echo "bench exit $?"; python -c "import json; d=json.loads(open('${T}_bench_n1.json').read()); r=d['roofline']; print(round(d['ms_per_step'],4), d['fresh_tables_ms_per_step'], d['eager_ms_per_step'], d['e2e']['ms_per_step'], r['frac'], r['traffic'], r['pipeline_dram_frac_of_peak'], d['gpu_launches'])"
timeout 300 python tools/bench_c1.py --no-O0 --csv-rows 0 > ${T}_bench_c1.json 2> ${T}_bench_c1.err; echo "c1 exit $?"; python -c "import json; d=json.loads(open('${T}_bench_c1.json').read()); print(d['ms_per_run'], d['ours'], d['vs_reference']['ratio_vs_O2'])"
