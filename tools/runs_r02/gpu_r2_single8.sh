#!/bin/bash
# Round 2, eighth 1-GPU round trip: parity (graph replay on new tables, the 8-column case on 8 ranks), a C4-like shape on
# 8 ranks on one GPU against the oracle, C2 bench with the fresh-table and eager timings.
mkdir -p gpurun_out
T=gpurun_out/r2aa
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -p no:cacheprovider --deselect tests/test_gpu_full_size.py -k "not (end_to_end and (2] or 4]))" > ${T}_tests.log 2>&1
echo "pytest exit $?" | tee -a ${T}_tests.log; tail -4 ${T}_tests.log | cut -c1-300
timeout 400 python tools/dist_onegpu.py 8 4000000 2 1 8 0.1 800000 > ${T}_onegpu_c4like.txt 2>&1; echo "onegpu c4-like check exit $?"; tail -2 ${T}_onegpu_c4like.txt | cut -c1-300
timeout 400 python tools/dist_onegpu.py 4 4000000 2 1 8 0.1 800000 > ${T}_onegpu_c4like_g4.txt 2>&1; echo "onegpu c4-like G=4 check exit $?"; tail -1 ${T}_onegpu_c4like_g4.txt | cut -c1-300
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > ${T}_bench_c2.json 2> ${T}_bench_c2.err; echo "bench c2 exit $?"; python -c "import json; d=json.loads(open('${T}_bench_c2.json').read()); print(round(d['ms_per_step'],4), 'fresh', d['fresh_tables_ms_per_step'], d['fresh_tables_graph_replayed'], 'eager', d['eager_ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['config']['result_checksum'])"
