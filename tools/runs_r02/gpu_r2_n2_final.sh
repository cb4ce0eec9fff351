#!/bin/bash
# Round 2, last 2-GPU round trip: the multi-GPU tests and the C driver with 1 / 2 GPUs on the round's last build, the driver's
# own N=2 bench command, and C1 through the C driver after the allocation warm-up.
mkdir -p gpurun_out
T=gpurun_out/r2m
timeout 500 python -m pytest tests/test_multi_gpu.py tests/test_host_csv.py -m gpu -x -q --timeout 400 -p no:cacheprovider > ${T}_tests.log 2>&1
echo "pytest exit $?" | tee -a ${T}_tests.log; tail -3 ${T}_tests.log | cut -c1-200
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 10 --warmup 3 > ${T}_bench_c2_n2.json 2> ${T}_bench_c2_n2.err
echo "bench n2 exit $?"; grep '^{' ${T}_bench_c2_n2.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), d['stage_ms'], d['config']['rows_joined'], d['config'].get('result_checksum'), d['config'].get('result_checksum_after_timed_steps'), d['e2e']['ms_per_step'])"
timeout 300 python tools/bench_c1.py --no-O0 --csv-rows 0 > ${T}_bench_c1.json 2> ${T}_bench_c1.err; echo "c1 exit $?"; python -c "import json; d=json.loads(open('${T}_bench_c1.json').read()); print(d['ms_per_run'], d['ours'], d['vs_reference']['ratio_vs_O2'])"
