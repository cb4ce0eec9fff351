#!/bin/bash
# Round 2, second 2-GPU round trip (short): one-process mode + fabric parity at world 2, N=2 bench lines (C2 weak, C4 strong).
mkdir -p gpurun_out
T=gpurun_out/r2n2b
timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_host_csv.py -m gpu -v --timeout 400 -p no:cacheprovider -k "(multi_gpu and (2-peer or 2-overflow or 2-default or 2-nccl or 2-merge)) or (end_to_end and (2] or g2-1]))" > ${T}_tests.log 2>&1
echo "multi pytest exit $?" | tee -a ${T}_tests.log; grep -E "PASSED|FAILED|SKIPPED|passed|failed|Error" ${T}_tests.log | tail -30 | cut -c1-220
run() {
  local name=$1 np=$2; shift 2
  SMJ_DIST_TRACE=1 timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $np "$@" > ${T}_${name}.json 2> ${T}_${name}.err
  echo "bench $name exit $?"; grep '^{' ${T}_${name}.json | cut -c1-330; grep "\[dist\]" ${T}_${name}.err | tail -2
}
run bench_c2_n2 2 --steps 20 --warmup 5 --no-e2e
SMJ_DIST_STREAMS=1 run bench_c2_n2_s1 2 --steps 20 --warmup 5 --no-e2e
run bench_c4_n2 2 --workload c4 --scaling strong --steps 5 --warmup 3 --no-e2e
