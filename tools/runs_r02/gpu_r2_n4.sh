#!/bin/bash
# Round 2, 4-GPU round trip: the multi-GPU parity tests at world 2 and 4 (fabric / overflow / nccl / merge paths, torchrun and
# one-process), the C driver with SMJ_NR_GPUS = 2 and 4, then bench lines: C2 weak N=4, C4 strong N=4 and N=2.
mkdir -p gpurun_out
T=gpurun_out/r2n4
nvidia-smi --query-gpu=index,name,memory.total --format=csv > ${T}_gpu.txt 2>&1
timeout 1500 python -m pytest tests/test_multi_gpu.py tests/test_host_csv.py -m gpu -v --timeout 600 -p no:cacheprovider -k "${MULTI_K:-multi_gpu or end_to_end}" > ${T}_tests_multi.log 2>&1
echo "multi pytest exit $?" | tee -a ${T}_tests_multi.log; grep -E "PASSED|FAILED|SKIPPED|passed|failed" ${T}_tests_multi.log | tail -40 | cut -c1-200
run() { # name nproc args...
  local name=$1 np=$2; shift 2
  SMJ_DIST_TRACE=1 timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $np "$@" > ${T}_${name}.json 2> ${T}_${name}.err
  echo "bench $name exit $?"; grep '^{' ${T}_${name}.json | cut -c1-330; grep "\[dist\]" ${T}_${name}.err | tail -1
}
run bench_c2_n4 4 --steps 20 --warmup 5
run bench_c2_n2 2 --steps 20 --warmup 5
run bench_c4_n4 4 --workload c4 --scaling strong --steps 5 --warmup 3 --no-e2e
run bench_c4_n2 2 --workload c4 --scaling strong --steps 5 --warmup 3 --no-e2e
