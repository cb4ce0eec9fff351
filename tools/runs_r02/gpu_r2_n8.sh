#!/bin/bash
# Round 2, 8-GPU round trip (charged 8x: keep it short): parity at world 8, C2 weak N=8 (fabric / NCCL / merge paths),
# C4 strong N=8, C5 (2B x 2B) N=8.
mkdir -p gpurun_out
T=gpurun_out/r2n8
nvidia-smi --query-gpu=index,name,memory.total --format=csv > ${T}_gpu.txt 2>&1
timeout 500 python -m pytest tests/test_multi_gpu.py -m gpu -v --timeout 400 -p no:cacheprovider -k "8-peer or 8-default" > ${T}_tests_multi.log 2>&1
echo "multi pytest exit $?" | tee -a ${T}_tests_multi.log; grep -E "PASSED|FAILED|SKIPPED|passed|failed" ${T}_tests_multi.log | tail -6 | cut -c1-200
run() {
  local name=$1 np=$2; shift 2
  SMJ_DIST_TRACE=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $np "$@" > ${T}_${name}.json 2> ${T}_${name}.err
  echo "bench $name exit $?"; grep '^{' ${T}_${name}.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), round(d['value']), d['stage_ms'], d['config']['rows_joined'], d['config'].get('result_checksum'), d['roofline'].get('nvlink_gbs_per_gpu'), d['e2e'] and round(d['e2e']['ms_per_step'],2))"; grep "\[dist\]" ${T}_${name}.err | tail -1
}
run bench_c2_n8 8 --steps 20 --warmup 5
run bench_c4_n8 8 --workload c4 --scaling strong --steps 5 --warmup 3 --no-e2e
run bench_c5_n8 8 --workload c5 --scaling strong --steps 3 --warmup 3 --no-e2e
SMJ_DIST_EXCHANGE=nccl run bench_c2_n8_nccl 8 --steps 10 --warmup 3 --no-e2e
SMJ_DIST_MODE=merge run bench_c2_n8_merge 8 --steps 10 --warmup 3 --no-e2e
SMJ_DIST_STREAMS=1 run bench_c2_n8_s1 8 --steps 10 --warmup 3 --no-e2e
