#!/bin/bash
# Round 2, second 8-GPU round trip (diagnostic, short): C4 strong at N=8 with poisoned receive buffers and first/last-step
# checksums (fabric and NCCL paths), C2 weak with poison.
mkdir -p gpurun_out
T=gpurun_out/r2n8b
run() {
  local name=$1 np=$2; shift 2
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $np "$@" > ${T}_${name}.json 2> ${T}_${name}.err
  echo "bench $name exit $?"; grep '^{' ${T}_${name}.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); c=d['config']; print(round(d['ms_per_step'],4), c['rows_joined'], c.get('result_checksum'), c.get('result_checksum_after_timed_steps'), c.get('rows_joined_first_and_last_step'))"
}
SMJ_DIST_POISON=1 run c4_n8_poison 8 --workload c4 --scaling strong --steps 3 --warmup 3 --no-e2e
SMJ_DIST_EXCHANGE=nccl run c4_n8_nccl 8 --workload c4 --scaling strong --steps 3 --warmup 3 --no-e2e
SMJ_DIST_POISON=1 run c2_n8_poison 8 --steps 5 --warmup 3 --no-e2e
