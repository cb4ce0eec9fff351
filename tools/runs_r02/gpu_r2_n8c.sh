#!/bin/bash
# Round 2, third 8-GPU round trip (short): C4 strong at N=8 twice WITHOUT poison (first/last-step checksums must equal the
# one-GPU run's 7bfe1f4ebd35bd47), C2 weak N=8 for the final line.
mkdir -p gpurun_out
T=gpurun_out/r2n8c
run() {
  local name=$1 np=$2; shift 2
  SMJ_DIST_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $np "$@" > ${T}_${name}.json 2> ${T}_${name}.err
  echo "bench $name exit $?"; grep '^{' ${T}_${name}.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); c=d['config']; print(round(d['ms_per_step'],4), round(d['value']), c['rows_joined'], c.get('result_checksum'), c.get('result_checksum_after_timed_steps'), c.get('rows_joined_first_and_last_step'), d['e2e'] and round(d['e2e']['ms_per_step'],2))"
}
run c4_n8_a 8 --workload c4 --scaling strong --steps 5 --warmup 3 --no-e2e
run c4_n8_b 8 --workload c4 --scaling strong --steps 5 --warmup 3 --no-e2e
run c2_n8 8 --steps 20 --warmup 5
