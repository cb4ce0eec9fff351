#!/bin/bash
# Round 2, third 2-GPU round trip (short): C4 strong at N=2 with the result checksum beside N=1's, C2 weak N=2 with the trace.
mkdir -p gpurun_out
T=gpurun_out/r2n2c
run() {
  local name=$1 np=$2; shift 2
  SMJ_DIST_TRACE=1 timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $np "$@" > ${T}_${name}.json 2> ${T}_${name}.err
  echo "bench $name exit $?"; grep '^{' ${T}_${name}.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), d['stage_ms'], d['config']['rows_joined'], d['config'].get('result_checksum'), d['roofline'].get('nvlink_gbs_per_gpu'))"; grep "\[dist\]" ${T}_${name}.err | tail -1
}
run bench_c2_n2 2 --steps 20 --warmup 5 --no-e2e
SMJ_DIST_STREAMS=1 run bench_c2_n2_s1 2 --steps 20 --warmup 5 --no-e2e
SMJ_PT_CTAS=3 run bench_c2_n2_pt3 2 --steps 20 --warmup 5 --no-e2e
run bench_c4_n2 2 --workload c4 --scaling strong --steps 5 --warmup 3 --no-e2e
timeout 300 python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > ${T}_bench_c4_n1.json 2> ${T}_bench_c4_n1.err
echo "bench c4 n1 exit $?"; python -c "import sys,json; d=json.loads(open('${T}_bench_c4_n1.json').read()); print(round(d['ms_per_step'],4), d['config']['rows_joined'], d['config'].get('result_checksum'))"
