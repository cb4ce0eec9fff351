#!/bin/bash
# 2-GPU round trip: the multi-GPU parity tests on the default (peer-store) path, then the N=2 bench line.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_multi_gpu.py -m gpu -q -x --timeout 900 -p no:cacheprovider -k "2-peer or 2-merge" > gpurun_out/tests_n2.log 2>&1
echo "pytest exit $?" >> gpurun_out/tests_n2.log
tail -4 gpurun_out/tests_n2.log
SMJ_DIST_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "bench n=2 exit $?"; cut -c1-300 gpurun_out/bench_n2.json; grep "\[dist\]" gpurun_out/bench_n2.err | tail -1
