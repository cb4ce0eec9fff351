#!/bin/bash
# Round 2, 2-GPU round trip: single-GPU regression, the fabric path's parity tests at world 2, the N=1/N=2 bench lines.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/r2_gpu.txt 2>&1
timeout 600 python -m pytest tests -m gpu -q -x --timeout 300 -p no:cacheprovider --ignore=tests/test_multi_gpu.py -k "not nr_gpus or 1-" > gpurun_out/r2_tests_single.log 2>&1
echo "single pytest exit $?" | tee -a gpurun_out/r2_tests_single.log; tail -3 gpurun_out/r2_tests_single.log
timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_host_csv.py -m gpu -q --timeout 600 -p no:cacheprovider -k "${MULTI_K:-2-peer or 2-overflow or 2-default or (end_to_end and 2-)}" > gpurun_out/r2_tests_n2.log 2>&1
echo "multi pytest exit $?" | tee -a gpurun_out/r2_tests_n2.log; tail -30 gpurun_out/r2_tests_n2.log | cut -c1-400
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
echo "bench n=1 exit $?"; cut -c1-260 gpurun_out/r2_bench_n1.json
for s in 2 1; do
SMJ_DIST_STREAMS=$s SMJ_DIST_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$s bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2_s$s.json 2> gpurun_out/r2_bench_n2_s$s.err
echo "bench n=2 streams=$s exit $?"; cut -c1-300 gpurun_out/r2_bench_n2_s$s.json; grep "\[dist\]" gpurun_out/r2_bench_n2_s$s.err | tail -2
done
SMJ_DIST_EXCHANGE=nccl timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e > gpurun_out/r2_bench_n2_nccl.json 2> gpurun_out/r2_bench_n2_nccl.err
echo "bench n=2 nccl exit $?"; cut -c1-300 gpurun_out/r2_bench_n2_nccl.json
