#!/bin/bash
# Round-end capture: parity tests, smoke, the bench lines (own arm, reference arm, c3 / c4), launch list + full ncu captures.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/tests.log; tail -3 gpurun_out/tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_r01_n1.json 2> gpurun_out/bench_r01_n1.err; echo "bench exit $?"; cat gpurun_out/bench_r01_n1.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r01_ref.json 2> gpurun_out/bench_r01_ref.err; echo "ref exit $?"; cut -c1-400 gpurun_out/bench_r01_ref.json
for w in c4 c3; do
  timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  echo "bench $w exit $?"; cut -c1-330 gpurun_out/bench_$w.json
done
KERNELS="radix_pass:6 select_tma:2 bloom_filter:1 plan_compact:1 join_match:1 join_materialize:1" bash tools/gpu_profile.sh > gpurun_out/profile_run.log 2>&1; tail -3 gpurun_out/profile_run.log
