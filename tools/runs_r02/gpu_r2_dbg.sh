#!/bin/bash
mkdir -p gpurun_out
T=gpurun_out/r2dbg
CUDA_LAUNCH_BLOCKING=1 timeout 600 python -m pytest tests/test_gpu_full_size.py -m gpu -x -q --timeout 500 -p no:cacheprovider -k "zipf_200M" > ${T}_zipf_blocking.log 2>&1
echo "zipf alone (blocking) exit $?"; grep -E "Error|error|passed|failed" ${T}_zipf_blocking.log | head -5 | cut -c1-300
timeout 900 python -m pytest tests/test_gpu_dpu_stages.py tests/test_gpu_full_size.py -m gpu -x -q --timeout 800 -p no:cacheprovider > ${T}_full_size.log 2>&1
echo "dpu stages + full size exit $?"; grep -E "Error|error|passed|failed" ${T}_full_size.log | head -5 | cut -c1-300
