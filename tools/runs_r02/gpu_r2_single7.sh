#!/bin/bash
# Round 2, seventh 1-GPU round trip (short): parity incl. the all-ranks-on-one-GPU variants after restoring the slot-based
# partition kernels; G = 8 on one GPU against the oracle.
mkdir -p gpurun_out
T=gpurun_out/r2z
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -p no:cacheprovider --deselect tests/test_gpu_full_size.py -k "not (end_to_end and (2] or 4]))" > ${T}_tests.log 2>&1
echo "pytest exit $?" | tee -a ${T}_tests.log; tail -4 ${T}_tests.log | cut -c1-300
timeout 300 python tools/dist_onegpu.py 8 2000000 2 1 > ${T}_onegpu_check.txt 2>&1; echo "onegpu check exit $?"; tail -1 ${T}_onegpu_check.txt | cut -c1-400
