#!/bin/bash
# Round 2: the two-batch radix pass (radix_pass2_kernel) against the one-batch kernel: tools/radix_lab at six sizes (checked
# against std::stable_sort up to 50 M pairs), the parity tests that sort, bench lines of C2 and c3 with either kernel.
mkdir -p gpurun_out
T=gpurun_out/r2j
for B in 1 2; do
  for spec in 1000:31:1 100003:31:1 3333954:24:1 5000000:25:1 50000000:31:1 400000000:31:0; do
    n=${spec%%:*}; r=${spec#*:}; bits=${r%%:*}; chk=${r##*:}
    echo "== batches=$B n=$n bits=$bits"
    SMJ_RADIX_BATCHES=$B timeout 200 tools/bin/radix_lab $n $bits $chk 2>&1 | tail -3
  done
done > ${T}_radix_lab.txt 2>&1
grep -c "check: 0 mismatches, device flag 0" ${T}_radix_lab.txt; grep "best\|mismatch" ${T}_radix_lab.txt | cut -c1-150
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_zz_gpu_scan_chunk.py tests/test_gpu_dpu_stages.py -m gpu -x -q --timeout 300 -p no:cacheprovider > ${T}_tests.log 2>&1
echo "pytest exit $?"; tail -3 ${T}_tests.log | cut -c1-200
for B in 1 2; do
  SMJ_RADIX_BATCHES=$B SMJ_BENCH_NO_EAGER=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > ${T}_bench_c2_b$B.json 2> ${T}_bench_c2_b$B.err
  echo "c2 batches=$B exit $?"; python -c "import json; d=json.loads(open('${T}_bench_c2_b$B.json').read()); print(round(d['ms_per_step'],4), d['stage_ms'], round(d['roofline']['frac'],3), d['config']['result_checksum'])"
done
for B in 1 2; do
  SMJ_RADIX_BATCHES=$B timeout 300 python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > ${T}_bench_c3_b$B.json 2> ${T}_bench_c3_b$B.err
  echo "c3 batches=$B exit $?"; python -c "import json; d=json.loads(open('${T}_bench_c3_b$B.json').read()); print(round(d['ms_per_step'],3), d['stage_ms'], round(d['roofline']['frac'],3), d['config']['rows_joined'], d['config']['many_to_many_count'])"
done
