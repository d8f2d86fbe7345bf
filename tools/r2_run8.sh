#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
K="bit_exact or config_flags or reduced_system or edge_cases or first_iterations or exact_staging or tensor_core or huber"
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$K" > gpurun_out/r8_pytest_subset.txt 2>&1; tail -5 gpurun_out/r8_pytest_subset.txt
export AAR_RIG_CACHE=/tmp/rigs
timeout 600 python tools/quick_time.py --workload cfg4 --frames 20000 --iters 6 > gpurun_out/r8_variants.txt 2>&1
grep "==\|ms/iter\|rror" gpurun_out/r8_variants.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_asm" -c 2 -f -o gpurun_out/r8_asm python tools/quick_time.py --workload cfg4 --frames 20000 --iters 1 > gpurun_out/r8_ncu.log 2>&1
tail -3 gpurun_out/r8_ncu.log
