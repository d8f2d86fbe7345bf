#!/bin/bash
# round 2: camera-major pair pass with warp-private camera x marker tables
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
K="bit_exact or config_flags or reduced_system or edge_cases or first_iterations or exact_staging or tensor_core or huber"
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$K" > gpurun_out/r7_pytest_subset.txt 2>&1; tail -15 gpurun_out/r7_pytest_subset.txt
timeout 600 python tools/quick_time.py --workload cfg4 --frames 20000 --iters 6 > gpurun_out/r7_variants.txt 2>&1
grep "==\|ms/iter\|rror" gpurun_out/r7_variants.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_asm|k_jac|k_pair" -c 12 --csv --log-file gpurun_out/r7_launches.csv python tools/quick_time.py --workload cfg4 --frames 20000 --iters 1 > gpurun_out/r7_ncu.log 2>&1
grep -v "^==" gpurun_out/r7_launches.csv | awk -F'","' '{print $5, $NF}' | tail -8
