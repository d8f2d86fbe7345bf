#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
export AAR_RIG_CACHE=/tmp/rigs
for c in 0 14; do
  AAR_LIB=$PWD/automatic-ar_b200/variants/solve_t$c.so timeout 300 python tools/quick_time.py --workload cfg4 --frames 2000 --iters 2 2>&1 | grep "solve timing" | tail -1
done
K="reduced_system or first_iterations or lm_solve or huber or edge_cases"
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$K" > gpurun_out/r22_pytest_subset.txt 2>&1; tail -5 gpurun_out/r22_pytest_subset.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_reduced_solve" -c 6 --csv --log-file gpurun_out/r22_solve_launches.csv python tools/quick_time.py --workload cfg4 --frames 5000 --iters 3 > /dev/null 2>&1
grep k_reduced gpurun_out/r22_solve_launches.csv | tail -3 | cut -c1-60,200-
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_reduced_solve" -c 6 --csv --log-file gpurun_out/r22_solve_launches_cfg2.csv python tools/quick_time.py --workload cfg2 --iters 3 > /dev/null 2>&1
grep k_reduced gpurun_out/r22_solve_launches_cfg2.csv | tail -2 | cut -c1-60,200-
