#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
export AAR_RIG_CACHE=/tmp/rigs
K="reduced_system or first_iterations or lm_solve or huber or edge_cases"
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$K" > gpurun_out/r18_pytest_subset.txt 2>&1; tail -5 gpurun_out/r18_pytest_subset.txt
for v in 1 0; do
  echo "== AAR_CLUSTER_SOLVE_V1=$v" >> gpurun_out/r18_variants.txt
  AAR_CLUSTER_SOLVE_V1=$v timeout 600 python tools/quick_time.py --workload cfg4 --frames 20000 --iters 6 >> gpurun_out/r18_variants.txt 2>&1
  AAR_CLUSTER_SOLVE_V1=$v timeout 600 python tools/quick_time.py --workload cfg2 --iters 6 >> gpurun_out/r18_variants.txt 2>&1
done
grep "==\|ms/iter\|rror" gpurun_out/r18_variants.txt
AAR_CLUSTER_SOLVE_V1=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_reduced_solve" -c 6 --csv --log-file gpurun_out/r18_solve_launches.csv python tools/quick_time.py --workload cfg4 --frames 5000 --iters 3 > /dev/null 2>&1
AAR_CLUSTER_SOLVE_V1=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_reduced_solve" -c 6 --csv --log-file gpurun_out/r18_solve_launches_v1.csv python tools/quick_time.py --workload cfg4 --frames 5000 --iters 3 > /dev/null 2>&1
grep k_reduced gpurun_out/r18_solve_launches.csv | tail -3 | cut -c1-60,200-; grep k_reduced gpurun_out/r18_solve_launches_v1.csv | tail -3 | cut -c1-60,200-
