#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
export AAR_RIG_CACHE=/tmp/rigs
K="reduced_system or first_iterations or graph_resident or edge_cases or huber"
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$K" > gpurun_out/r32_pytest_subset.txt 2>&1; tail -5 gpurun_out/r32_pytest_subset.txt
AAR_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_schur_prepare|k_pair_tab|k_residual|k_backsub" -s 8 -c 8 --csv --log-file gpurun_out/r32_launches.csv python tools/quick_time.py --workload cfg4 --frames 20000 --iters 3 > /dev/null 2>&1
grep "k_" gpurun_out/r32_launches.csv | tail -4 | awk -F'","' '{print substr($5,1,30), $(NF)}'
