#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
export AAR_RIG_CACHE=/tmp/rigs
AAR_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_pair_tab|k_schur_prepare|k_residual|k_backsub" -s 8 -c 4 -f -o gpurun_out/r25_small python tools/quick_time.py --workload cfg4 --frames 20000 --iters 3 > gpurun_out/r25_ncu.log 2>&1
tail -2 gpurun_out/r25_ncu.log
