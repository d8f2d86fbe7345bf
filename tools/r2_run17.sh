#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
export AAR_RIG_CACHE=/tmp/rigs
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_reduced_solve" -s 2 -c 1 -f -o gpurun_out/r17_solve python tools/quick_time.py --workload cfg4 --frames 5000 --iters 3 > gpurun_out/r17_ncu.log 2>&1
tail -2 gpurun_out/r17_ncu.log
