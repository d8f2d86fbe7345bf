#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
export AAR_RIG_CACHE=/tmp/rigs
for c in 0 7 14; do
  AAR_LIB=$PWD/automatic-ar_b200/variants/solve_t$c.so timeout 300 python tools/quick_time.py --workload cfg4 --frames 2000 --iters 2 2>&1 | grep "solve timing" | tail -1
done
timeout 300 python tools/quick_time.py --workload cfg4 --frames 5000 --iters 3 2>&1 | grep "ms/iter"
