#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
export AAR_RIG_CACHE=/tmp/rigs
timeout 900 python -m pytest tests/test_gpu_init.py -x -q -m gpu > gpurun_out/r14_pytest_init.txt 2>&1; tail -15 gpurun_out/r14_pytest_init.txt
for w in "cfg1 --check" "cfg2 --check" "cfg3 --consensus-max 256 --check" "cfg3 --frames 1000 --check" "cfg5 --frames 5000 --objects-only" "cfg5 --frames 5000 --objects-only --consensus-max 128"; do
  timeout 900 python tools/time_init.py --workload $w >> gpurun_out/r14_time_init.txt 2>&1
done
cat gpurun_out/r14_time_init.txt
