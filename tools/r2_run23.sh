#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
export AAR_RIG_CACHE=/tmp/rigs
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "intrinsics" > gpurun_out/r23_pytest_intr.txt 2>&1; tail -30 gpurun_out/r23_pytest_intr.txt
