#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
export AAR_RIG_CACHE=/tmp/rigs
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r35_pytest_gpu.txt 2>&1; tail -6 gpurun_out/r35_pytest_gpu.txt
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1
