#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
export AAR_RIG_CACHE=/tmp/rigs
timeout 900 python -m pytest tests/test_host_facade.py tests/test_gpu_init.py -x -q -m gpu > gpurun_out/r15_pytest_facade.txt 2>&1; tail -25 gpurun_out/r15_pytest_facade.txt
