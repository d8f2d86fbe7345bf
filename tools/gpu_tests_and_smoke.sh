#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
export AAR_RIG_CACHE=/tmp/rigs
rm -f gpurun_out/parity_achieved.jsonl
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/last_pytest_gpu.txt 2>&1; tail -4 gpurun_out/last_pytest_gpu.txt
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1 | tee gpurun_out/last_smoke.txt
timeout 600 python bench.py --workload cfg2 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | head -c 300; echo
