#!/bin/bash
# GPU check of the analytic variant, then the whole GPU suite and smoke (one gpurun call)
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out; rm -f gpurun_out/parity_achieved.jsonl
timeout 150 python -m pytest tests/test_gpu_analytic.py -q -m gpu -s 2>&1 | tail -40 > gpurun_out/r40_pytest_analytic.txt
tail -5 gpurun_out/r40_pytest_analytic.txt
cp gpurun_out/parity_achieved.jsonl gpurun_out/r40_parity_analytic.jsonl 2>/dev/null
timeout 240 python -m pytest tests -q -m gpu --deselect tests/test_gpu_analytic.py 2>&1 | tail -15 > gpurun_out/r40_pytest_gpu.txt
tail -3 gpurun_out/r40_pytest_gpu.txt
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 > gpurun_out/r40_smoke.txt
cat gpurun_out/r40_smoke.txt
