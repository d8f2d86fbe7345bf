#!/bin/bash
# Development aid: one gpurun call = divcheck + parity subset + timing of the library variants under automatic-ar_b200/variants/
cd "${GRAFT_REPO_ROOT:-.}"
TAG=${1:-r1x}
mkdir -p gpurun_out
timeout 120 ./automatic-ar_b200/divcheck > gpurun_out/${TAG}_divcheck.txt 2>&1
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "bit_exact or config_flags or reduced_system or edge_cases or first_iterations" > gpurun_out/${TAG}_pytest_subset.txt 2>&1
tail -3 gpurun_out/${TAG}_pytest_subset.txt
LIBS=default; for f in automatic-ar_b200/variants/*.so; do [ -f "$f" ] && LIBS=$LIBS,$f; done
timeout 400 python tools/quick_time.py --workload cfg4 --frames 20000 --iters 6 --libs $LIBS > gpurun_out/${TAG}_variants.txt 2>&1
grep "==\|ms/iter" gpurun_out/${TAG}_variants.txt
cat gpurun_out/${TAG}_divcheck.txt
