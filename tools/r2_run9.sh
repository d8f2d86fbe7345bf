#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
export AAR_RIG_CACHE=/tmp/rigs
K="bit_exact or config_flags or reduced_system or edge_cases or first_iterations or exact_staging or tensor_core or huber or track"
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$K" > gpurun_out/r9_pytest_subset.txt 2>&1; tail -5 gpurun_out/r9_pytest_subset.txt
timeout 600 python tools/quick_time.py --workload cfg4 --frames 20000 --iters 6 > gpurun_out/r9_variants.txt 2>&1
grep "==\|ms/iter\|rror" gpurun_out/r9_variants.txt
timeout 300 python tools/time_track.py --frames 5000 > gpurun_out/r9_track.txt 2>&1
AAR_TRACK=warp timeout 300 python tools/time_track.py --frames 5000 >> gpurun_out/r9_track.txt 2>&1
cat gpurun_out/r9_track.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_asm|k_track_cta" -c 3 -f -o gpurun_out/r9_asm python tools/quick_time.py --workload cfg4 --frames 20000 --iters 1 > gpurun_out/r9_ncu.log 2>&1
tail -3 gpurun_out/r9_ncu.log
