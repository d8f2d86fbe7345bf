#!/bin/bash
# round 2, first GPU call: microbenchmarks that decide the assembly design + baseline parity + padded-stride A/B
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv | tee gpurun_out/r2a_gpu.txt
timeout 120 ./automatic-ar_b200/red_bench | tee gpurun_out/r2a_red_bench.txt
timeout 120 ./automatic-ar_b200/dmma_peak | tee gpurun_out/r2a_dmma_peak.txt
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2a_pytest_gpu.txt 2>&1; tail -5 gpurun_out/r2a_pytest_gpu.txt
timeout 600 python tools/quick_time.py --workload cfg4 --frames 20000 --iters 6 --libs default,automatic-ar_b200/variants/r1.so > gpurun_out/r2a_variants.txt 2>&1
grep "==\|ms/iter\|rror" gpurun_out/r2a_variants.txt
