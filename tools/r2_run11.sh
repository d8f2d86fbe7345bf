#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
export AAR_RIG_CACHE=/tmp/rigs
K="bit_exact or config_flags or reduced_system or edge_cases or first_iterations or exact_staging or tensor_core or huber or track"
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$K" > gpurun_out/r11_pytest_subset.txt 2>&1; tail -5 gpurun_out/r11_pytest_subset.txt
timeout 600 python tools/quick_time.py --workload cfg4 --frames 20000 --iters 6 > gpurun_out/r11_variants.txt 2>&1
grep "==\|ms/iter\|rror" gpurun_out/r11_variants.txt
timeout 300 python tools/time_track.py --frames 5000 > gpurun_out/r11_track.txt 2>&1; cat gpurun_out/r11_track.txt
timeout 900 python bench.py --steps 6 --warmup 3 > gpurun_out/r11_bench_cfg4.json 2> gpurun_out/r11_bench_cfg4.err; tail -c 3000 gpurun_out/r11_bench_cfg4.json; tail -5 gpurun_out/r11_bench_cfg4.err
timeout 900 python bench.py --workload cfg5 --frames 20000 --steps 3 --warmup 3 > gpurun_out/r11_bench_cfg5.json 2> gpurun_out/r11_bench_cfg5.err; tail -c 2500 gpurun_out/r11_bench_cfg5.json; tail -5 gpurun_out/r11_bench_cfg5.err
