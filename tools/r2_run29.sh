#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
export AAR_RIG_CACHE=/tmp/rigs
K="bit_exact or config_flags or first_iterations or intrinsics or edge_cases"
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$K" > gpurun_out/r29_pytest_subset.txt 2>&1; tail -3 gpurun_out/r29_pytest_subset.txt
for lib in default mb3 mb4; do
  if [ $lib = default ]; then L=""; else L=$PWD/automatic-ar_b200/variants/$lib.so; fi
  AAR_LIB=$L AAR_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_pair_tab|k_residual|k_schur_prepare|k_backsub" -s 8 -c 8 --csv --log-file gpurun_out/r29_launches_$lib.csv python tools/quick_time.py --workload cfg4 --frames 20000 --iters 3 > /dev/null 2>&1
  echo "== $lib"; grep "k_" gpurun_out/r29_launches_$lib.csv | tail -4 | awk -F'","' '{print substr($5,1,30), $(NF)}'
done
