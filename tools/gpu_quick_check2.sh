#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
export AAR_RIG_CACHE=/tmp/rigs
K="bit_exact or config_flags or first_iterations"
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$K" > gpurun_out/quick2_pytest.txt 2>&1; tail -2 gpurun_out/quick2_pytest.txt
AAR_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_pair_tab" -s 2 -c 3 --csv --log-file gpurun_out/quick2_launches.csv python tools/quick_time.py --workload cfg4 --frames 20000 --iters 3 > /dev/null 2>&1
grep "k_" gpurun_out/quick2_launches.csv | tail -3 | awk -F'","' '{print substr($5,1,30), $(NF)}'
