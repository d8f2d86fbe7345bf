#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 600 python tools/quick_time.py --workload cfg4 --frames 20000 --iters 6 --libs default,default:AAR_ASM_SKIP=1,default:AAR_ASM_SKIP=2 > gpurun_out/r2e_variants.txt 2>&1
grep "==\|ms/iter\|rror" gpurun_out/r2e_variants.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_asm" -c 4 --csv --log-file gpurun_out/r2e_launches.csv env AAR_ASM_SKIP=1 python tools/quick_time.py --workload cfg4 --frames 20000 --iters 1 > gpurun_out/r2e_ncu.log 2>&1
grep -v "^==" gpurun_out/r2e_launches.csv | awk -F'","' '{print $5, $NF}' | tail -5
