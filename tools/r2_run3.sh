#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 600 python tools/quick_time.py --workload cfg4 --frames 20000 --iters 6 --libs default,default:AAR_ASM_SKIP=1,default:AAR_ASM_SKIP=2,automatic-ar_b200/variants/mb4.so > gpurun_out/r2c_variants.txt 2>&1
grep "==\|ms/iter\|rror" gpurun_out/r2c_variants.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_asm" -c 2 -f -o gpurun_out/r2c_asm python tools/quick_time.py --workload cfg4 --frames 20000 --iters 1 > gpurun_out/r2c_ncu.log 2>&1
tail -3 gpurun_out/r2c_ncu.log
