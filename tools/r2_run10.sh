#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
export AAR_RIG_CACHE=/tmp/rigs
K="bit_exact or config_flags or reduced_system or edge_cases or first_iterations or exact_staging or tensor_core or huber or track"
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$K" > gpurun_out/r10_pytest_subset.txt 2>&1; tail -5 gpurun_out/r10_pytest_subset.txt
timeout 600 python tools/quick_time.py --workload cfg4 --frames 20000 --iters 6 > gpurun_out/r10_variants.txt 2>&1
grep "==\|ms/iter\|rror" gpurun_out/r10_variants.txt
for lib in default automatic-ar_b200/variants/trk3.so automatic-ar_b200/variants/trk2.so; do
  echo "== $lib" >> gpurun_out/r10_track.txt
  if [ $lib = default ]; then timeout 300 python tools/time_track.py --frames 5000 >> gpurun_out/r10_track.txt 2>&1; else AAR_LIB=$PWD/$lib timeout 300 python tools/time_track.py --frames 5000 >> gpurun_out/r10_track.txt 2>&1; fi
done
cat gpurun_out/r10_track.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_asm" -c 2 -f -o gpurun_out/r10_asm python tools/quick_time.py --workload cfg4 --frames 20000 --iters 1 > gpurun_out/r10_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_track_cta" -c 1 -f -o gpurun_out/r10_trk python tools/time_track.py --frames 5000 > gpurun_out/r10_ncu2.log 2>&1
tail -2 gpurun_out/r10_ncu.log gpurun_out/r10_ncu2.log
