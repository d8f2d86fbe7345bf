#!/bin/bash
# ncu --set full of the track kernel (cfg 5 density, 5000 frames) and of the initialisation kernels
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
export AAR_RIG_CACHE=/tmp/rigs
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_track_cta" -s 1 -c 1 -f -o gpurun_out/ncu_track python tools/time_track.py --frames 5000 > gpurun_out/ncu_track.log 2>&1; tail -2 gpurun_out/ncu_track.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_ippe|k_consensus|k_build_object" -c 4 -f -o gpurun_out/ncu_init python tools/time_init.py --workload cfg5 --frames 5000 --objects-only > gpurun_out/ncu_init.log 2>&1; tail -2 gpurun_out/ncu_init.log
