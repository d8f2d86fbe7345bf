#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
export AAR_RIG_CACHE=/tmp/rigs
for lib in default trk_mb3 trk_mb2; do
  if [ $lib = default ]; then L=""; else L=$PWD/automatic-ar_b200/variants/$lib.so; fi
  echo "== $lib"; AAR_LIB=$L timeout 300 python tools/time_track.py --frames 5000 2>&1 | tail -2
done
