#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
export AAR_RIG_CACHE=/tmp/rigs
AAR_CREATE_TIMING=1 python - <<'PY' 2>&1 | tee gpurun_out/create_timing.txt
import sys, time
sys.path.insert(0, "automatic-ar_b200/python")
from aar_b200 import binding, synth
rig = synth.make_config("cfg4")
for k in range(2):
    t = time.time(); p = binding.Problem(rig); print("Problem(rig) total %.3f s" % (time.time() - t), flush=True); p.close()
PY
