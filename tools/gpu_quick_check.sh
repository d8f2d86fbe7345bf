#!/bin/bash
# quick regression after a kernel change: LM parity subset + kernel times at cfg 4 (20 k frames) + the intrinsics block at cfg 3
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
export AAR_RIG_CACHE=/tmp/rigs
K="reduced_system or first_iterations or graph_resident or edge_cases or huber or lm_solve or intrinsics"
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$K" > gpurun_out/quick_pytest.txt 2>&1; tail -3 gpurun_out/quick_pytest.txt
AAR_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_backsub|k_schur_prepare|k_residual|k_pair_tab" -s 8 -c 8 --csv --log-file gpurun_out/quick_launches.csv python tools/quick_time.py --workload cfg4 --frames 20000 --iters 3 > /dev/null 2>&1
grep "k_" gpurun_out/quick_launches.csv | tail -4 | awk -F'","' '{print substr($5,1,30), $(NF)}'
python - <<'PY'
import sys, time
sys.path.insert(0, "automatic-ar_b200/python")
import numpy as np
from aar_b200 import binding, synth
rig = synth.make_config("cfg3")
for intr in (False, True):
    p = binding.Problem(rig, intrinsics=intr)
    z0 = p.mats2evec(); prm = binding.Problem.default_params(max_iters=12, ignore_stop_rules=1)
    p.solve(z0, prm); t = time.time(); z, c, it, tr = p.solve(z0, prm); dt = time.time() - t
    print("cfg3 intrinsics=%s: %.3f ms per LM iteration (aar_lm_solve, host z in/out), n_r %d, cost %.6g" % (intr, 1e3 * dt / it, p.n_r, c))
PY
