#!/bin/bash
# compute-sanitizer memcheck over the kernels of the analytic variant (tiny rig)
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
cat > /tmp/an_san.py <<'PY'
import sys
sys.path.insert(0, "automatic-ar_b200/python")
import numpy as np
from aar_b200 import binding, synth
rig = synth.make_rig(3, 6, 40, 6.0, seed=4)
rig.det_frame = np.concatenate([rig.det_frame, rig.det_frame[[3]]]); rig.det_cam = np.concatenate([rig.det_cam, rig.det_cam[[3]]])
rig.det_marker = np.concatenate([rig.det_marker, rig.det_marker[[3]]]); rig.det_xy = np.concatenate([rig.det_xy, rig.det_xy[[3]] + 0.25])
for huber in (False, True):
    p = binding.Problem(rig, with_huber=huber, analytic=True)
    z = p.mats2evec()
    r, ss = p.residual(z); cp, ri, v = p.jacobian(z); S, b, c = p.reduced_system(z, 10.0)
    z1, fc, it, tr = p.solve(z, binding.Problem.default_params(max_iters=6))
    print("analytic huber=%d" % huber, it, fc, len(v)); p.close()
PY
timeout 55 compute-sanitizer --tool memcheck --error-exitcode 9 python /tmp/an_san.py 2>&1 | tail -8 | tee gpurun_out/r44_sanitizer_analytic.txt
