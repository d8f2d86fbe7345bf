#!/bin/bash
# compute-sanitizer over the small end-to-end paths: LM solve (graph-resident loop, cluster Cholesky), Initializer, intrinsics block
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
cat > /tmp/san_small.py <<'PY'
import sys
sys.path.insert(0, "automatic-ar_b200/python")
import numpy as np
from aar_b200 import binding, synth
rig = synth.make_rig(C=4, M=12, F=40, obs_per_frame=14.0, seed=2)          # n_r = 84: three block rows in the cluster Cholesky
p = binding.Problem(rig); z, c, it, tr = p.solve(p.mats2evec()); print("solve", it, c)
q = binding.Problem(rig, intrinsics=True, with_huber=True); z, c, it, tr = q.solve(q.mats2evec(), binding.Problem.default_params(max_iters=4)); print("intrinsics + huber", it, c)
g = binding.Initializer.from_rig(rig, consensus_max=20); g.init_transforms(); g.init_object_transforms(); print("init", len(g.results()["objects"][0]))
t = binding.Problem(rig, cams=False, markers=False, objects=True); zz, cc, ii = t.track_batch(t.mats2evec().reshape(-1, 6)); print("track", ii.max())
PY
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python /tmp/san_small.py > gpurun_out/sanitizer_memcheck.txt 2>&1; echo "memcheck rc=$?"; tail -6 gpurun_out/sanitizer_memcheck.txt
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python /tmp/san_small.py > gpurun_out/sanitizer_racecheck.txt 2>&1; echo "racecheck rc=$?"; tail -6 gpurun_out/sanitizer_racecheck.txt
