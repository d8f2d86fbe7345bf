"""Development aid: where do the GPU and oracle LM trajectories part?  (run under gpurun)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "automatic-ar_b200", "python")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
from aar_b200 import binding, synth
import oracle_py
rig = synth.make_config("cfg1")
o = oracle_py.Oracle(rig); p = binding.Problem(rig)
z0 = o.mats2evec()
z_o, fc_o, it_o, tr_o = o.solve(z0)
runs = [p.solve(z0) for _ in range(3)]
for z_g, fc_g, it_g, tr_g in runs:
    n = min(it_o, it_g)
    print("iters", it_o, it_g, "final", fc_o, fc_g, "rel", abs(fc_g - fc_o) / fc_o, "zdiff", np.abs(z_g - z_o).max())
    print(" per-iter rel cost diff:", np.abs(tr_g[:n, 0] - tr_o[:n, 0]) / tr_o[:n, 0])
print("gpu run-to-run final cost:", [r[1] for r in runs])
# one-step check: from the same z, compare delta of one iteration
prm = binding.Problem.default_params(max_iters=1, ignore_stop_rules=1)
z1_g = p.solve(z0, prm)[0]
o2 = oracle_py.Oracle(rig)
import ctypes
# oracle single iteration via port with maxIters... use trace of z after 1 iteration by running port with min_average_step huge is not exposed; compare through reduced system instead
S_o, b_o, c_o = o.reduced_system(z0, tr_o[0, 1] / 0.33 if False else 5.33493861e+08 / 1.0)
S_g, b_g, c_g = p.reduced_system(z0, 5.33493861e+08)
iu = np.triu_indices(p.n_r)
print("reduced S rel diff", np.abs(S_g[iu] - S_o[iu]).max() / np.abs(S_o).max(), "b", np.abs(b_g - b_o).max() / np.abs(b_o).max())
Sf = np.triu(S_o) + np.triu(S_o, 1).T
print("cond(S+muI)", np.linalg.cond(Sf + 5.33493861e+08 * np.eye(p.n_r)), "diag range", np.diag(Sf).min(), np.diag(Sf).max())
