"""Development aid: run the C-ABI entry points one by one on a small rig, flushing progress (run under gpurun with a timeout)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "automatic-ar_b200", "python")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
from aar_b200 import binding, synth
import oracle_py
def say(*a):
    print(*a, flush=True)
F = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rig = synth.make_rig(C=3, M=6, F=F, obs_per_frame=6.0, seed=3)
o = oracle_py.Oracle(rig); say("oracle ok")
p = binding.Problem(rig); say("problem created", p.num_obs, p.num_vars)
z = o.mats2evec()
r, ss = p.residual(z); say("residual ok", np.array_equal(r, o.error(z)))
cp, ri, v = p.jacobian(z); cpo, rio, vo = o.jacobian(z); say("jacobian dump ok", np.array_equal(v, vo), np.abs(v - vo).max())
t = time.time(); S, b, c = p.reduced_system(z, 1234.5); say("reduced system done in", time.time() - t)
So, bo, co = o.reduced_system(z, 1234.5)
iu = np.triu_indices(p.n_r)
say("S rel err", np.abs(S[iu] - So[iu]).max() / np.abs(So).max(), "b rel err", np.abs(b - bo).max() / np.abs(bo).max(), "cost", c, co)
for k in (1, 3, 50):
    t = time.time(); zg, fc, it, tr = p.solve(z, binding.Problem.default_params(max_iters=k)); say("solve", k, "iters", it, "cost", fc, "time", time.time() - t)
o.set_max_iters(50); zo, fco, ito, tro = o.solve(z); say("oracle", ito, fco)
