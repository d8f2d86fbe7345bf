"""Development aid: time LM iterations of one workload on cuda:0 with per-phase device times."""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "automatic-ar_b200", "python"))
import numpy as np
from aar_b200 import binding, synth

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg3"); ap.add_argument("--frames", type=int, default=None); ap.add_argument("--iters", type=int, default=8)
a = ap.parse_args()
t = time.time(); rig = synth.make_config(a.workload, frames=a.frames); t_gen = time.time() - t
t = time.time(); p = binding.Problem(rig); t_create = time.time() - t
z0 = p.mats2evec()
print(f"{a.workload}: C={rig.C} M={rig.M} F={rig.F} N={p.num_obs} n_r={p.n_r} gen {t_gen:.1f}s create {t_create:.2f}s", flush=True)
prm = binding.Problem.default_params(ignore_stop_rules=1)
p.lm_begin(z0, prm); p.lm_iterate(2)
p.set_profiling(True)
import torch
torch.cuda.synchronize(); t = time.time()
rep, tr = p.lm_iterate(a.iters, trace_capacity=a.iters)
torch.cuda.synchronize(); dt = time.time() - t
print("trace cost:", tr[:, 0], "tries", tr[:, 3])
ph = p.phase_ms()
print(f"{dt / a.iters * 1e3:.3f} ms/iter  {4 * p.num_obs * a.iters / dt / 1e9:.3f} G corner-obs/s; phases (ms/iter):", {k: round(v / a.iters, 3) for k, v in ph.items()})
print(f"jacobian kernel: {p.num_obs * 9.0e3 / (ph['jacobian'] / a.iters * 1e-3) / 1e12:.2f} TFLOP/s algorithmic (9.0 kflop/obs)")
