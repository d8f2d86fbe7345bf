"""Development aid: one opt-in accumulation mode (AAR_ACC_MMA=<mode>) against the default kernel — normal equations on a small rig
with duplicates / an emptied frame, and device time of the accumulation phase on cfg 4 restricted to a few thousand frames.
Sized to finish in a few seconds of GPU-box time.  Usage: mini_check.py MODE [frames]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "automatic-ar_b200", "python"))
import numpy as np
from aar_b200 import binding, synth

mode = sys.argv[1]; frames = int(sys.argv[2]) if len(sys.argv) > 2 else 4000


def setmode(m):
    if m is None: os.environ.pop("AAR_ACC_MMA", None)
    else: os.environ["AAR_ACC_MMA"] = m


rig = synth.make_rig(C=3, M=6, F=120, obs_per_frame=6.0, seed=31)
dup = np.array([5, 40, 41], np.int64)
rig.det_frame = np.concatenate([rig.det_frame, rig.det_frame[dup]]); rig.det_cam = np.concatenate([rig.det_cam, rig.det_cam[dup]])
rig.det_marker = np.concatenate([rig.det_marker, rig.det_marker[dup]]); rig.det_xy = np.concatenate([rig.det_xy, rig.det_xy[dup] + 0.5])
rig.det_marker = rig.det_marker.copy(); rig.det_marker[rig.det_frame == rig.frame_ids[7]] = 99999
big = synth.make_config("cfg4", frames=frames)
out = {}
for m in (None, mode):
    setmode(m)
    p = binding.Problem(rig); z0 = p.mats2evec()
    S, b, c = p.reduced_system(z0, 5.0)
    q = binding.Problem(big); zb = q.mats2evec()
    q.lm_begin(zb, binding.Problem.default_params(ignore_stop_rules=1)); q.lm_iterate(2); q.set_profiling(True)
    rep, tr = q.lm_iterate(3, trace_capacity=3)
    ph = q.phase_ms()
    out[m] = (S, b, c, tr[:, 0].copy(), ph["accumulate_kernel"] / 3, ph["jacobian_kernel"] / 3)
S0, b0, c0, t0, a0, j0 = out[None]; S1, b1, c1, t1, a1, j1 = out[mode]
iu = np.triu_indices(len(b0))
dS = np.abs(S1[iu] - S0[iu]).max() / np.abs(S0).max(); db = np.abs(b1 - b0).max() / np.abs(b0).max()
dt = np.abs(t1 - t0).max() / np.abs(t0).max()
print(f"mode {mode}: N={q.num_obs} dS={dS:.2e} db={db:.2e} cost_equal={c1 == c0} trace_rel_diff={dt:.2e} "
      f"accumulate {a0:.3f} -> {a1:.3f} ms, project {j0:.3f} -> {j1:.3f} ms  {'OK' if dS <= 1e-12 and db <= 1e-12 and dt <= 1e-9 else 'MISMATCH'}")
