"""Development aid: throughput of aar_track_batch (BASELINE config 5: independent per-frame pose LM against the fixed rig)."""
import argparse, copy, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "automatic-ar_b200", "python"))
import numpy as np
from aar_b200 import binding, synth
ap = argparse.ArgumentParser(); ap.add_argument("--frames", type=int, default=5000); ap.add_argument("--workload", default="cfg5")
a = ap.parse_args()
rig = copy.copy(synth.make_config(a.workload, frames=a.frames))
rig.T_cam_init, rig.T_marker_init = rig.T_cam_true, rig.T_marker_true
p = binding.Problem(rig, cams=False, markers=False, objects=True)
z0 = p.mats2evec().reshape(-1, 6)
p.track_batch(z0)                                   # warm-up
t = time.time(); z, cost, its = p.track_batch(z0); dt = time.time() - t
p.track_upload(z0); p.set_profiling(True)
for _ in range(3): p.track_run()
ms, runs = p.track_ms()
print(f"device time of the solves alone: {ms / runs:.2f} ms per pass ({os.environ.get('AAR_TRACK', 'cta')} kernel)")
print(f"{a.workload}: {rig.F} frames, {p.num_obs} observations: {dt * 1e3:.1f} ms -> {rig.F / dt:.0f} frames/s, {4 * p.num_obs / dt / 1e9:.3f} G corner-obs/s per solve; "
      f"iterations mean {its.mean():.1f} max {its.max()}, rms {np.sqrt(cost.sum() / (8 * p.num_obs)):.3f} px, "
      f"translation error max {np.abs(z[:, 3:] - rig.T_frame_true[:, :3, 3]).max():.2e} m")
