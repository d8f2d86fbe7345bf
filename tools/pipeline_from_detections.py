"""find_solution end to end on a BASELINE workload, from raw detections: Initializer on the device (IPPE per detection, rig and
object-pose consensus; include/aar_init.h) -> MultiCamMapper::solve on the device (include/aar_cuda.h).
Usage: pipeline_from_detections.py --workload cfg3 [--frames N] [--consensus-max K] [--max-iters I]"""
import argparse, copy, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "automatic-ar_b200", "python"))
import numpy as np
from aar_b200 import binding, synth

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg3"); ap.add_argument("--frames", type=int, default=None)
ap.add_argument("--consensus-max", type=int, default=256); ap.add_argument("--max-iters", type=int, default=10000)
a = ap.parse_args()
rig = synth.make_config(a.workload, frames=a.frames)
out = {"workload": a.workload, "cameras": rig.C, "markers": rig.M, "frames": rig.F, "detections": int(rig.N), "consensus_max": a.consensus_max}
t0 = time.time(); g = binding.Initializer.from_rig(rig, consensus_max=a.consensus_max); t1 = time.time()
g.init_transforms(); t2 = time.time(); g.init_object_transforms(); t3 = time.time()
r = g.results(); tm = g.timings()
out["init_s"] = {"create_and_ippe": t1 - t0, "rig": t2 - t1, "objects": t3 - t2, "device_ms": tm}
ci, cT = r["cams"]; mi, mT = r["markers"]; fi, fT = r["objects"]
assert np.array_equal(ci, rig.cam_ids) and np.array_equal(mi, rig.marker_ids), "rig not fully connected"
rig2 = copy.copy(rig)
keep = np.isin(rig.det_frame, fi)
rig2.frame_ids = fi; rig2.T_frame_true = rig.T_frame_true[np.searchsorted(rig.frame_ids, fi)]
rig2.det_frame, rig2.det_cam, rig2.det_marker, rig2.det_xy = rig.det_frame[keep], rig.det_cam[keep], rig.det_marker[keep], rig.det_xy[keep]
rig2.T_cam_init, rig2.T_marker_init, rig2.T_frame_init = cT, mT, fT
rig2.root_cam, rig2.root_marker = r["root_cam"], r["root_marker"]
out["init_dev_vs_truth"] = {"cams": float(np.abs(cT - rig.T_cam_true).max()), "markers": float(np.abs(mT - rig.T_marker_true).max()), "objects": float(np.abs(fT - rig2.T_frame_true).max())}
t0 = time.time(); p = binding.Problem(rig2); t1 = time.time()
z0 = p.mats2evec()
z, cost, iters, trace = p.solve(z0, binding.Problem.default_params(max_iters=a.max_iters)); t2 = time.time()
Tc, Tm, Tf = p.evec2mats(z)
out["solve"] = {"create_s": t1 - t0, "solve_s": t2 - t1, "iterations": int(iters), "initial_cost": float(trace[0, 0]) if len(trace) else None, "final_cost": float(cost),
                "rms_px": float(np.sqrt(cost / (8 * p.num_obs))), "cams_vs_truth": float(np.abs(Tc - rig.T_cam_true).max()), "markers_vs_truth": float(np.abs(Tm - rig.T_marker_true).max())}
print(json.dumps(out))
