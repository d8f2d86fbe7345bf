"""Times the initialisation path (include/aar_init.h) on a BASELINE workload: IPPE per detection, rig consensus, per-frame object
consensus.  Usage: time_init.py --workload cfg3 [--frames N] [--consensus-max K] [--check]   (--check compares with the CPU oracle)"""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "automatic-ar_b200", "python")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
from aar_b200 import binding, synth

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg2"); ap.add_argument("--frames", type=int, default=None)
ap.add_argument("--consensus-max", type=int, default=0); ap.add_argument("--check", action="store_true"); ap.add_argument("--objects-only", action="store_true")
a = ap.parse_args()
rig = synth.make_config(a.workload, frames=a.frames)
t0 = time.time(); g = binding.Initializer.from_rig(rig, consensus_max=a.consensus_max); t1 = time.time()
if a.objects_only:
    g.set_rig(rig.cam_ids, rig.T_cam_true, rig.marker_ids, rig.T_marker_true)
else:
    g.init_transforms()
t2 = time.time(); g.init_object_transforms(); t3 = time.time()
r = g.results(); tm = g.timings()
print(f"{a.workload}: {rig.N} detections, {rig.F} frames, consensus_max {a.consensus_max}: create {t1 - t0:.3f} s (IPPE kernel {tm['ippe_ms']:.3f} ms), "
      f"rig {t2 - t1:.3f} s (device {tm['rig_ms']:.3f} ms), objects {t3 - t2:.3f} s (device {tm['objects_ms']:.3f} ms), launches {tm['launches']}")
ci, cT = r["cams"]; fi, fT = r["objects"]
print("  cams with transform", len(ci), "max |T - truth|", float(np.abs(cT - rig.T_cam_true[np.searchsorted(rig.cam_ids, ci)]).max()),
      " objects", len(fi), "max |T - truth|", float(np.abs(fT - rig.T_frame_true[np.searchsorted(rig.frame_ids, fi)]).max()))
if a.check:
    import oracle_py
    nF = int(rig.frame_ids.max()) + 1
    t0 = time.time()
    o = oracle_py.InitOracle(rig.C, rig.K, rig.dist, float(rig.marker_size), nF, rig.det_frame, rig.det_cam, rig.det_marker, rig.det_xy, consensus_max=a.consensus_max)
    o.obtain_pose_estimations()
    if a.objects_only:
        o.set_rig(rig.cam_ids, rig.T_cam_true, rig.marker_ids, rig.T_marker_true); o.init_object_transforms()
    else:
        o.init_transforms()
    ro = o.results()
    print(f"  oracle (1 core): {time.time() - t0:.2f} s; bit-exact:", all(np.array_equal(ro[k][1], r[k][1]) for k in ("cams", "markers", "objects")))
