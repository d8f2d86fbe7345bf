"""Generates tests/golden/init_golden.npz — OpenCV (cv2 4.13, build container) known answers for the initialisation
path, which pin oracle/init_oracle.cpp:

    cv::undistortPoints without P   3rdparty/aruco/aruco/ippe.cpp:164      -> bit-exact
    aruco IPPE (solvePnP_)          ippe.cpp:118-219                      -> cv2.solvePnPGeneric(SOLVEPNP_IPPE_SQUARE), OpenCV's own
                                                                             port of the same algorithm (Collins & Bartoli): both
                                                                             solutions and their order, to a tolerance
    find_best_transformation        libs/initializer.cpp:156-205          -> a twin that calls cv2 for every matrix operation
                                                                             (gemm, invert, subtract, multiply, reduce, sqrt, sumElems):
                                                                             winner index and consensus error bit-exact
Run once in the build container:  python tests/golden/make_golden_init.py   (the GPU box only reads the .npz).
"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "automatic-ar_b200", "python"))
from aar_b200 import synth  # noqa: E402


def gemm(A, B):
    return cv2.gemm(np.ascontiguousarray(A), np.ascontiguousarray(B), 1.0, None, 0.0)


def twin_consensus(marker_size, T, T1inv, T2inv):
    """Initializer::find_best_transformation with cv2 calls (initializer.cpp:156-205)."""
    h = marker_size / 2
    pts = np.array([[-h, h, h, -h], [h, h, -h, -h], [0, 0, 0, 0], [1, 1, 1, 1]], dtype=np.float64)
    best, best_err = -1, np.finfo(np.float64).max
    for i in range(len(T)):
        curr = 0.0
        for j in range(len(T)):
            p2 = gemm(gemm(gemm(T2inv[j], T[i]), T1inv[j]), pts)
            diff = cv2.subtract(pts, p2)[0:3]
            sq = cv2.multiply(diff, diff)
            red = cv2.reduce(sq, 0, cv2.REDUCE_SUM)
            red = cv2.sqrt(red)
            curr += cv2.sumElems(red)[0]
        if curr < best_err:
            best, best_err = i, curr
    return best, best_err


def main():
    rng = np.random.default_rng(20260002)
    out = {}
    # ---- detections of a distorted rig: realistic corner quadruples with their camera
    rig = synth.make_rig(C=4, M=8, F=40, obs_per_frame=10.0, seed=11, distorted=True)
    sel = rng.choice(rig.N, size=min(400, rig.N), replace=False)
    xy = rig.det_xy[sel]; cam = rig.det_cam[sel]
    out["K"] = rig.K; out["dist"] = rig.dist; out["marker_size"] = np.float64(rig.marker_size)
    out["xy"] = xy; out["cam"] = cam
    # ---- undistortPoints (normalised)
    und = np.zeros_like(xy)
    for n in range(len(xy)):
        c = int(cam[n])
        und[n] = cv2.undistortPoints(xy[n].reshape(4, 1, 2), rig.K[c], rig.dist[c]).reshape(8)
    out["und_norm"] = und
    # ---- cv2 IPPE_SQUARE: both solutions (sorted by reprojection error like aruco's)
    s = float(np.float32(rig.marker_size)); hs = s / 2
    obj = np.array([[-hs, hs, 0], [hs, hs, 0], [hs, -hs, 0], [-hs, -hs, 0]], dtype=np.float64)
    R_all = np.zeros((len(xy), 2, 3, 3)); t_all = np.zeros((len(xy), 2, 3)); e_all = np.zeros((len(xy), 2))
    for n in range(len(xy)):
        c = int(cam[n])
        ok, rvecs, tvecs, errs = cv2.solvePnPGeneric(obj, xy[n].reshape(4, 1, 2).astype(np.float64), rig.K[c], rig.dist[c], flags=cv2.SOLVEPNP_IPPE_SQUARE)
        assert ok and len(rvecs) == 2
        for k in range(2):
            R_all[n, k] = cv2.Rodrigues(rvecs[k])[0]; t_all[n, k] = tvecs[k].reshape(3); e_all[n, k] = float(errs[k])
    out["ippe_R"] = R_all; out["ippe_t"] = t_all; out["ippe_rms"] = e_all
    # ---- consensus twin on random candidate triples (realistic: noisy versions of one transform)
    cases = []
    for case in range(12):
        n = int(rng.integers(2, 40))
        base = np.eye(4); base[:3, :3] = cv2.Rodrigues(rng.normal(0, 1, 3))[0]; base[:3, 3] = rng.normal(0, 0.5, 3)
        T = np.zeros((n, 4, 4)); A = np.zeros((n, 4, 4)); B = np.zeros((n, 4, 4))
        for i in range(n):
            def noisy(M, s):
                D = np.eye(4); D[:3, :3] = cv2.Rodrigues(rng.normal(0, s, 3))[0]; D[:3, 3] = rng.normal(0, s / 4, 3)
                return gemm(D, M).astype(np.float32).astype(np.float64)
            P1 = noisy(np.eye(4), 0.8); P1[:3, 3] += [0, 0, 1.5]
            P2 = noisy(gemm(base, P1), 0.03)
            T[i] = gemm(P2, cv2.invert(P1)[1]); A[i] = P1; B[i] = cv2.invert(P2)[1]
        bi, be = twin_consensus(0.05, T, A, B)
        cases.append((T, A, B, bi, be))
    out["cons_n"] = np.array([len(c[0]) for c in cases])
    out["cons_T"] = np.concatenate([c[0] for c in cases]); out["cons_T1inv"] = np.concatenate([c[1] for c in cases]); out["cons_T2inv"] = np.concatenate([c[2] for c in cases])
    out["cons_best"] = np.array([c[3] for c in cases]); out["cons_err"] = np.array([c[4] for c in cases])
    np.savez_compressed(os.path.join(HERE, "init_golden.npz"), **out)
    print("wrote init_golden.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
