"""Generates tests/golden/intrinsics_golden.npz — a cv2-driven twin (cv2 4.13, build container) of the camera-intrinsics block of
MultiCamMapper::jacobian_function (libs/multicam_mapper.cpp:835-893: fx, cx, fy, cy and the five distortion coefficients of every
camera perturbed by +-J_delta, central difference of float32 projections of the RAW corners) and of error_function with the camera
matrices taken from io_vec (intrinsics_vec2mats, :580-593), on the rig stored in cv2_golden.npz.  Every matrix operation is a cv2 call.
Run once in the build container:  python tests/golden/make_golden_intrinsics.py"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..")); sys.path.insert(0, os.path.join(HERE, "..", "..", "automatic-ar_b200", "python"))
from conftest import rig_from_golden  # noqa: E402
from make_golden_cv2 import gemm, twin_project, vec2mat  # noqa: E402


def main():
    g = np.load(os.path.join(HERE, "cv2_golden.npz"))
    rig = rig_from_golden(g)
    C, M, F, N = rig.C, rig.M, rig.F, rig.N
    z_pose = g["twin_z"]
    intr = np.concatenate([np.concatenate([[rig.K[c, 0, 0], rig.K[c, 0, 2], rig.K[c, 1, 1], rig.K[c, 1, 2]], rig.dist[c]]) for c in range(C)])
    rng = np.random.default_rng(77)
    intr = intr + np.tile(np.concatenate([rng.normal(0, 0.7, 4), rng.normal(0, 1e-3, 5)]), C) * 1.0      # io_vec need not equal the camera files
    z = np.concatenate([z_pose, intr])
    h = np.float32(rig.marker_size) / np.float32(2)
    X = np.array([[-h, h, 0, 1], [h, h, 0, 1], [h, -h, 0, 1], [-h, -h, 0, 1]], dtype=np.float64).T.copy()
    Tc = [np.eye(4)] + [vec2mat(z[6 * i:6 * i + 6]) for i in range(C - 1)]
    o = 6 * (C - 1)
    Tm = [np.eye(4)] + [vec2mat(z[o + 6 * i:o + 6 * i + 6]) for i in range(M - 1)]
    o += 6 * (M - 1)
    To = [vec2mat(z[o + 6 * i:o + 6 * i + 6]) for i in range(F)]
    o += 6 * F
    cam_index = {int(c): i for i, c in enumerate(rig.cam_ids)}; marker_index = {int(m): i for i, m in enumerate(rig.marker_ids)}
    frame_index = {int(f): i for i, f in enumerate(rig.frame_ids)}

    def Kof(c, dK=None):
        K = np.eye(3); v = z[o + 9 * c:o + 9 * c + 4].copy()
        if dK is not None:
            v[dK[0]] += dK[1]
        K[0, 0], K[0, 2], K[1, 1], K[1, 2] = v
        return K

    und = g["twin_und"]; raw = rig.det_xy
    r = np.zeros(8 * N)
    for i in range(N):
        ci = cam_index[int(rig.det_cam[i])]; mi = marker_index[int(rig.det_marker[i])]; fi = frame_index[int(rig.det_frame[i])]
        px, py = twin_project(Tc[ci], To[fi], Tm[mi], Kof(ci), X, ci == 0, mi == 0)
        r[8 * i + 0:8 * i + 8:2] = (und[i, 0::2] - px).astype(np.float64); r[8 * i + 1:8 * i + 8:2] = (und[i, 1::2] - py).astype(np.float64)
    delta = 0.001
    J = np.zeros((8 * N, 9 * C))
    for c in range(C):
        for k in range(9):
            for i in range(N):
                ci = cam_index[int(rig.det_cam[i])]
                if ci != c:
                    continue
                mi = marker_index[int(rig.det_marker[i])]; fi = frame_index[int(rig.det_frame[i])]
                e = []
                for sg in (+delta, -delta):
                    K = Kof(c, (k, sg)) if k < 4 else Kof(c)          # the distortion coefficients never reach project_marker (:608-649)
                    px, py = twin_project(Tc[ci], To[fi], Tm[mi], K, X, ci == 0, mi == 0)
                    v = np.zeros(8); v[0::2] = (raw[i, 0::2] - px).astype(np.float64); v[1::2] = (raw[i, 1::2] - py).astype(np.float64)
                    e.append(v)
                J[8 * i:8 * i + 8, 9 * c + k] = (e[0] - e[1]) / (2 * delta)
    np.savez_compressed(os.path.join(HERE, "intrinsics_golden.npz"), z=z, r=r, J_intr=J)
    print("wrote intrinsics_golden.npz", z.shape, r.shape, J.shape, "non-zero columns", int((np.abs(J).sum(0) > 0).sum()))


if __name__ == "__main__":
    main()
