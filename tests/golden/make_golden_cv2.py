"""Generates tests/golden/cv2_golden.npz — known answers from OpenCV (cv2 4.13 in the build
container) for every OpenCV call the reference makes on the hot path, plus a cv2-driven twin of
MultiCamMapper::error_function / jacobian_function on a small rig.

The reference has no tests or golden vectors and cannot be built here (no OpenCV C++), so these
vectors are what pins the CPU oracle (oracle/mcm_oracle.cpp) to the arithmetic the reference
reaches through OpenCV:
    cv::Rodrigues        libs/multicam_mapper.cpp:470, 693, 910-911
    cv::Mat::inv()       :294, 539, 619, 621, 704
    cv::Mat operator*    :619-628, 640, 434, 704-705
    cv::undistortPoints  :570
Run once in the build container:  python tests/golden/make_golden_cv2.py
(the GPU box never runs this; it only reads the committed .npz).
"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "automatic-ar_b200", "python"))
from aar_b200 import synth  # noqa: E402


def rand_rigid(rng, n):
    T = np.zeros((n, 4, 4))
    for i in range(n):
        R, _ = cv2.Rodrigues(rng.normal(0, 1.5, 3))
        T[i, :3, :3] = R; T[i, :3, 3] = rng.normal(0, 1.0, 3); T[i, 3, 3] = 1
    return T


def gemm(A, B):
    return cv2.gemm(np.ascontiguousarray(A), np.ascontiguousarray(B), 1.0, None, 0.0)


def twin_project(Tc, To, Tm, K, X, cam_is_root, marker_is_root):
    """MultiCamMapper::project_marker (multicam_mapper.cpp:608-649) with cv2 calls."""
    T = To.copy()
    if not cam_is_root:
        T = gemm(cv2.invert(Tc)[1], T)
    if not marker_is_root:
        T = gemm(T, Tm)
    P = gemm(gemm(K, T[0:3, :]), X)
    return (P[0] / P[2]).astype(np.float32), (P[1] / P[2]).astype(np.float32)


def vec2mat(v):
    T = np.eye(4)
    T[:3, :3] = cv2.Rodrigues(np.ascontiguousarray(v[:3]).reshape(3, 1))[0]
    T[:3, 3] = v[3:6]
    return T


def main():
    rng = np.random.default_rng(20260001)
    out = {}
    # --- Rodrigues
    rv = np.concatenate([rng.normal(0, 1.2, (1500, 3)), rng.normal(0, 1e-3, (300, 3)), rng.normal(0, 3.0, (200, 3)),
                         np.array([[0, 0, 0], [1e-17, 0, 0], [2e-16, 0, 0], [3e-16, 0, 0], [np.pi, 0, 0]])])
    out["rod_r"] = rv
    out["rod_R"] = np.stack([cv2.Rodrigues(r.reshape(3, 1))[0] for r in rv])
    # --- R -> r (host-only; compared with a tolerance)
    Rm = out["rod_R"][:500].astype(np.float32).astype(np.float64)  # float32-rounded like initializer output
    out["rodinv_R"] = Rm
    out["rodinv_r"] = np.stack([cv2.Rodrigues(R)[0].reshape(3) for R in Rm])
    # --- inv 4x4
    T = rand_rigid(rng, 1000)
    out["inv_A"] = T
    out["inv_B"] = np.stack([cv2.invert(t)[1] for t in T])
    # --- gemm shapes on the path
    A = rand_rigid(rng, 500); B = rand_rigid(rng, 500)
    out["mm_A"] = A; out["mm_B"] = B
    out["mm_C"] = np.stack([gemm(a, b) for a, b in zip(A, B)])
    # --- undistortPoints
    K = np.array([[1003.7, 0, 641.3], [0, 998.2, 358.9], [0, 0, 1]])
    d = np.array([-0.1, 0.05, 1e-3, -2e-3, 0.01])
    pts = np.stack([rng.uniform(0, 1280, 3000), rng.uniform(0, 720, 3000)], axis=1).astype(np.float32)
    out["und_K"] = K; out["und_d"] = d; out["und_in"] = pts
    out["und_out"] = cv2.undistortPoints(pts.reshape(-1, 1, 2), K, d, None, K).reshape(-1, 2)
    out["und_out_zero"] = cv2.undistortPoints(pts.reshape(-1, 1, 2), K, np.zeros(5), None, K).reshape(-1, 2)

    # --- cv2 twin of error_function and jacobian_function on a small rig (zero distortion)
    rig = synth.make_rig(C=3, M=5, F=12, obs_per_frame=7.0, seed=11)
    for k, v in rig.__dict__.items():
        out["rig_" + k] = np.asarray(v)
    C, M, F = rig.C, rig.M, rig.F
    # z from cv2.Rodrigues(R->r) on the initial matrices, in the reference's packing order
    def mat2vec(Tm_):
        r = cv2.Rodrigues(np.ascontiguousarray(Tm_[:3, :3]))[0].reshape(3)
        return np.concatenate([r, Tm_[:3, 3]])
    z = np.concatenate([mat2vec(rig.T_cam_init[i]) for i in range(1, C)] + [mat2vec(rig.T_marker_init[i]) for i in range(1, M)] +
                       [mat2vec(rig.T_frame_init[i]) for i in range(F)])
    out["twin_z"] = z
    h = np.float32(rig.marker_size) / np.float32(2)
    X = np.array([[-h, h, 0, 1], [h, h, 0, 1], [h, -h, 0, 1], [-h, -h, 0, 1]], dtype=np.float64).T.copy()
    cam_index = {int(c): i for i, c in enumerate(rig.cam_ids)}
    marker_index = {int(m): i for i, m in enumerate(rig.marker_ids)}
    frame_index = {int(f): i for i, f in enumerate(rig.frame_ids)}

    def unpack(zv):
        Tc = [np.eye(4)] + [vec2mat(zv[6 * i:6 * i + 6]) for i in range(C - 1)]
        o = 6 * (C - 1)
        Tm_ = [np.eye(4)] + [vec2mat(zv[o + 6 * i:o + 6 * i + 6]) for i in range(M - 1)]
        o += 6 * (M - 1)
        To = [vec2mat(zv[o + 6 * i:o + 6 * i + 6]) for i in range(F)]
        return Tc, Tm_, To

    # zero distortion: cv2.undistortPoints still runs (remove_distortions, :554-578)
    und = np.zeros_like(rig.det_xy)
    for i in range(rig.N):
        ci = cam_index[int(rig.det_cam[i])]
        und[i] = cv2.undistortPoints(rig.det_xy[i].reshape(-1, 1, 2), rig.K[ci], rig.dist[ci], None, rig.K[ci]).reshape(-1)
    out["twin_und"] = und

    def residual(zv, override=None):
        Tc, Tm_, To = unpack(zv)
        if override is not None:
            kind, idx, Tnew = override
            {"c": Tc, "m": Tm_, "f": To}[kind][idx] = Tnew
        r = np.zeros(8 * rig.N)
        for i in range(rig.N):  # det order is already (frame, cam, detection order)
            ci = cam_index[int(rig.det_cam[i])]; mi = marker_index[int(rig.det_marker[i])]; fi = frame_index[int(rig.det_frame[i])]
            px, py = twin_project(Tc[ci], To[fi], Tm_[mi], rig.K[ci], X, ci == 0, mi == 0)
            r[8 * i + 0:8 * i + 8:2] = (und[i, 0::2] - px).astype(np.float64)   # float - float
            r[8 * i + 1:8 * i + 8:2] = (und[i, 1::2] - py).astype(np.float64)
        return r

    out["twin_r"] = residual(z)

    # Jacobian twin: dense (rows x n) from the quantised central difference with the RAW corners
    n = len(z); delta = 0.001
    J = np.zeros((8 * rig.N, n))
    raw = rig.det_xy
    def errs(zv, override):
        Tc, Tm_, To = unpack(zv)
        kind, idx, Tnew = override
        {"c": Tc, "m": Tm_, "f": To}[kind][idx] = Tnew
        e = np.zeros(8 * rig.N)
        for i in range(rig.N):
            ci = cam_index[int(rig.det_cam[i])]; mi = marker_index[int(rig.det_marker[i])]; fi = frame_index[int(rig.det_frame[i])]
            if (kind == "c" and ci != idx) or (kind == "m" and mi != idx) or (kind == "f" and fi != idx):
                continue
            px, py = twin_project(Tc[ci], To[fi], Tm_[mi], rig.K[ci], X, ci == 0, mi == 0)
            e[8 * i + 0:8 * i + 8:2] = (raw[i, 0::2] - px).astype(np.float64)
            e[8 * i + 1:8 * i + 8:2] = (raw[i, 1::2] - py).astype(np.float64)
        return e
    blocks = [("c", i + 1, 6 * i) for i in range(C - 1)] + [("m", i + 1, 6 * (C - 1) + 6 * i) for i in range(M - 1)] + \
             [("f", i, 6 * (C - 1) + 6 * (M - 1) + 6 * i) for i in range(F)]
    for kind, idx, col in blocks:
        v = z[col:col + 6]
        T0 = vec2mat(v)
        for dof in range(6):
            Ta = T0.copy(); Ts = T0.copy()
            if dof < 3:
                ra = v[:3].copy(); rs = v[:3].copy(); ra[dof] += delta; rs[dof] -= delta
                Ta[:3, :3] = cv2.Rodrigues(ra.reshape(3, 1))[0]; Ts[:3, :3] = cv2.Rodrigues(rs.reshape(3, 1))[0]
            else:
                Ta[dof - 3, 3] += delta; Ts[dof - 3, 3] -= delta
            J[:, col + dof] = (errs(z, (kind, idx, Ta)) - errs(z, (kind, idx, Ts))) / (2 * delta)
    out["twin_J"] = J
    np.savez_compressed(os.path.join(HERE, "cv2_golden.npz"), **out)
    print("wrote cv2_golden.npz:", {k: v.shape for k, v in out.items() if not k.startswith("rig_")})


if __name__ == "__main__":
    main()
