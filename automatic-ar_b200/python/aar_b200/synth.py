"""Deterministic synthetic rigs in the reference's dataset conventions.

There is no network, so the reference's sample datasets ("box", "pentagonal", README.md:55-56 of the
reference) are replaced by seeded generators that emit exactly what the hot path consumes:

* the `Initializer` output the 8-argument `MultiCamMapper` constructor takes
  (libs/multicam_mapper.h:17): root ids, `map<int,4x4>` of camera->root-camera, marker->root-marker
  and object(root marker)->root-camera transforms (float32-rounded entries like the IPPE output,
  libs/initializer.cpp:403-409), camera matrices and distortion coefficients, and the
  `frame -> cam -> [aruco::Marker]` detections;
* optionally a dataset folder (`<cam>/calib.yml` + `aruco.detections`).

Everything is numpy; generation is chunked over frames so the 100k-frame config fits in RAM.
"""
from __future__ import annotations

import dataclasses
import os
import struct

import numpy as np

# BASELINE.json configs: (cameras, markers, frames, target marker observations per frame)
CONFIGS = {
    "cfg1": dict(C=3, M=6, F=200, obs_per_frame=6.0),
    "cfg2": dict(C=5, M=12, F=1000, obs_per_frame=15.0),
    "cfg3": dict(C=8, M=24, F=10000, obs_per_frame=50.0),
    "cfg4": dict(C=16, M=64, F=100000, obs_per_frame=256.0),
    "cfg5": dict(C=16, M=64, F=100000, obs_per_frame=256.0),  # track mode: same density, frames independent
}


@dataclasses.dataclass
class Rig:
    cam_ids: np.ndarray        # int32 [C], ascending
    marker_ids: np.ndarray     # int32 [M], ascending
    frame_ids: np.ndarray      # int32 [F], ascending (frames with < 2 detections dropped)
    K: np.ndarray              # float64 [C,3,3]
    dist: np.ndarray           # float64 [C,5]  k1 k2 p1 p2 k3
    image_size: tuple
    marker_size: np.float32
    root_cam: int
    root_marker: int
    T_cam_true: np.ndarray     # [C,4,4] camera -> root camera
    T_marker_true: np.ndarray  # [M,4,4] marker -> root marker
    T_frame_true: np.ndarray   # [F,4,4] root marker -> root camera
    T_cam_init: np.ndarray
    T_marker_init: np.ndarray
    T_frame_init: np.ndarray
    det_frame: np.ndarray      # int32 [N] frame id, file order (frame, cam, detection order)
    det_cam: np.ndarray        # int32 [N]
    det_marker: np.ndarray     # int32 [N]
    det_xy: np.ndarray         # float32 [N,8] x0 y0 x1 y1 x2 y2 x3 y3 (raw, i.e. distorted, pixels)

    @property
    def C(self): return len(self.cam_ids)
    @property
    def M(self): return len(self.marker_ids)
    @property
    def F(self): return len(self.frame_ids)
    @property
    def N(self): return len(self.det_frame)


def rodrigues(r):
    """(...,3) rotation vectors -> (...,3,3); plain numpy, generator use only."""
    r = np.asarray(r, dtype=np.float64)
    th = np.linalg.norm(r, axis=-1)[..., None, None]
    k = r / np.maximum(np.linalg.norm(r, axis=-1, keepdims=True), 1e-300)
    Kx = np.zeros(r.shape[:-1] + (3, 3))
    Kx[..., 0, 1] = -k[..., 2]; Kx[..., 0, 2] = k[..., 1]
    Kx[..., 1, 0] = k[..., 2]; Kx[..., 1, 2] = -k[..., 0]
    Kx[..., 2, 0] = -k[..., 1]; Kx[..., 2, 1] = k[..., 0]
    I = np.broadcast_to(np.eye(3), Kx.shape)
    kkT = k[..., :, None] * k[..., None, :]
    return np.cos(th) * I + (1 - np.cos(th)) * kkT + np.sin(th) * Kx


def se3(R, t):
    T = np.zeros(R.shape[:-2] + (4, 4))
    T[..., :3, :3] = R
    T[..., :3, 3] = t
    T[..., 3, 3] = 1.0
    return T


def se3_inv(T):
    R = T[..., :3, :3]; t = T[..., :3, 3]
    Rt = np.swapaxes(R, -1, -2)
    return se3(Rt, -(Rt @ t[..., None])[..., 0])


def _look_at(eye, target=np.zeros(3), up=np.array([0.0, 0.0, 1.0])):
    """camera->world pose with +z looking from eye to target (OpenCV camera axes: x right, y down)."""
    z = target - eye; z /= np.linalg.norm(z)
    x = np.cross(z, up); x /= np.linalg.norm(x)
    y = np.cross(z, x)
    return se3(np.stack([x, y, z], axis=1), eye)


def _fibonacci_sphere(n):
    i = np.arange(n) + 0.5
    phi = np.arccos(1 - 2 * i / n)
    th = np.pi * (1 + 5 ** 0.5) * i
    return np.stack([np.cos(th) * np.sin(phi), np.sin(th) * np.sin(phi), np.cos(phi)], axis=1)


def _perturb(T, rng, rot_sigma, trans_sigma):
    n = T.shape[0]
    dT = se3(rodrigues(rng.normal(0, rot_sigma, (n, 3))), rng.normal(0, trans_sigma, (n, 3)))
    out = T @ dT
    out = out.astype(np.float32).astype(np.float64)  # IPPE matrices are CV_32F (initializer.cpp:403-409)
    out[:, 3, :] = [0, 0, 0, 1]
    return out


def _distort(xn, yn, d):
    k1, k2, p1, p2, k3 = d
    r2 = xn * xn + yn * yn
    rad = 1 + ((k3 * r2 + k2) * r2 + k1) * r2
    xd = xn * rad + 2 * p1 * xn * yn + p2 * (r2 + 2 * xn * xn)
    yd = yn * rad + p1 * (r2 + 2 * yn * yn) + 2 * p2 * xn * yn
    return xd, yd


def make_rig(C, M, F, obs_per_frame, seed=0, marker_size=0.05, noise_px=0.3, distorted=False,
             rot_sigma=0.02, trans_sigma=0.005, shuffle_detections=True, sparse_marker_ids=True,
             chunk_frames=4000) -> Rig:
    rng = np.random.default_rng(seed)
    W, H = 1280, 720
    # cameras on a ring of radius 1.5 m looking at the origin (world frame), slightly varied height
    ang = 2 * np.pi * np.arange(C) / C + rng.normal(0, 0.05, C)
    eyes = np.stack([1.5 * np.cos(ang), 1.5 * np.sin(ang), rng.uniform(-0.3, 0.5, C)], axis=1)
    T_cw = np.stack([_look_at(e) for e in eyes])            # camera -> world
    T_cam_true = se3_inv(T_cw[0:1]) @ T_cw                  # camera -> root camera (cam 0 is root: identity)
    T_cam_true[0] = np.eye(4)
    f = 1000.0 * (1 + rng.uniform(-0.02, 0.02, C))
    K = np.zeros((C, 3, 3)); K[:, 0, 0] = f; K[:, 1, 1] = f; K[:, 0, 2] = 640.0; K[:, 1, 2] = 360.0; K[:, 2, 2] = 1.0
    dist = np.zeros((C, 5))
    if distorted:
        dist[:] = [-0.1, 0.05, 1e-3, -2e-3, 0.01]
        dist += rng.normal(0, 1e-3, dist.shape) * np.array([10, 5, 0.1, 0.1, 1])

    # markers on a polyhedron of circumradius 0.15 m, outward-facing (object frame)
    nrm = _fibonacci_sphere(M)
    Tm_obj = np.zeros((M, 4, 4))
    for i in range(M):
        z = nrm[i]
        a = np.array([0.0, 0.0, 1.0]) if abs(z[2]) < 0.9 else np.array([1.0, 0.0, 0.0])
        x = np.cross(a, z); x /= np.linalg.norm(x)
        spin = rng.uniform(0, 2 * np.pi)
        y = np.cross(z, x)
        x, y = np.cos(spin) * x + np.sin(spin) * y, -np.sin(spin) * x + np.cos(spin) * y
        Tm_obj[i] = se3(np.stack([x, y, z], axis=1), 0.15 * z)
    T_marker_true = se3_inv(Tm_obj[0:1]) @ Tm_obj            # marker -> root marker (marker 0 is root)
    T_marker_true[0] = np.eye(4)
    marker_ids = (np.arange(M) * 3 + 7 if sparse_marker_ids else np.arange(M)).astype(np.int32)
    cam_ids = np.arange(C, dtype=np.int32)

    h = np.float32(marker_size) / np.float32(2)
    Xm = np.array([[-h, h, 0, 1], [h, h, 0, 1], [h, -h, 0, 1], [-h, -h, 0, 1]], dtype=np.float64).T  # 4x4, columns=corners

    # smooth object trajectory (world frame): low-frequency sinusoids inside a 0.5 m cube, slow tumbling
    tt = np.arange(F) / max(F, 1)
    ph = rng.uniform(0, 2 * np.pi, (3, 3)); fr = rng.uniform(0.5, 3.0, (3, 3))
    pos = 0.25 * np.stack([np.sin(2 * np.pi * fr[0, i] * tt * (1 + F / 2000.0) + ph[0, i]) for i in range(3)], axis=1) * 0.8
    rv = np.stack([1.2 * np.sin(2 * np.pi * fr[1, i] * tt * (1 + F / 3000.0) + ph[1, i]) + 0.8 * np.sin(2 * np.pi * fr[2, i] * tt * 7 + ph[2, i]) for i in range(3)], axis=1)
    T_ow = se3(rodrigues(rv), pos)                           # object -> world
    # root marker -> root camera = inv(T_cw0) * T_ow * Tm_obj0
    T_frame_all = se3_inv(T_cw[0]) @ T_ow @ Tm_obj[0]

    det_f, det_c, det_m, det_xy = [], [], [], []
    # estimate visibility rate on the first chunk to set the Bernoulli thinning probability
    p_keep = None
    Tci = se3_inv(T_cam_true)                                # root camera -> camera
    Yall = np.concatenate([T_marker_true[m] @ Xm for m in range(M)], axis=1)                      # (4, 4M) corners in the root-marker frame
    NCm = np.stack([np.stack([T_marker_true[m][:, 2], T_marker_true[m][:, 3]], axis=1) for m in range(M)], axis=1).reshape(4, 2 * M)
    for f0 in range(0, F, chunk_frames):
        f1 = min(F, f0 + chunk_frames)
        n = f1 - f0
        # T1[f,c] = inv(Tc) * To ; corners and normals of every marker in the object frame are fixed, so
        # the camera-frame corners are one (n*C*3, 4) x (4, 4M) product
        T1 = np.einsum("cij,fjk->fcik", Tci, T_frame_all[f0:f1])           # (n,C,4,4)
        P = (T1[:, :, :3, :].reshape(n * C * 3, 4) @ Yall).reshape(n, C, 3, M, 4).transpose(0, 1, 3, 2, 4)  # (n,C,M,3,4)
        z = P[..., 2, :]
        xn = P[..., 0, :] / z; yn = P[..., 1, :] / z
        if distorted:
            xd = np.empty_like(xn); yd = np.empty_like(yn)
            for c in range(C):
                xd[:, c], yd[:, c] = _distort(xn[:, c], yn[:, c], dist[c])
        else:
            xd, yd = xn, yn
        u = xd * K[None, :, None, None, 0, 0] + K[None, :, None, None, 0, 2]
        v = yd * K[None, :, None, None, 1, 1] + K[None, :, None, None, 1, 2]
        NC = (T1[:, :, :3, :].reshape(n * C * 3, 4) @ NCm).reshape(n, C, 3, M, 2)      # normal (w=0) and centre (w=1)
        normal = NC[..., 0].transpose(0, 1, 3, 2); centre = NC[..., 1].transpose(0, 1, 3, 2)
        view = centre / np.linalg.norm(centre, axis=-1, keepdims=True)
        facing = (normal * view).sum(-1) < -0.2
        inside = ((u > 1) & (u < W - 2) & (v > 1) & (v < H - 2) & (z > 0.1)).all(-1)
        vis = facing & inside
        if p_keep is None:
            rate = vis.sum() / max(n, 1)
            p_keep = min(1.0, obs_per_frame / max(rate, 1e-9))
        keep = vis & (rng.random(vis.shape) < p_keep)
        fi, ci, mi = np.nonzero(keep)                        # sorted by (frame, cam, marker)
        if shuffle_detections and len(fi):
            # random detection order inside every (frame, cam) group
            key = rng.random(len(fi))
            order = np.lexsort((key, ci, fi))
            fi, ci, mi = fi[order], ci[order], mi[order]
        xy = np.stack([u[fi, ci, mi], v[fi, ci, mi]], axis=-1)            # (n_obs,4,2)
        xy = xy + rng.normal(0, noise_px, xy.shape)
        det_f.append((fi + f0).astype(np.int32)); det_c.append(cam_ids[ci]); det_m.append(marker_ids[mi])
        det_xy.append(xy.reshape(-1, 8).astype(np.float32))
    det_f = np.concatenate(det_f); det_c = np.concatenate(det_c); det_m = np.concatenate(det_m); det_xy = np.concatenate(det_xy)
    # frames with fewer than 2 detections are dropped (initializer.h:53, initializer.cpp:379)
    cnt = np.bincount(det_f, minlength=F)
    ok = cnt >= 2
    sel = ok[det_f]
    det_f, det_c, det_m, det_xy = det_f[sel], det_c[sel], det_m[sel], det_xy[sel]
    frame_ids = np.nonzero(ok)[0].astype(np.int32)
    T_frame_true = T_frame_all[frame_ids]

    T_cam_init = _perturb(T_cam_true, rng, rot_sigma, trans_sigma); T_cam_init[0] = np.eye(4)
    T_marker_init = _perturb(T_marker_true, rng, rot_sigma, trans_sigma); T_marker_init[0] = np.eye(4)
    T_frame_init = _perturb(T_frame_true, rng, rot_sigma, trans_sigma)
    return Rig(cam_ids=cam_ids, marker_ids=marker_ids, frame_ids=frame_ids, K=K, dist=dist, image_size=(W, H),
               marker_size=np.float32(marker_size), root_cam=int(cam_ids[0]), root_marker=int(marker_ids[0]),
               T_cam_true=T_cam_true, T_marker_true=T_marker_true, T_frame_true=T_frame_true,
               T_cam_init=T_cam_init, T_marker_init=T_marker_init, T_frame_init=T_frame_init,
               det_frame=det_f, det_cam=det_c.astype(np.int32), det_marker=det_m.astype(np.int32), det_xy=det_xy)


def make_config(name: str, seed=None, frames=None, **kw) -> Rig:
    cfg = dict(CONFIGS[name])
    if frames is not None:
        cfg["F"] = frames
    if seed is None:
        seed = int(name[-1])
    cache = os.environ.get("AAR_RIG_CACHE")          # development aid: several processes of one profiling session share the rig
    path = os.path.join(cache, f"{name}_{cfg['F']}_{seed}.pkl") if cache and not kw else None
    if path and os.path.exists(path):
        import pickle
        with open(path, "rb") as fh:
            return pickle.load(fh)
    rig = make_rig(cfg["C"], cfg["M"], cfg["F"], cfg["obs_per_frame"], seed=seed, **kw)
    if path:
        import pickle
        os.makedirs(cache, exist_ok=True)
        with open(path, "wb") as fh:
            pickle.dump(rig, fh, protocol=4)
    return rig


# ------------------------------------------------------------------ dataset files (Appendix A.1/A.2)
def write_detections_file(path, rig: Rig):
    """`aruco.detections`: size_t num_cams, then per frame, per cam: size_t n, n x {int id, 8 float}
    (reader: libs/initializer.cpp:316-362; record: libs/aruco_serdes.cpp:9-24).  The frame index in
    the file is the ordinal, so every frame up to the last id is written (possibly empty)."""
    C = rig.C
    nF = int(rig.frame_ids.max()) + 1 if rig.F else 0
    order = np.lexsort((np.arange(rig.N), rig.det_cam, rig.det_frame))
    starts = np.searchsorted(rig.det_frame[order] * C + rig.det_cam[order], np.arange(nF * C + 1))
    with open(path, "wb") as fh:
        fh.write(struct.pack("<Q", C))
        for f in range(nF):
            for c in range(C):
                a, b = starts[f * C + c], starts[f * C + c + 1]
                fh.write(struct.pack("<Q", b - a))
                for k in order[a:b]:
                    fh.write(struct.pack("<i8f", int(rig.det_marker[k]), *[float(x) for x in rig.det_xy[k]]))


def write_calib_files(folder, rig: Rig):
    """<folder>/<cam>/calib.yml in OpenCV FileStorage YAML (reader: libs/cam_config.cpp:52-78)."""
    for i, cid in enumerate(rig.cam_ids):
        d = os.path.join(folder, str(int(cid)))
        os.makedirs(d, exist_ok=True)
        Kv = ", ".join(repr(float(x)) for x in rig.K[i].reshape(-1))
        dv = ", ".join(repr(float(x)) for x in rig.dist[i])
        with open(os.path.join(d, "calib.yml"), "w") as fh:
            fh.write("%YAML:1.0\n---\n")
            fh.write(f"image_width: {rig.image_size[0]}\nimage_height: {rig.image_size[1]}\n")
            fh.write(f"camera_matrix: !!opencv-matrix\n   rows: 3\n   cols: 3\n   dt: d\n   data: [ {Kv} ]\n")
            fh.write(f"distortion_coefficients: !!opencv-matrix\n   rows: 1\n   cols: 5\n   dt: d\n   data: [ {dv} ]\n")


def write_dataset(folder, rig: Rig):
    os.makedirs(folder, exist_ok=True)
    write_calib_files(folder, rig)
    write_detections_file(os.path.join(folder, "aruco.detections"), rig)


# ------------------------------------------------------------------ .solution files
def rot2vec(R):
    """(...,3,3) -> (...,3) rotation vectors (cv::Rodrigues matrix -> vector for proper rotations, angle < pi)."""
    R = np.asarray(R, dtype=np.float64)
    v = np.stack([R[..., 2, 1] - R[..., 1, 2], R[..., 0, 2] - R[..., 2, 0], R[..., 1, 0] - R[..., 0, 1]], axis=-1)
    s = np.linalg.norm(v, axis=-1) / 2
    c = np.clip((np.trace(R, axis1=-2, axis2=-1) - 1) / 2, -1, 1)
    th = np.arctan2(s, c)
    k = np.where(s > 1e-12, th / np.maximum(2 * s, 1e-300), 0.5)
    return v * k[..., None]


def full_vector(rig: Rig, Tc, Tm, Tf):
    """io_vec for the full Config: cams (root skipped), markers (root skipped), frames, per cam [fx cx fy cy k1 k2 p1 p2 k3]."""
    def six(T):
        return np.concatenate([rot2vec(T[:, :3, :3]), T[:, :3, 3]], axis=1).reshape(-1)
    ci = rig.cam_ids != rig.root_cam; mi = rig.marker_ids != rig.root_marker
    intr = np.concatenate([np.stack([rig.K[:, 0, 0], rig.K[:, 0, 2], rig.K[:, 1, 1], rig.K[:, 1, 2]], axis=1), rig.dist], axis=1).reshape(-1)
    return np.concatenate([six(Tc[ci]), six(Tm[mi]), six(Tf), intr])


def write_solution_file(path, rig: Rig, use_init=True, flags=(True, True, True, False)):
    """The reference's binary .solution (libs/multicam_mapper.cpp:1053-1099).  Corners must already be undistorted:
    only valid for rigs generated with distorted=False (undistortPoints with zero coefficients is the identity)."""
    assert not np.any(rig.dist), "write_solution_file needs undistorted corners"
    Tc, Tm, Tf = (rig.T_cam_init, rig.T_marker_init, rig.T_frame_init) if use_init else (rig.T_cam_true, rig.T_marker_true, rig.T_frame_true)
    with open(path, "wb") as fh:
        fh.write(struct.pack("<Q", rig.C)); fh.write(rig.cam_ids.astype("<i4").tobytes())
        fh.write(struct.pack("<Q", int(rig.root_cam)))
        for _ in range(rig.C):
            fh.write(struct.pack("<ii", int(rig.image_size[0]), int(rig.image_size[1])))
        fh.write(struct.pack("<Q", rig.M)); fh.write(rig.marker_ids.astype("<i4").tobytes())
        fh.write(struct.pack("<Q", int(rig.root_marker)))
        fh.write(struct.pack("<d", float(np.float32(rig.marker_size))))
        fh.write(struct.pack("<Q", rig.F)); fh.write(rig.frame_ids.astype("<i4").tobytes())
        fh.write(full_vector(rig, Tc, Tm, Tf).astype("<f8").tobytes())
        # frame_cam_markers: frames ascending, cams ascending, detection order
        order = np.lexsort((np.arange(rig.N), rig.det_cam, rig.det_frame))
        df, dc, dm, xy = rig.det_frame[order], rig.det_cam[order], rig.det_marker[order], rig.det_xy[order]
        frames, fstart = np.unique(df, return_index=True)
        fh.write(struct.pack("<Q", len(frames)))
        fend = list(fstart[1:]) + [len(df)]
        rec = np.zeros(len(df), dtype=[("id", "<i4"), ("xy", "<f4", 8)]); rec["id"] = dm; rec["xy"] = xy
        for f, a, b in zip(frames, fstart, fend):
            cams, cstart = np.unique(dc[a:b], return_index=True)
            fh.write(struct.pack("<iQ", int(f), len(cams)))
            cend = list(cstart[1:]) + [b - a]
            for c, ca, cb in zip(cams, cstart, cend):
                fh.write(struct.pack("<iQ", int(c), cb - ca))
                fh.write(rec[a + ca:a + cb].tobytes())
        fh.write(struct.pack("<4?", *flags))


def read_solution_file(path):
    """-> dict(cam_ids, root_cam, image_sizes, marker_ids, root_marker, marker_size, frame_ids, vec, det_frame, det_cam, det_marker, det_xy, flags)"""
    buf = open(path, "rb").read(); pos = 0

    def take(fmt):
        nonlocal pos
        v = struct.unpack_from(fmt, buf, pos); pos += struct.calcsize(fmt); return v
    C, = take("<Q"); cam_ids = np.frombuffer(buf, "<i4", C, pos).copy(); pos += 4 * C
    root_cam, = take("<Q"); sizes = [take("<ii") for _ in range(C)]
    M, = take("<Q"); marker_ids = np.frombuffer(buf, "<i4", M, pos).copy(); pos += 4 * M
    root_marker, = take("<Q"); marker_size, = take("<d")
    F, = take("<Q"); frame_ids = np.frombuffer(buf, "<i4", F, pos).copy(); pos += 4 * F
    n = 6 * (C - 1) + 6 * (M - 1) + 6 * F + 9 * C
    vec = np.frombuffer(buf, "<f8", n, pos).copy(); pos += 8 * n
    nf, = take("<Q"); df, dc, dm, xy = [], [], [], []
    for _ in range(nf):
        f, nc = take("<iQ")
        for _ in range(nc):
            c, nm = take("<iQ")
            rec = np.frombuffer(buf, [("id", "<i4"), ("xy", "<f4", 8)], nm, pos); pos += 36 * nm
            df += [f] * nm; dc += [c] * nm; dm += list(rec["id"]); xy.append(rec["xy"].copy())
    flags = take("<4?")
    assert pos == len(buf)
    return dict(cam_ids=cam_ids, root_cam=root_cam, image_sizes=sizes, marker_ids=marker_ids, root_marker=root_marker, marker_size=marker_size,
                frame_ids=frame_ids, vec=vec, det_frame=np.array(df, np.int32), det_cam=np.array(dc, np.int32), det_marker=np.array(dm, np.int32),
                det_xy=np.concatenate(xy) if xy else np.zeros((0, 8), np.float32), flags=flags)
