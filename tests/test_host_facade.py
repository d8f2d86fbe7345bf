"""The C++ host facade (automatic-ar_b200/host): file formats of the reference and, on the GPU,
the find_solution / track apps calling MultiCamMapper::solve() / track() through the C ABI."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from aar_b200 import synth

HOST = os.path.join(ROOT, "automatic-ar_b200", "host")


@pytest.fixture(scope="module")
def host_bins():
    from aar_b200 import binding
    binding.build()
    subprocess.run(["make", "-s", "-C", HOST], check=True)
    return HOST


def test_solution_file_round_trip(host_bins, tmp_path):
    rig = synth.make_rig(C=3, M=5, F=12, obs_per_frame=5.0, seed=4)
    a, b, y = str(tmp_path / "a.solution"), str(tmp_path / "b.solution"), str(tmp_path / "b.yaml")
    synth.write_solution_file(a, rig)
    subprocess.run([os.path.join(host_bins, "solution_tool"), "roundtrip", a, b, y], check=True)
    A, B = synth.read_solution_file(a), synth.read_solution_file(b)
    for k in ("cam_ids", "marker_ids", "frame_ids", "det_frame", "det_cam", "det_marker", "det_xy"):
        assert np.array_equal(A[k], B[k]), k                          # integer / byte content is exact
    assert A["root_cam"] == B["root_cam"] and A["root_marker"] == B["root_marker"] and A["marker_size"] == B["marker_size"]
    assert A["image_sizes"] == B["image_sizes"] and A["flags"] == B["flags"] and os.path.getsize(a) == os.path.getsize(b)
    assert np.abs(A["vec"] - B["vec"]).max() < 1e-12                  # r -> R -> r (cv::Rodrigues both ways)
    # the YAML export has the layout cv::FileStorage gives `fs << "{:" << "cam_id" << id << "transform" << Mat << "}"`
    # (multicam_mapper.cpp:1233-1268; compare tests/golden/make_golden_cv2.py) and carries the same matrices.
    # (cv2 4.13 cannot re-read a matrix inside a flow mapping — not even its own output — so the check parses the text.)
    import re
    txt = open(y).read()
    assert txt.startswith("%YAML:1.0\n---\nmarker_size: ")
    assert abs(float(txt.split("marker_size:")[1].split()[0]) - float(np.float32(rig.marker_size))) < 1e-15
    sec = txt.split("transforms_to_root_marker:")[1].split("root_marker_to_root_cam:")[0]
    ids = [int(v) for v in re.findall(r"marker_id:(-?\d+)", sec)]
    assert ids == list(rig.marker_ids)
    mats = re.findall(r"data: \[([^\]]*)\]", sec)
    T = np.array([float(v) for v in mats[2].replace("\n", " ").split(",")]).reshape(4, 4)
    assert np.abs(T - rig.T_marker_init[2]).max() < 1e-6      # the .solution stores rotation VECTORS: float32-rounded matrices come back orthonormalised
    assert len(re.findall(r"frame_id:", txt)) == rig.F and "!!opencv-matrix" in txt and "dt: d" in txt


def test_detections_file_round_trip(host_bins, tmp_path):
    rig = synth.make_rig(C=3, M=5, F=12, obs_per_frame=5.0, seed=5)
    a, b = str(tmp_path / "aruco.detections"), str(tmp_path / "b.detections")
    synth.write_detections_file(a, rig)
    subprocess.run([os.path.join(host_bins, "solution_tool"), "detections", a, b], check=True)
    assert open(a, "rb").read() == open(b, "rb").read()              # byte exact
    # a truncated last frame is discarded (initializer.cpp:335-347)
    raw = open(a, "rb").read()
    open(a, "wb").write(raw[:-20])
    out = subprocess.run([os.path.join(host_bins, "solution_tool"), "detections", a, b], check=True, capture_output=True, text=True).stdout
    assert int(out.split()[0]) == int(rig.frame_ids.max())            # one frame fewer than written (ids 0..max)


def test_stereo_calib_and_ground_truth_files(host_bins, tmp_path):
    """read/write_stereo_calib and read/write_ground_truth (multicam_mapper.cpp:86-184): records of (rotation vector, translation) as
    the reference lays them out; matrices against cv2.Rodrigues, the added inverse edges against cv2.invert (cv::Mat::inv, LU)."""
    import struct
    import cv2
    rng = np.random.default_rng(12)
    edges = {0: {1: rng.normal(0, 0.7, 6), 2: rng.normal(0, 0.7, 6)}, 5: {3: rng.normal(0, 0.7, 6)}}
    a, b = str(tmp_path / "stereo.calib"), str(tmp_path / "stereo2.calib")
    with open(a, "wb") as fh:
        fh.write(struct.pack("<Q", len(edges)))
        for n1, sec in edges.items():
            fh.write(struct.pack("<iQ", n1, len(sec)))
            for n2, v in sec.items():
                fh.write(struct.pack("<i6d", n2, *v))
        fh.write(struct.pack("<i", 5))
    out = subprocess.run([os.path.join(host_bins, "solution_tool"), "stereo", a, b], check=True, capture_output=True, text=True).stdout.strip().splitlines()
    assert out[0] == "root 5"
    got = {(int(l.split()[0]), int(l.split()[1])): np.array([float(x) for x in l.split()[2:]]).reshape(4, 4) for l in out[1:]}
    assert set(got) == {(0, 1), (1, 0), (0, 2), (2, 0), (5, 3), (3, 5)}
    for n1, sec in edges.items():
        for n2, v in sec.items():
            T = np.eye(4); T[:3, :3] = cv2.Rodrigues(v[:3].reshape(3, 1))[0]; T[:3, 3] = v[3:]
            assert np.abs(got[(n1, n2)] - T).max() <= 2.3e-16                  # shared sincos: within one ulp of libm's (tests/test_oracle_pin.py)
            assert np.abs(got[(n2, n1)] - cv2.invert(got[(n1, n2)])[1]).max() <= 1e-15
    # written back: same structure, every stored edge and its inverse, ids ascending like std::map; vectors to R -> r accuracy
    raw = open(b, "rb").read()
    (n1,) = struct.unpack_from("<Q", raw, 0); off = 8; seen = {}
    for _ in range(n1):
        node1, n2 = struct.unpack_from("<iQ", raw, off); off += 12
        for _ in range(n2):
            rec = struct.unpack_from("<i6d", raw, off); off += 52
            seen[(node1, rec[0])] = np.array(rec[1:])
    assert struct.unpack_from("<i", raw, off)[0] == 5 and off + 4 == len(raw)
    assert list(seen) == sorted(seen) and set(seen) == set(got)
    for (n1_, n2_), v in seen.items():
        assert np.abs(cv2.Rodrigues(v[:3].reshape(3, 1))[0] - got[(n1_, n2_)][:3, :3]).max() <= 1e-12 and np.array_equal(v[3:], got[(n1_, n2_)][:3, 3])
    # ground truth: frame number + pose records until EOF; a record cut short is an error like in the reference
    poses = {3: rng.normal(0, 0.5, 6), 11: rng.normal(0, 0.5, 6), 4: rng.normal(0, 0.5, 6)}
    g, g2 = str(tmp_path / "truth.bin"), str(tmp_path / "truth2.bin")
    with open(g, "wb") as fh:
        for f, v in poses.items():
            fh.write(struct.pack("<Q6d", f, *v))
    out = subprocess.run([os.path.join(host_bins, "solution_tool"), "truth", g, g2], check=True, capture_output=True, text=True).stdout.strip().splitlines()
    assert [int(l.split()[0]) for l in out] == [3, 4, 11]
    for l in out:
        f = int(l.split()[0]); T = np.array([float(x) for x in l.split()[1:]]).reshape(4, 4); v = poses[f]
        assert np.abs(T[:3, :3] - cv2.Rodrigues(v[:3].reshape(3, 1))[0]).max() <= 2.3e-16 and np.array_equal(T[:3, 3], v[3:]) and np.array_equal(T[3], [0, 0, 0, 1])
    raw2 = open(g2, "rb").read()
    assert len(raw2) == 3 * 56
    for i, f in enumerate([3, 4, 11]):
        rec = struct.unpack_from("<Q6d", raw2, 56 * i)
        assert rec[0] == f and np.abs(np.array(rec[1:]) - poses[f]).max() <= 1e-12
    cut = open(g, "rb").read()[:-10]
    open(g, "wb").write(cut)
    r = subprocess.run([os.path.join(host_bins, "solution_tool"), "truth", g, g2], capture_output=True, text=True)
    assert r.returncode == 2 and "Unexpected end of input ground truth file" in r.stderr


def test_calib_reader_on_cv2_written_files(host_bins, tmp_path):
    """CamConfig::read_cam_configs (libs/cam_config.cpp:52-95) on calib.yml files written by OpenCV's own cv::FileStorage (cv2 4.13),
    the writer the reference's datasets come from: every value must come back exactly; folders are taken in numeric order."""
    import cv2
    rng = np.random.default_rng(7)
    want = {}
    for cam in (0, 1, 2, 10):                                   # "10" sorts before "2" as a string: numeric order is what indexes cam_configs[cam_id]
        K = np.array([[1000 * (1 + rng.normal(0, 0.02)), 0, 640 + rng.normal(0, 3)], [0, 1000 * (1 + rng.normal(0, 0.02)), 360 + rng.normal(0, 3)], [0, 0, 1]])
        d = rng.normal(0, 0.05, (1, 5)) if cam != 2 else rng.normal(0, 0.05, (1, 4))      # a 4-coefficient file: zero padded (setDistCoeffs)
        os.makedirs(tmp_path / str(cam))
        fs = cv2.FileStorage(str(tmp_path / str(cam) / "calib.yml"), cv2.FILE_STORAGE_WRITE)
        fs.write("image_width", 1280 + cam); fs.write("image_height", 720); fs.write("camera_matrix", K); fs.write("distortion_coefficients", d)
        fs.release()
        want[cam] = (1280 + cam, 720, K.reshape(-1), np.concatenate([d.reshape(-1), np.zeros(5 - d.size)]))
    out = subprocess.run([os.path.join(host_bins, "solution_tool"), "calib", str(tmp_path), "-"], check=True, capture_output=True, text=True).stdout
    rows = [l.split() for l in out.strip().splitlines()]
    assert len(rows) == 4
    for row, cam in zip(rows, sorted(want)):
        w, h, K, d = want[cam]
        assert int(row[0]) == w and int(row[1]) == h
        assert np.array_equal(np.array([float(v) for v in row[2:11]]), K) and np.array_equal(np.array([float(v) for v in row[11:16]]), d)
    # and the files synth.py writes for the datasets of the other tests parse to the rig's values
    rig = synth.make_rig(C=2, M=3, F=4, obs_per_frame=3.0, seed=6, distorted=True)
    synth.write_calib_files(str(tmp_path / "ds"), rig)
    out = subprocess.run([os.path.join(host_bins, "solution_tool"), "calib", str(tmp_path / "ds"), "-"], check=True, capture_output=True, text=True).stdout
    rows = [l.split() for l in out.strip().splitlines()]
    for i, row in enumerate(rows):
        assert np.array_equal(np.array([float(v) for v in row[2:11]]), rig.K[i].reshape(-1)) and np.array_equal(np.array([float(v) for v in row[11:16]]), rig.dist[i])


@pytest.mark.gpu
def test_find_solution_app_matches_binding(host_bins, tmp_path):
    from aar_b200 import binding
    rig = synth.make_config("cfg1")
    synth.write_solution_file(str(tmp_path / "initial.solution"), rig)
    out = subprocess.run([os.path.join(host_bins, "find_solution"), str(tmp_path), "0.05", "-init", str(tmp_path / "initial.solution")], check=True, capture_output=True, text=True).stdout
    assert "The algorithm took:" in out
    fin = synth.read_solution_file(str(tmp_path / "final.solution"))
    p = binding.Problem(rig)
    z, fc, it, tr = p.solve(p.mats2evec())
    app_cost = float(out.split("final_error:")[1].split()[0])
    assert abs(app_cost - fc) <= 2e-5 * fc                            # same solve up to the reproducibility envelope
    n = p.num_vars
    # compare poses, not rotation vectors: the file holds cv::Rodrigues(R) of the final matrices, which maps a vector
    # whose angle went past pi during the optimisation back to its equivalent below pi
    a, b = fin["vec"][:n].reshape(-1, 6), z.reshape(-1, 6)
    assert np.abs(synth.rodrigues(a[:, :3]) - synth.rodrigues(b[:, :3])).max() <= 5e-5
    assert np.abs(a[:, 3:] - b[:, 3:]).max() <= 5e-5
    assert fin["flags"] == (True, True, True, False)
    assert os.path.exists(tmp_path / "final.solution.yaml")


@pytest.mark.gpu
def test_find_solution_app_analytic_option(host_bins, tmp_path):
    """find_solution -analytic: MultiCamMapper::set_analytic_jacobian through the facade (include/aar_analytic.h) gives the solve of the
    binding's analytic variant (no float32 quantisation in this variant: the two agree far inside the faithful path's envelope)."""
    from aar_b200 import binding
    rig = synth.make_config("cfg1")
    synth.write_solution_file(str(tmp_path / "initial.solution"), rig)
    out = subprocess.run([os.path.join(host_bins, "find_solution"), str(tmp_path), "0.05", "-init", str(tmp_path / "initial.solution"), "-analytic"], check=True, capture_output=True, text=True).stdout
    app_cost = float(out.split("final_error:")[1].split()[0])
    p = binding.Problem(rig, analytic=True)
    z, fc, it, tr = p.solve(p.mats2evec())
    assert int(out.split("iterations:")[1].split()[0]) == it
    assert abs(app_cost - fc) <= 1e-9 * fc
    pf = binding.Problem(rig)
    _, fc_f, _, _ = pf.solve(pf.mats2evec())
    assert app_cost != fc_f and abs(app_cost - fc_f) <= 1e-4 * fc_f       # it really is the other variant, next to the same optimum


@pytest.mark.gpu
@pytest.mark.parametrize("workload", ["cfg1", "distorted"])
def test_find_solution_from_detections_and_calib(host_bins, oracle_mod, tmp_path, workload):
    """apps/find_solution.cpp:102-177 end to end: <cam>/calib.yml + aruco.detections -> Initializer -> MultiCamMapper::solve ->
    final.solution, against the CPU oracle pipeline (restated Initializer -> restated MultiCamMapper driving sparselevmarq.h)."""
    import copy
    from conftest import parity_record
    # "distorted": the Jacobian must difference the RAW corners while the residual uses the undistorted ones (the facade's
    # init -> set flags -> solve flow used to rebuild its handle from the undistorted corners)
    rig = synth.make_config("cfg1") if workload == "cfg1" else synth.make_rig(C=3, M=6, F=150, obs_per_frame=7.0, seed=31, distorted=True)
    synth.write_dataset(str(tmp_path), rig)
    out = subprocess.run([os.path.join(host_bins, "find_solution"), str(tmp_path), "0.05"], check=True, capture_output=True, text=True).stdout
    assert "The algorithm took:" in out and os.path.exists(tmp_path / "initial.solution") and os.path.exists(tmp_path / "final.solution.yaml")
    # the oracle pipeline on the same files' content
    nF = int(rig.frame_ids.max()) + 1
    io = oracle_mod.InitOracle(rig.C, rig.K, rig.dist, 0.05, nF, rig.det_frame, rig.det_cam, rig.det_marker, rig.det_xy)
    io.obtain_pose_estimations(); io.init_transforms()
    r = io.results()
    ini = synth.read_solution_file(str(tmp_path / "initial.solution"))
    assert np.array_equal(ini["cam_ids"], r["cams"][0]) and np.array_equal(ini["marker_ids"], r["markers"][0]) and np.array_equal(ini["frame_ids"], r["objects"][0])
    assert ini["root_cam"] == r["root_cam"] and ini["root_marker"] == r["root_marker"]
    rig_o = copy.copy(rig)
    rig_o.T_cam_init, rig_o.T_marker_init, rig_o.T_frame_init = r["cams"][1], r["markers"][1], r["objects"][1]
    o = oracle_mod.Oracle(rig_o)
    z0 = o.mats2evec()
    # the initial.solution holds the same starting point (as rotation vectors)
    n = len(z0)
    a, b = ini["vec"][:n].reshape(-1, 6), z0.reshape(-1, 6)
    d_init = max(np.abs(synth.rodrigues(a[:, :3]) - synth.rodrigues(b[:, :3])).max(), np.abs(a[:, 3:] - b[:, 3:]).max())
    assert d_init <= 1e-9, d_init
    zo, co, ito, tro = o.solve(z0)
    app_cost = float(out.split("final_error:")[1].split()[0]); app_it = int(out.split("iterations:")[1].split()[0])
    rel = abs(app_cost - co) / co
    parity_record("find_solution_app_from_detections_vs_oracle_pipeline", workload=workload, initial_solution_max_abs_dev=float(d_init), final_cost_app=app_cost,
                  final_cost_oracle=float(co), rel_dev_final_cost=float(rel), iterations=[app_it, int(ito)], bar=2e-5)
    assert rel <= 2e-5, (app_cost, co)                        # the reference's own reproducibility envelope


@pytest.mark.gpu
def test_track_app_from_detections(host_bins, tmp_path):
    """apps/track.cpp:85-133: a solved rig + the detections of new frames -> per-frame IPPE + object-pose consensus -> track()."""
    import copy
    rig = copy.copy(synth.make_rig(C=4, M=8, F=30, obs_per_frame=10.0, seed=18))
    rig.T_cam_init, rig.T_marker_init = rig.T_cam_true, rig.T_marker_true
    a, d, b = str(tmp_path / "rig.solution"), str(tmp_path / "aruco.detections"), str(tmp_path / "tracked.solution")
    synth.write_solution_file(a, rig); synth.write_detections_file(d, rig)
    out = subprocess.run([os.path.join(host_bins, "track"), a, d, b], check=True, capture_output=True, text=True).stdout
    assert "tracked 30 frames" in out
    res = synth.read_solution_file(b)
    assert np.array_equal(res["frame_ids"], rig.frame_ids)
    off = 6 * (rig.C - 1) + 6 * (rig.M - 1)
    t_est = res["vec"][off:off + 6 * rig.F].reshape(-1, 6)[:, 3:]
    assert np.abs(t_est - rig.T_frame_true[:, :3, 3]).max() < 5e-3


@pytest.mark.gpu
def test_eight_argument_init_path_matches_solution_file_path(host_bins, tmp_path):
    """MultiCamMapper(root_c, T_to_root_cam, ..., fcm, m_size, cam_confs) — the Initializer-output constructor: raw corners are
    undistorted on the device and pulled back into frame_cam_markers — must solve like the handle built from the file."""
    rig = synth.make_config("cfg1")
    a = str(tmp_path / "initial.solution")
    synth.write_solution_file(a, rig)
    out1 = subprocess.run([os.path.join(host_bins, "find_solution"), str(tmp_path), "0.05", "-init", a], check=True, capture_output=True, text=True).stdout
    out2 = subprocess.run([os.path.join(host_bins, "solution_tool"), "resolve", a, str(tmp_path / "resolved.solution")], check=True, capture_output=True, text=True).stdout
    c1 = float(out1.split("final_error:")[1].split()[0]); c2 = float(out2.split("final_error:")[1].split()[0])
    assert abs(c1 - c2) <= 2e-5 * c1
    A, B = synth.read_solution_file(str(tmp_path / "final.solution")), synth.read_solution_file(str(tmp_path / "resolved.solution"))
    assert np.array_equal(A["det_xy"], B["det_xy"])          # zero distortion: the device undistortion returns the corners unchanged
    assert np.abs(A["vec"][-9 * rig.C:] - B["vec"][-9 * rig.C:]).max() == 0      # intrinsics untouched


@pytest.mark.gpu
def test_track_app(host_bins, tmp_path):
    import copy
    rig = copy.copy(synth.make_rig(C=3, M=6, F=20, obs_per_frame=6.0, seed=8))
    rig.T_cam_init, rig.T_marker_init = rig.T_cam_true, rig.T_marker_true
    a, b = str(tmp_path / "frames.solution"), str(tmp_path / "tracked.solution")
    synth.write_solution_file(a, rig)
    out = subprocess.run([os.path.join(host_bins, "track"), a, b], check=True, capture_output=True, text=True).stdout
    assert "tracked 20 frames" in out
    res = synth.read_solution_file(b)
    off = 6 * (rig.C - 1) + 6 * (rig.M - 1)
    t_est = res["vec"][off:off + 6 * rig.F].reshape(-1, 6)[:, 3:]
    assert np.abs(t_est - rig.T_frame_true[:, :3, 3]).max() < 5e-3    # object positions recovered to a few mm


@pytest.mark.gpu
def test_default_config_with_intrinsics_through_the_facade(host_bins, tmp_path):
    """MultiCamMapper's default Config (multicam_mapper.h:75-81) optimises the camera intrinsics as well: the facade must hand the
    9 values per camera to the device path and read them back into the camera configurations written to the .solution file."""
    from aar_b200 import binding
    rig = synth.make_config("cfg1")
    a, b = str(tmp_path / "initial.solution"), str(tmp_path / "full.solution")
    synth.write_solution_file(a, rig)
    out = subprocess.run([os.path.join(host_bins, "solution_tool"), "resolve_full", a, b], check=True, capture_output=True, text=True).stdout
    cost = float(out.split("final_error:")[1].split()[0])
    p = binding.Problem(rig, intrinsics=True)
    z, fc, it, tr = p.solve(p.mats2evec())
    assert abs(cost - fc) <= 2e-5 * fc
    res = synth.read_solution_file(b)
    got = res["vec"][-9 * rig.C:].reshape(-1, 9); want = z[-9 * rig.C:].reshape(-1, 9)
    assert np.abs(got - want).max() <= 1e-6 * np.abs(want).max() and np.abs(got[:, :4] - rig.K[:, [0, 0, 1, 1], [0, 2, 1, 2]]).max() > 1e-6   # they moved
    assert res["flags"] == (True, True, True, True)
