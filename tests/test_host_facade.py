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


def test_calib_reader(host_bins, tmp_path):
    rig = synth.make_rig(C=2, M=3, F=4, obs_per_frame=3.0, seed=6, distorted=True)
    synth.write_calib_files(str(tmp_path), rig)
    txt = open(tmp_path / "1" / "calib.yml").read()
    assert "camera_matrix" in txt and "distortion_coefficients" in txt


@pytest.mark.gpu
def test_find_solution_app_matches_binding(host_bins, tmp_path):
    from aar_b200 import binding
    rig = synth.make_config("cfg1")
    synth.write_solution_file(str(tmp_path / "initial.solution"), rig)
    out = subprocess.run([os.path.join(host_bins, "find_solution"), str(tmp_path), "0.05"], check=True, capture_output=True, text=True).stdout
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
def test_eight_argument_init_path_matches_solution_file_path(host_bins, tmp_path):
    """MultiCamMapper(root_c, T_to_root_cam, ..., fcm, m_size, cam_confs) — the Initializer-output constructor: raw corners are
    undistorted on the device and pulled back into frame_cam_markers — must solve like the handle built from the file."""
    rig = synth.make_config("cfg1")
    a = str(tmp_path / "initial.solution")
    synth.write_solution_file(a, rig)
    out1 = subprocess.run([os.path.join(host_bins, "find_solution"), str(tmp_path), "0.05"], check=True, capture_output=True, text=True).stdout
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
