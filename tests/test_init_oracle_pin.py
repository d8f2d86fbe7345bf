"""CPU tests pinning the initialisation oracle (oracle/init_oracle.cpp) to OpenCV: tests/golden/init_golden.npz holds cv2 4.13
answers (generator: tests/golden/make_golden_init.py).  The reference ships no tests or vectors for this path."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(ROOT, "tests", "golden", "init_golden.npz"))


def test_undistort_normalised_bit_exact(oracle_mod, g):
    ms = np.float32(g["marker_size"])
    for n in range(len(g["xy"])):
        c = int(g["cam"][n])
        q = oracle_mod.init_ippe_raw(ms, g["xy"][n], g["K"][c], g["dist"][c])[0]
        assert np.array_equal(q, g["und_norm"][n])


def test_ippe_matches_opencv_ippe_square(oracle_mod, g):
    """aruco's IPPE (ippe.cpp:118-219) against OpenCV's own port of the algorithm: both solutions, in the same order; the oracle's
    poses are float32-rounded (getRTMatrix CV_32F) and computed from float32 normalised points, hence the tolerance."""
    ms = np.float32(g["marker_size"]); worst_R = worst_t = 0.0
    for n in range(len(g["xy"])):
        c = int(g["cam"][n])
        T, e = oracle_mod.init_solve_pnp(ms, g["xy"][n], g["K"][c], g["dist"][c], sincos_mode=0)
        assert e[0] <= e[1]
        for k in range(2):
            worst_R = max(worst_R, np.abs(T[k, :3, :3] - g["ippe_R"][n, k]).max())
            worst_t = max(worst_t, np.abs(T[k, :3, 3] - g["ippe_t"][n, k]).max() / np.abs(g["ippe_t"][n, k]).max())
            assert np.array_equal(T[k], T[k].astype(np.float32).astype(np.float64)) and np.array_equal(T[k, 3], [0, 0, 0, 1])
    assert worst_R < 1e-5 and worst_t < 2e-6, (worst_R, worst_t)


def test_shared_acos_and_sincos_rarely_flip_a_float(oracle_mod, g):
    """Mode 1 (aar_acos / aar_sincos, what the device runs) against mode 0 (libm, what the reference calls): the float32-rounded
    poses agree except where an ulp-level difference straddles a float32 rounding boundary."""
    ms = np.float32(g["marker_size"]); flips = total = 0
    for n in range(len(g["xy"])):
        c = int(g["cam"][n])
        T0, e0 = oracle_mod.init_solve_pnp(ms, g["xy"][n], g["K"][c], g["dist"][c], sincos_mode=0)
        T1, e1 = oracle_mod.init_solve_pnp(ms, g["xy"][n], g["K"][c], g["dist"][c], sincos_mode=1)
        flips += int((T0 != T1).sum()); total += 24
        assert np.array_equal(e0, e1) and np.abs(T0 - T1).max() <= 2.4e-7
    assert flips <= total * 1e-3, (flips, total)
    rng = np.random.default_rng(5)
    xs = np.concatenate([rng.uniform(-1, 1, 20000), [1.0, -1.0, 0.0, 0.5, -0.5, 1 - 1e-16, -1 + 1e-16, 1e-300]])
    err = max(abs(oracle_mod.acos_shared(float(x)) - np.arccos(x)) / max(np.arccos(x), 1e-300) for x in xs if abs(x) < 1)
    assert err <= 2.3e-16, err
    assert oracle_mod.acos_shared(1.0) == 0.0 and oracle_mod.acos_shared(-1.0) == np.pi


def test_consensus_bit_exact_against_cv2_twin(oracle_mod, g):
    o = 0
    for i, n in enumerate(g["cons_n"]):
        sl = slice(o, o + n); o += n
        bi, w = oracle_mod.init_consensus(0.05, g["cons_T"][sl], g["cons_T1inv"][sl], g["cons_T2inv"][sl])
        assert bi == int(g["cons_best"][i]) and w == float(g["cons_err"][i])


def test_initializer_recovers_a_synthetic_rig(oracle_mod):
    from aar_b200 import synth
    rig = synth.make_rig(C=3, M=6, F=60, obs_per_frame=6.0, seed=1)
    nF = int(rig.frame_ids.max()) + 1
    o = oracle_mod.InitOracle(rig.C, rig.K, rig.dist, float(rig.marker_size), nF, rig.det_frame, rig.det_cam, rig.det_marker, rig.det_xy)
    T, err, nc = o.obtain_pose_estimations()
    assert set(np.unique(nc)) <= {1, 2} and (err[:, 0] <= err[:, 1]).all()
    o.init_transforms()
    r = o.results()
    assert np.array_equal(r["cam_ids"], rig.cam_ids) and np.array_equal(r["marker_ids"], rig.marker_ids)
    assert r["root_cam"] == rig.root_cam and r["root_marker"] == rig.root_marker
    ci, cT = r["cams"]; mi, mT = r["markers"]; fi, fT = r["objects"]
    assert np.array_equal(ci, rig.cam_ids) and np.array_equal(mi, rig.marker_ids) and np.array_equal(fi, rig.frame_ids)
    assert np.abs(cT - rig.T_cam_true).max() < 0.1 and np.abs(mT - rig.T_marker_true).max() < 0.1 and np.abs(fT - rig.T_frame_true).max() < 0.2
    # the sampled consensus (consensus_max) with a bound above every list length is the exhaustive one
    o2 = oracle_mod.InitOracle(rig.C, rig.K, rig.dist, float(rig.marker_size), nF, rig.det_frame, rig.det_cam, rig.det_marker, rig.det_xy, consensus_max=10 ** 6)
    o2.obtain_pose_estimations(); o2.init_transforms()
    r2 = o2.results()
    assert np.array_equal(r2["cams"][1], cT) and np.array_equal(r2["objects"][1], fT)
