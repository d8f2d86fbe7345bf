"""GPU parity tests of the initialisation path (include/aar_init.h) against the restated Initializer
(oracle/init_oracle.cpp): IPPE poses, candidate counts, consensus winners and errors, spanning-tree transforms and per-frame
object poses must be BIT-EXACT — every operation on this path is an IEEE +, -, *, /, sqrt in the reference's order, and the two
transcendental calls (acos, sin / cos) are shared implementations (include/aar_acos.h, include/aar_crsincos.h)."""
import os

import numpy as np
import pytest

from aar_b200 import synth
from conftest import parity_record

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def binding():
    from aar_b200 import binding as b
    b.lib()
    return b


def _pair(binding, oracle_mod, rig, **kw):
    nF = int(rig.frame_ids.max()) + 1
    o = oracle_mod.InitOracle(rig.C, rig.K, rig.dist, float(rig.marker_size), nF, rig.det_frame, rig.det_cam, rig.det_marker, rig.det_xy, **kw)
    g = binding.Initializer.from_rig(rig, **kw)
    return o, g


def _same_results(ro, rg):
    assert np.array_equal(ro["cam_ids"], rg["cam_ids"]) and np.array_equal(ro["marker_ids"], rg["marker_ids"])
    assert ro["root_cam"] == rg["root_cam"] and ro["root_marker"] == rg["root_marker"]
    for k in ("cams", "markers", "objects"):
        assert np.array_equal(ro[k][0], rg[k][0]), k
        assert np.array_equal(ro[k][1], rg[k][1]), (k, np.abs(ro[k][1] - rg[k][1]).max())


@pytest.mark.parametrize("distorted", [False, True])
def test_ippe_estimations_bit_exact(binding, oracle_mod, distorted):
    rig = synth.make_rig(C=4, M=8, F=80, obs_per_frame=10.0, seed=21, distorted=distorted)
    o, g = _pair(binding, oracle_mod, rig)
    To, eo, no = o.obtain_pose_estimations()
    Tg, eg, ng = g.estimations()
    assert np.array_equal(no, ng) and set(np.unique(no)) <= {1, 2}
    assert np.array_equal(eo, eg)
    assert np.array_equal(To, Tg), np.abs(To - Tg).max()
    parity_record("init_ippe_vs_oracle", detections=int(rig.N), distorted=bool(distorted), poses_bit_exact=True, errors_bit_exact=True,
                  second_solution_kept=int((no == 2).sum()))


def test_consensus_kernel_against_cv2_twin_goldens(binding, oracle_mod):
    g = np.load(os.path.join(ROOT, "tests", "golden", "init_golden.npz"))
    o = 0
    for i, n in enumerate(g["cons_n"]):
        sl = slice(o, o + n); o += n
        bi, w = binding.init_consensus(0.05, g["cons_T"][sl], g["cons_T1inv"][sl], g["cons_T2inv"][sl])
        assert bi == int(g["cons_best"][i]) and w == float(g["cons_err"][i]), (i, bi, w)
    # a list longer than one CTA (several jobs, cross-CTA pick): against the oracle
    rng = np.random.default_rng(3)
    T = np.tile(g["cons_T"][:30], (20, 1, 1)); A = np.tile(g["cons_T1inv"][:30], (20, 1, 1)); B = np.tile(g["cons_T2inv"][:30], (20, 1, 1))
    T[:, :3, 3] += rng.normal(0, 1e-3, (len(T), 3))
    bo, wo = oracle_mod.init_consensus(0.05, T, A, B)
    bg, wg = binding.init_consensus(0.05, T, A, B)
    assert (bo, wo) == (bg, wg)


@pytest.mark.parametrize("cfg", ["small", "cfg1", "distorted"])
def test_rig_and_object_initialisation_bit_exact(binding, oracle_mod, cfg):
    if cfg == "small":
        rig = synth.make_rig(C=3, M=6, F=60, obs_per_frame=6.0, seed=1)
    elif cfg == "cfg1":
        rig = synth.make_config("cfg1")
    else:
        rig = synth.make_rig(C=5, M=10, F=120, obs_per_frame=12.0, seed=4, distorted=True)
    o, g = _pair(binding, oracle_mod, rig)
    o.obtain_pose_estimations(); o.init_transforms()
    g.init_transforms(); g.init_object_transforms()
    ro, rg = o.results(), g.results()
    _same_results(ro, rg)
    for cams in (True, False):
        eo, eg = o.edges(cams), g.edges(cams)
        for a, b in zip(eo, eg):
            assert np.array_equal(a, b)
    dev_c = float(np.abs(rg["cams"][1] - rig.T_cam_true).max()); dev_f = float(np.abs(rg["objects"][1] - rig.T_frame_true).max())
    parity_record("init_rig_and_objects_vs_oracle", workload=cfg, cameras=int(rig.C), markers=int(rig.M), frames=int(rig.F), detections=int(rig.N),
                  bit_exact=True, cam_edges=int(len(g.edges(True)[0])), marker_edges=int(len(g.edges(False)[0])),
                  max_abs_dev_cam_T_vs_truth=dev_c, max_abs_dev_object_T_vs_truth=dev_f, timings=g.timings())


def test_sampled_consensus_excluded_cameras_and_sparse_frames(binding, oracle_mod):
    rig = synth.make_rig(C=4, M=8, F=90, obs_per_frame=9.0, seed=9)
    # frames 3 and 4 lose all but one detection (skipped by min_detections); a duplicated detection of one (frame, cam, marker)
    keep = np.ones(rig.N, bool)
    for f in rig.frame_ids[3:5]:
        idx = np.nonzero(rig.det_frame == f)[0]; keep[idx[1:]] = False
    rig.det_frame, rig.det_cam, rig.det_marker, rig.det_xy = rig.det_frame[keep], rig.det_cam[keep], rig.det_marker[keep], rig.det_xy[keep]
    d = int(np.nonzero(rig.det_frame == rig.frame_ids[10])[0][0])
    ins = d + 1
    rig.det_frame = np.insert(rig.det_frame, ins, rig.det_frame[d]); rig.det_cam = np.insert(rig.det_cam, ins, rig.det_cam[d])
    rig.det_marker = np.insert(rig.det_marker, ins, rig.det_marker[d]); rig.det_xy = np.insert(rig.det_xy, ins, rig.det_xy[d] + 0.25, axis=0)
    for kw in (dict(consensus_max=16), dict(excluded=[2]), dict(threshold=1.2), dict(consensus_max=7, excluded=[0])):
        o, g = _pair(binding, oracle_mod, rig, **kw)
        To, eo, no = o.obtain_pose_estimations()
        Tg, eg, ng = g.estimations()
        assert np.array_equal(no, ng) and np.array_equal(To, Tg) and (no == 0).any()
        o.init_transforms(); g.init_transforms(); g.init_object_transforms()
        _same_results(o.results(), g.results())
        assert len(g.results()["objects"][0]) == rig.F - 2
        for cams in (True, False):
            for a, b in zip(o.edges(cams), g.edges(cams)):
                assert np.array_equal(a, b)


def test_track_flow_object_transforms_against_a_given_rig(binding, oracle_mod):
    """apps/track.cpp:128-131: set_transforms_to_root_* from a solved rig, obtain_pose_estimations, init_object_transforms."""
    rig = synth.make_rig(C=6, M=12, F=150, obs_per_frame=24.0, seed=12)
    o, g = _pair(binding, oracle_mod, rig)
    o.obtain_pose_estimations()
    # cameras / markers without a transform fall back to the identity (initializer.cpp:83-100): leave one of each out
    o.set_rig(rig.cam_ids[:-1], rig.T_cam_true[:-1], rig.marker_ids[1:], rig.T_marker_true[1:])
    g.set_rig(rig.cam_ids[:-1], rig.T_cam_true[:-1], rig.marker_ids[1:], rig.T_marker_true[1:])
    o.init_object_transforms(); g.init_object_transforms()
    fo, To = o.results()["objects"]; fg, Tg = g.results()["objects"]
    assert np.array_equal(fo, fg) and np.array_equal(To, Tg)
    parity_record("init_track_flow_objects_vs_oracle", frames=int(len(fg)), detections=int(rig.N), bit_exact=True, timings=g.timings())
