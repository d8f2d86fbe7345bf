"""Pins the analytic-Jacobian / full-FP64 variant (include/aar_analytic.h) on the CPU.

The variant has no counterpart in the reference (it is the delta -> 0, no-float32 limit of the reference's central
differences), so its anchor is numerical differentiation of an INDEPENDENTLY written projection: the oracle's restated
OpenCV chain (inv44 / mul44 / K * T34 * X of MultiCamMapper::project_marker, multicam_mapper.cpp:608-649) kept in double.
The GPU tests (tests/test_gpu_analytic.py) then hold the device against this oracle."""
import numpy as np
import pytest

from aar_b200 import synth


def _csc_to_dense(colptr, rowidx, vals, rows):
    J = np.zeros((rows, len(colptr) - 1))
    for c in range(len(colptr) - 1):
        J[rowidx[colptr[c]:colptr[c + 1]], c] = vals[colptr[c]:colptr[c + 1]]
    return J


def test_rodrigues_derivatives_against_central_differences(oracle_mod):
    rng = np.random.default_rng(11)
    worst = 0.0
    for r in list(rng.normal(0, 1.0, (40, 3))) + [np.array([3.0, 0.3, -0.2]), np.array([1e-3, -2e-3, 5e-4])]:
        dR = oracle_mod.rodrigues_derivs(r)
        for k in range(3):
            d = np.zeros(3); d[k] = 1e-6
            fd = (oracle_mod.rodrigues(r + d, sincos_mode=1) - oracle_mod.rodrigues(r - d, sincos_mode=1)) / 2e-6
            worst = max(worst, np.abs(dR[k] - fd).max())
    assert worst < 2e-9, worst


def test_rodrigues_derivatives_small_angle_branch(oracle_mod):
    """|r| < 1e-5 switches to the derivative of the second-order expansion: both branches agree across the switch and the
    zero vector gives the generators [e_k]x."""
    dR0 = oracle_mod.rodrigues_derivs(np.zeros(3))
    gen = np.zeros((3, 3, 3))
    gen[0, 2, 1] = gen[1, 0, 2] = gen[2, 1, 0] = 1; gen[0, 1, 2] = gen[1, 2, 0] = gen[2, 0, 1] = -1
    assert np.array_equal(dR0, gen)
    u = np.array([0.6, -0.48, 0.64])
    below, above = oracle_mod.rodrigues_derivs(u * 0.99e-5), oracle_mod.rodrigues_derivs(u * 1.01e-5)
    assert np.abs(below - above).max() < 1e-6          # they differ by the 2 % step in r itself (first-order term ~ 1e-7)
    # against the exact first-order expansion at the same point
    for s, D in ((0.99e-5, below), (1.01e-5, above)):
        r = u * s
        for k in range(3):
            ek = np.eye(3)[k]
            lin = gen[k] + 0.5 * (np.outer(ek, r) + np.outer(r, ek)) - r[k] * np.eye(3)
            assert np.abs(D[k] - lin).max() < 1e-9


@pytest.mark.parametrize("flags", [(True, True, True), (True, False, True), (False, True, True), (True, True, False), (False, False, True)])
def test_analytic_jacobian_against_central_differences_of_the_independent_chain(oracle_mod, flags):
    rig = synth.make_rig(3, 5, 12, 6, seed=21)
    # a duplicated (frame, cam, marker) detection: only the last one owns Jacobian rows (multicam_mapper.cpp:368-370)
    rig.det_frame = np.concatenate([rig.det_frame, rig.det_frame[[3]]]); rig.det_cam = np.concatenate([rig.det_cam, rig.det_cam[[3]]])
    rig.det_marker = np.concatenate([rig.det_marker, rig.det_marker[[3]]]); rig.det_xy = np.concatenate([rig.det_xy, rig.det_xy[[3]] + 0.25])
    o = oracle_mod.Oracle(rig)
    o.set_config(cams=flags[0], markers=flags[1], objects=flags[2])
    o.set_analytic(True)
    z = o.mats2evec()
    rows, n = o.num_rows, o.num_vars
    # the shared-header residual equals the independent chain to rounding
    e_an, e_chain = o.error(z), o.error_fp64_chain(z)
    assert np.abs(e_an - e_chain).max() < 1e-9 * max(1.0, np.abs(e_chain).max())
    J = _csc_to_dense(*o.jacobian(z), rows)
    Jfd = np.zeros((rows, n))
    h = 1e-6
    for c in range(n):
        d = np.zeros(n); d[c] = h
        Jfd[:, c] = (o.error_fp64_chain(z + d) - o.error_fp64_chain(z - d)) / (2 * h)
    obs = o.observations()
    nojac_rows = np.repeat(obs["has_jac"] == 0, 8)
    assert nojac_rows.sum() == 8
    assert not J[nojac_rows].any()                       # an overwritten duplicate has residual rows but no Jacobian rows
    scale = np.abs(Jfd).max()
    dev = np.abs(J - Jfd)[~nojac_rows].max() / scale
    assert dev < 5e-8, dev
    # and it is the limit of the reference's scheme: the faithful float32 / delta = 1e-3 Jacobian is close, not equal
    o.set_analytic(False)
    Jref = _csc_to_dense(*o.jacobian(z), rows)
    rel = np.abs(J - Jref)[~nojac_rows].max() / scale
    assert 1e-12 < rel < 5e-3, rel


def test_analytic_jacobian_on_a_planar_board(oracle_mod):
    """All markers coplanar and aligned with the root marker (a printed board): their rotation vectors are exactly zero, which is the
    small-angle branch of the Rodrigues derivatives inside a full Jacobian; one camera pose is the identity rotation as well."""
    rig = synth.make_rig(3, 5, 10, 6, seed=33)
    for T in (rig.T_marker_init, rig.T_marker_true):
        T[:, :3, :3] = np.eye(3)
    rig.T_cam_init[1, :3, :3] = np.eye(3)
    o = oracle_mod.Oracle(rig); o.set_analytic(True)
    z = o.mats2evec()
    assert np.count_nonzero(z[6 * 2:6 * 2 + 6 * 4].reshape(-1, 6)[:, :3]) == 0 and not z[0:3].any()      # marker blocks and camera 1: r = 0
    rows, n = o.num_rows, o.num_vars
    J = _csc_to_dense(*o.jacobian(z), rows)
    Jfd = np.zeros((rows, n)); h = 1e-6
    for c in range(n):
        d = np.zeros(n); d[c] = h
        Jfd[:, c] = (o.error_fp64_chain(z + d) - o.error_fp64_chain(z - d)) / (2 * h)
    assert np.abs(J - Jfd).max() / np.abs(Jfd).max() < 5e-8


def test_analytic_solve_reaches_the_faithful_optimum(oracle_mod):
    """The two variants minimise (nearly) the same function: from the same start the reference solver, driven by the analytic
    functions, ends within the noise floor of the faithful solve."""
    rig = synth.make_rig(3, 6, 30, 6, seed=5)
    o = oracle_mod.Oracle(rig)
    z0 = o.mats2evec()
    z_f, cost_f, it_f, _ = o.solve(z0)
    o.set_analytic(True)
    z_a, cost_a, it_a, _ = o.solve(z0)
    assert it_a > 2 and cost_a < 0.05 * float(np.sum(o.error(z0) ** 2))
    assert abs(cost_a - cost_f) / cost_f < 1e-3
    assert np.abs(z_a - z_f).max() < 1e-3
