"""GPU tests of the analytic-Jacobian / full-FP64 variant (aar_problem_desc::analytic_jacobian): the CUDA path
against the CPU oracle running the same variant.  The arithmetic of one observation is ONE implementation for host and device
(include/aar_analytic.h; tests/test_analytic_cpu.py pins it against numerical differentiation of an independent projection chain),
so residuals and Jacobian entries are expected to agree bit for bit; the asserted bar is 1e-12 relative and the record says whether
they were identical.  Sums of products (reduced system, LM trajectory) carry the tolerances written at each assert."""
import numpy as np
import pytest

from aar_b200 import synth
from conftest import parity_record

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def binding():
    from aar_b200 import binding as b
    b.lib()
    return b


def _rig_with_erasures(seed=3, **kw):
    args = dict(C=3, M=6, F=40, obs_per_frame=6.0, seed=seed); args.update(kw)
    rig = synth.make_rig(**args)
    # a detection of an unknown camera, one of an unknown marker, and a duplicated (frame, cam, marker): the erasures of
    # fill_iteration_arrays (multicam_mapper.cpp:353-370) — the duplicate keeps its residual rows and loses its Jacobian rows
    extra_f = np.array([rig.det_frame[5], rig.det_frame[9], rig.det_frame[12]], np.int32)
    extra_c = np.array([77, rig.det_cam[9], rig.det_cam[12]], np.int32)
    extra_m = np.array([rig.det_marker[5], 9999, rig.det_marker[12]], np.int32)
    rig.det_frame = np.concatenate([rig.det_frame, extra_f]); rig.det_cam = np.concatenate([rig.det_cam, extra_c])
    rig.det_marker = np.concatenate([rig.det_marker, extra_m]); rig.det_xy = np.concatenate([rig.det_xy, rig.det_xy[[5, 9, 12]] + 0.5])
    return rig


@pytest.mark.parametrize("distorted", [False, True])
def test_analytic_residual_and_jacobian_match_the_oracle(binding, oracle_mod, distorted):
    rig = _rig_with_erasures(seed=4, distorted=distorted)
    worst_r = worst_j = 0.0; identical = True
    for cams, markers, objects in [(True, True, True), (False, True, True), (True, False, True), (True, True, False), (False, False, True)]:
        o = oracle_mod.Oracle(rig); o.set_config(cams=cams, markers=markers, objects=objects); o.set_analytic(True)
        p = binding.Problem(rig, cams=cams, markers=markers, objects=objects, analytic=True)
        z = o.mats2evec()
        assert p.num_vars == o.num_vars == len(z)
        r_o = o.error(z); r_g, ss = p.residual(z)
        dr = np.abs(r_g - r_o).max()
        assert dr <= 1e-12 * max(1.0, np.abs(r_o).max())
        assert abs(ss - r_o @ r_o) <= 1e-12 * (r_o @ r_o)
        cp_o, ri_o, v_o = o.jacobian(z); cp_g, ri_g, v_g = p.jacobian(z)
        assert np.array_equal(cp_o, cp_g) and np.array_equal(ri_o, ri_g)          # same sparsity pattern as the faithful Jacobian
        dj = np.abs(v_g - v_o).max() / np.abs(v_o).max()
        assert dj <= 1e-12
        worst_r = max(worst_r, dr); worst_j = max(worst_j, dj)
        identical = identical and np.array_equal(r_g, r_o) and np.array_equal(v_g, v_o)
        p.close()
    parity_record("analytic_residual_and_jacobian_vs_oracle" + ("_distorted" if distorted else ""), max_abs_dev_residual_px=worst_r,
                  max_rel_dev_jacobian=worst_j, bit_identical=bool(identical), bar=1e-12)


def test_analytic_huber_residual_and_unsupported_combination(binding, oracle_mod):
    rig = _rig_with_erasures(seed=8)
    rig.det_xy[::11] += 12.0
    o = oracle_mod.Oracle(rig); o.set_config(with_huber=True, huber_delta=1.5); o.set_analytic(True)
    p = binding.Problem(rig, with_huber=True, analytic=True)
    z = o.mats2evec()
    r_o = o.error(z); r_g, _ = p.residual(z, huber_delta=1.5)
    assert np.abs(r_g - r_o).max() <= 1e-12 * np.abs(r_o).max()
    with pytest.raises(binding.AarError):                      # the intrinsics block exists for the central-difference path only
        binding.Problem(rig, intrinsics=True, analytic=True)


def test_analytic_reduced_system_matches_the_oracle(binding, oracle_mod):
    rig = _rig_with_erasures(seed=9, F=50)
    o = oracle_mod.Oracle(rig); o.set_analytic(True)
    p = binding.Problem(rig, analytic=True)
    z = o.mats2evec(); mu = 1234.5
    S_o, b_o, c_o = o.reduced_system(z, mu)
    S_g, b_g, c_g = p.reduced_system(z, mu)
    iu = np.triu_indices(p.n_r)
    scale = np.abs(S_o).max()
    dS = np.abs(S_g[iu] - S_o[iu]).max() / scale; db = np.abs(b_g - b_o).max() / np.abs(b_o).max(); dc = abs(c_g - c_o) / c_o
    parity_record("analytic_reduced_schur_system_vs_oracle", n_r=int(p.n_r), rel_dev_S=dS, rel_dev_b=db, rel_dev_cost=dc, bar=1e-10)
    assert dS <= 1e-10 and db <= 1e-10 and dc <= 1e-12


@pytest.mark.parametrize("name,huber", [("cfg1", False), ("cfg2", False), ("cfg1", True)])
def test_analytic_lm_matches_the_reference_solver(binding, oracle_mod, name, huber):
    """SparseLevMarq::solve (the unmodified reference header when oracle/_ref is built) driven by the analytic residual and
    Jacobian of the oracle, against the device-resident loop in the same variant: single steps to 1e-10, the full solve to the
    north star's 1e-6 (no float32 quantisation in this variant, so no reproducibility envelope is needed)."""
    rig = synth.make_config(name, seed=21 if huber else None)
    if huber:
        rig.det_xy[::37] += 25.0
    o = oracle_mod.Oracle(rig); o.set_config(with_huber=huber); o.set_analytic(True)
    p = binding.Problem(rig, with_huber=huber, analytic=True)
    z0 = o.mats2evec()
    worst = {}
    for k in (1, 3):
        o.set_max_iters(k)
        z_o, fc_o, it_o, tr_o = o.solve(z0)
        z_g, fc_g, it_g, tr_g = p.solve(z0, binding.Problem.default_params(max_iters=k))
        assert it_o == it_g == k
        dc = abs(fc_g - fc_o) / fc_o; dz = np.abs(z_g - z_o).max() / np.abs(z_o).max()
        worst[k] = (dc, dz)
        assert dc <= 1e-10 and dz <= 1e-10, (k, dc, dz)
        assert np.allclose(tr_g[:, 1], tr_o[:, 1], rtol=1e-9, atol=0)
    cap = 40 if huber else 10000            # with outliers the solve runs for 500 iterations (hubberDelta annealing): 40 of them are compared
    o.set_max_iters(cap)
    z_o, fc_o, it_o, tr_o = o.solve(z0)
    z_g, fc_g, it_g, tr_g = p.solve(z0, binding.Problem.default_params(max_iters=cap))
    n = min(it_o, it_g)
    d_trace = np.abs(tr_g[:n, 0] - tr_o[:n, 0]) / tr_o[:n, 0]
    d_cost = abs(fc_g - fc_o) / fc_o; d_z = np.abs(z_g - z_o).max() / max(1.0, np.abs(z_o).max())
    r_fin = o.error(z_g)
    parity_record("analytic_lm_solve_" + name + ("_huber_outliers" if huber else ""), iterations_ref=it_o, iterations_gpu=it_g,
                  final_cost_ref=fc_o, final_cost_gpu=fc_g, rel_dev_final_cost=d_cost, rel_dev_z=d_z, per_iteration_cost_rel_dev=d_trace,
                  one_step_rel_dev_cost_z=worst[1], three_steps_rel_dev_cost_z=worst[3],
                  oracle_cost_at_gpu_z_rel_dev=abs(r_fin @ r_fin - fc_g) / fc_g, graph_loop=p.stats()["graph_loop"], north_star=1e-6)
    assert abs(it_o - it_g) <= 1
    assert d_cost <= 1e-6 and d_z <= 1e-6
    if not huber:                               # (with Huber the oracle's hubberDelta has moved on since the accepted evaluation)
        assert abs(r_fin @ r_fin - fc_g) <= 1e-12 * fc_g
    if huber:
        assert np.array_equal(tr_g[:min(n, 8), 5], tr_o[:min(n, 8), 5])          # hubberDelta schedule of optCallBack


def test_analytic_and_faithful_solves_agree_at_cfg3_size(binding, oracle_mod):
    """cfg 3 (0.5 M observations): the analytic variant's residual agrees with the oracle at full size, its solve decreases the cost
    monotonically to the 0.3 px noise floor and ends next to the faithful path's optimum (both minimise the same reprojection
    error; they differ by the reference's float32 / finite-difference quantisation)."""
    rig = synth.make_config("cfg3")
    o = oracle_mod.Oracle(rig); o.set_analytic(True)
    pa = binding.Problem(rig, analytic=True)
    z0 = o.mats2evec()
    r_o = o.error(z0); r_g, _ = pa.residual(z0)
    d_res = np.abs(r_g - r_o).max()
    assert d_res <= 1e-12 * np.abs(r_o).max()
    z_a, fc_a, it_a, tr_a = pa.solve(z0)
    assert np.all(np.diff(tr_a[:, 0]) < 0)
    rms = np.sqrt(fc_a / (8 * pa.num_obs))
    assert 0.25 < rms < 0.35
    r_fin = o.error(z_a)
    assert abs(r_fin @ r_fin - fc_a) <= 1e-12 * fc_a
    pa.close()
    pf = binding.Problem(rig)
    z_f, fc_f, it_f, tr_f = pf.solve(z0)
    d_cost = abs(fc_a - fc_f) / fc_f; d_z = np.abs(z_a - z_f).max()
    parity_record("analytic_vs_faithful_solve_cfg3", observations=int(pf.num_obs), residual_max_abs_dev_vs_oracle_px=d_res, residual_bit_identical=bool(np.array_equal(r_g, r_o)),
                  iterations_analytic=it_a, iterations_faithful=it_f, final_cost_analytic=fc_a, final_cost_faithful=fc_f, rel_dev_final_cost=d_cost,
                  max_abs_dev_z=d_z, rms_px=rms)
    assert d_cost <= 1e-3 and d_z <= 1e-3
