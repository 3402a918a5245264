"""Multi-period ACOPF on the CUDA path vs the reference's goldens and the oracle, through the C ABI
(run on the B200 box). Mirrors test/algorithms/mpacopf_update_gpu.jl: operator-level known
answers for case9 x 3 periods, then the end-to-end solve (Solved / 20 / 729 / 15901.48)."""
import json
from pathlib import Path

import numpy as np
import pytest

import exaadmm_b200 as ea
from exaadmm_b200 import operators as ops
from exaadmm_b200.admm_two_level import admm_two_level
from exaadmm_b200.environment import AdmmEnv, Parameters
from exaadmm_b200.mpacopf import ModelMpacopf, solve_mpacopf
from exaadmm_b200.solve_acopf import solve_acopf
from exaadmm_b200.synthetic import synthetic_case
from oracle.oracle import OracleMpModel, OracleModel

pytestmark = pytest.mark.gpu

GOLD = json.loads((Path(__file__).resolve().parent / "golden" / "mpacopf_case9_golden.json").read_text())
ITERATE_TOL = 1e-8


def write_profile(tmp_path, grid, scales, name="profile"):
    """<prefix>.Pd / .Qd: one row per bus, one column per period (opfdata.jl:121-130)."""
    prefix = tmp_path / name
    np.savetxt(str(prefix) + ".Pd", np.outer(np.asarray(grid.Pd), scales), fmt="%.17g")
    np.savetxt(str(prefix) + ".Qd", np.outer(np.asarray(grid.Qd), scales), fmt="%.17g")
    return str(prefix)


def make_models(tmp_path, case, grid, scales, rho_pq=4e2, rho_va=4e4, ramp_ratio=0.02, **params):
    prefix = write_profile(tmp_path, grid, scales)
    env = AdmmEnv(case, rho_pq, rho_va, use_gpu=True, verbose=0, load_prefix=prefix)
    mod = ModelMpacopf(env, start_period=1, end_period=len(scales), ramp_ratio=ramp_ratio)
    par = Parameters(); par.verbose = 0
    for k, v in params.items():
        setattr(env.params, k, v); setattr(par, k, v)
    Pd, Qd = np.outer(scales, grid.Pd), np.outer(scales, grid.Qd)
    om = OracleMpModel(grid, par, rho_pq, rho_va, Pd, Qd, ramp_ratio=ramp_ratio)
    return env, mod, om


def _presteps(env, mod):
    ops.admm_increment_outer(env, mod); ops.admm_outer_prestep(env, mod)
    ops.admm_increment_reset_inner(env, mod); ops.admm_increment_inner(env, mod); ops.admm_inner_prestep(env, mod)


def test_operator_level_goldens_case9_three_periods(tmp_path, case9_grid):
    atol, T = GOLD["atol"], GOLD["len_horizon"]
    env, mod, _ = make_models(tmp_path, ea.CASE9, case9_grid, GOLD["load_scale"], scale=1e-4, initial_beta=1e3, beta=1e3)
    assert mod.nvar == mod.models[0].nvar + mod.ngen and mod.len_horizon == T
    _presteps(env, mod)

    def check(field, gen, br):
        for t in range(T):
            x = getattr(mod.models[t].solution, field)
            np.testing.assert_allclose(x[:6], GOLD[gen][t], atol=atol, rtol=0, err_msg=f"{gen} t={t}")
            np.testing.assert_allclose(x[6:], GOLD[br][t], atol=atol, rtol=0, err_msg=f"{br} t={t}")

    ops.admm_update_x(env, mod);    check("u_curr", "U_GEN", "U_BR")
    ops.admm_update_xbar(env, mod); check("v_curr", "V_GEN", "V_BR")
    ops.admm_update_z(env, mod);    check("z_curr", "Z_GEN", "Z_BR")
    ops.admm_update_l(env, mod);    check("l_curr", "L_GEN", "L_BR")
    ops.admm_update_residual(env, mod)
    for t in range(T):
        s = mod.models[t].solution
        u, v, z = s.u_curr, s.v_curr, s.z_curr
        np.testing.assert_allclose(s.rp, u - v + z, atol=atol)
        np.testing.assert_allclose(s.rd, z - s.z_prev, atol=atol)
        np.testing.assert_allclose(s.Ax_plus_By, u - v, atol=atol)
        if t > 0:
            r = mod.solution[t]
            vp = mod.models[t - 1].solution.v_curr[0:2 * mod.ngen:2]
            np.testing.assert_allclose(r.rp, r.u_curr - vp + r.z_curr, atol=atol)
            np.testing.assert_allclose(r.rd, r.z_curr - r.z_prev, atol=atol)
            np.testing.assert_allclose(r.Ax_plus_By, r.u_curr - vp, atol=atol)
    lz_prev = [mod.models[t].solution.lz for t in range(T)]
    lzr_prev = [mod.solution[t].lz for t in range(T)]
    ops.admm_update_lz(env, mod)
    for t in range(T):
        np.testing.assert_allclose(mod.models[t].solution.lz, lz_prev[t] + env.params.beta * mod.models[t].solution.z_curr, atol=atol)
        if t > 0:
            np.testing.assert_allclose(mod.solution[t].lz, lzr_prev[t] + env.params.beta * mod.solution[t].z_curr, atol=atol)
    mod.close()


def _compare_state(mod, om, tol, what, beta=1.0):
    """lambda = -(lz + beta z): a difference of tol in z shows up as beta * tol in lambda, so the multipliers are
    compared with atol = beta * tol (beta = 1 for the short lock-step runs, where everything agrees to tol)."""
    for t in range(mod.len_horizon):
        for f in ("u_curr", "v_curr", "z_curr", "l_curr"):
            a = tol * beta if f == "l_curr" else tol
            np.testing.assert_allclose(getattr(mod.models[t].solution, f), om.models[t].vec(f), atol=a, rtol=tol,
                                       err_msg=f"{what}: {f} t={t}")
        if t > 0:
            for f in ("u_curr", "s_curr", "z_curr", "l_curr"):
                a = tol * beta if f == "l_curr" else tol
                np.testing.assert_allclose(getattr(mod.solution[t], f), om.ramp(t, f), atol=a, rtol=tol,
                                           err_msg=f"{what}: ramp {f} t={t}")


def test_stepwise_iterates_match_oracle(tmp_path, case9_grid):
    """Five inner iterations operator by operator; every vector (periods and ramp) against the oracle."""
    env, mod, om = make_models(tmp_path, ea.CASE9, case9_grid, [1.0, 0.97, 1.03, 0.99], scale=1e-4, initial_beta=1e3, beta=1e3)
    ops.admm_increment_outer(env, mod); om.admm_increment_outer()
    ops.admm_outer_prestep(env, mod); om.admm_outer_prestep()
    ops.admm_increment_reset_inner(env, mod); om.admm_increment_reset_inner()
    for it in range(5):
        ops.admm_increment_inner(env, mod); om.admm_increment_inner()
        ops.admm_inner_prestep(env, mod); om.admm_inner_prestep()
        ops.admm_update_x(env, mod); om.admm_update_x()
        ops.admm_update_xbar(env, mod); om.admm_update_xbar()
        ops.admm_update_z(env, mod); om.admm_update_z()
        ops.admm_update_l(env, mod); om.admm_update_l()
        ops.admm_update_residual(env, mod); res = om.admm_update_residual()
        _compare_state(mod, om, ITERATE_TOL, f"iteration {it + 1}")
        got = [mod.info.primres, mod.info.dualres, mod.info.norm_z_curr, mod.info.mismatch]
        np.testing.assert_allclose(got, res, rtol=1e-9)
    ops.admm_update_lz(env, mod); om.admm_update_lz()
    ops.admm_outer_prestep(env, mod); om.admm_outer_prestep()
    assert mod.info.norm_z_prev == pytest.approx(om.norm_z_prev, rel=1e-9)
    for t in range(1, mod.len_horizon):
        np.testing.assert_allclose(mod.solution[t].lz, om.ramp(t, "lz"), atol=ITERATE_TOL, rtol=ITERATE_TOL)
        np.testing.assert_allclose(mod.gen_membuf(t), om.gen_membuf(t)[6:8], rtol=1e-9, atol=1e-9)
    mod.close()


def test_fused_iterations_match_stepwise(tmp_path, case9_grid):
    """ea_mp_run_inner (device-side loop) == the operator sequence, iterate for iterate."""
    env, mod, om = make_models(tmp_path, ea.CASE9, case9_grid, [1.0, 0.98, 1.02], scale=1e-4, initial_beta=1e3, beta=1e3,
                               inner_iterlim=7)
    ops.admm_increment_outer(env, mod); om.admm_increment_outer()
    ops.admm_outer_prestep(env, mod); om.admm_outer_prestep()
    ops.admm_increment_reset_inner(env, mod)
    ops.admm_run_inner(env, mod)
    assert mod.info.inner == 7
    om.admm_increment_reset_inner()
    for _ in range(7):
        om.admm_increment_inner(); om.admm_inner_prestep(); om.admm_update_x(); om.admm_update_xbar()
        om.admm_update_z(); om.admm_update_l(); res = om.admm_update_residual()
    _compare_state(mod, om, ITERATE_TOL, "after 7 fused iterations")
    np.testing.assert_allclose([mod.info.primres, mod.info.dualres, mod.info.norm_z_curr, mod.info.mismatch], res, rtol=1e-9)
    # z_prev and the residual vectors are left as the reference's last iteration leaves them
    for t in range(mod.len_horizon):
        np.testing.assert_allclose(mod.models[t].solution.z_prev, om.models[t].vec("z_prev"), atol=ITERATE_TOL)
        np.testing.assert_allclose(mod.models[t].solution.rp, om.models[t].vec("rp"), atol=ITERATE_TOL)
        if t > 0:
            np.testing.assert_allclose(mod.solution[t].z_prev, om.ramp(t, "z_prev"), atol=ITERATE_TOL)
            np.testing.assert_allclose(mod.solution[t].rp, om.ramp(t, "rp"), atol=ITERATE_TOL)
    mod.close()


@pytest.mark.parametrize("mode", ["stepwise", "fused", "native"])
def test_case9_three_periods_known_answer(tmp_path, case9_grid, mode):
    pin = GOLD["solve_case9_T3"]
    prefix = write_profile(tmp_path, case9_grid, GOLD["load_scale"])
    env, mod = solve_mpacopf(ea.CASE9, prefix, use_gpu=True, verbose=0, mode=mode, **pin["kwargs"])
    assert mod.info.status == "Solved"
    assert (mod.info.outer, mod.info.cumul) == (pin["outer"], pin["cumul"])
    assert abs(mod.info.objval - pin["objval"]) <= pin["objval_atol"]
    assert mod.info.user.err_ramp <= 1e-6
    mod.close()


def test_warm_start_solve_matches_oracle(tmp_path, case9_grid):
    """solve_mpacopf's default warm start: every period solved alone first (keeps the line-limit multipliers),
    then init_solution! and the coupled solve (solve_mpacopf.jl:27-35)."""
    scales = [1.0, 0.975, 1.025]          # the 2 % ramp limit binds for two generators between periods 2 and 3
    prefix = write_profile(tmp_path, case9_grid, scales)
    env, mod = solve_mpacopf(ea.CASE9, prefix, use_gpu=True, verbose=0, end_period=3, outer_iterlim=25, outer_eps=2e-5)
    par = Parameters(); par.verbose = 0; par.outer_iterlim = 25; par.outer_eps = 2e-5
    om = OracleMpModel(case9_grid, par, 4e2, 4e4, np.outer(scales, case9_grid.Pd), np.outer(scales, case9_grid.Qd))
    for m in om.models:
        m.admm_two_level()
    om.init_solution(4e2, 4e4)
    oinfo = om.admm_two_level()
    assert (mod.info.outer, mod.info.cumul) == (oinfo.outer, oinfo.cumul)
    assert mod.info.status == "Solved" and oinfo.status == 2
    assert mod.info.objval == pytest.approx(oinfo.objval, rel=1e-8)
    _compare_state(mod, om, 1e-6, "final state", beta=env.params.beta)
    # the ramp constraint is what couples the periods: it holds to the ADMM tolerance and binds somewhere
    pg = np.array([m.solution.u_curr[0:2 * mod.ngen:2] for m in mod.models])
    ramp = 0.02 * np.asarray(case9_grid.pgmax)
    step = np.abs(np.diff(pg, axis=0))
    assert np.all(step <= ramp + 2e-4) and np.any(step >= 0.999 * ramp)
    assert mod.info.user.err_ramp == pytest.approx(om.err_ramp, abs=1e-7)
    mod.close()


def test_period_solo_solves_are_plain_single_period_solves(tmp_path):
    """The warm start of solve_mpacopf (solve_mpacopf.jl:27-32) solves every period ALONE as a ModelAcopf: the period
    handles of a multi-period model must not see the ramp coupling. Grid with binding line limits, so the multipliers
    the solo solves leave in membuf (rows 25-27) matter for the coupled solve that follows."""
    case = synthetic_case(120, 25, 170, seed=120, rate_margin=1.02)
    grid = ea.GridData.from_opfdata(case)
    scales = [1.0, 1.02, 0.98]
    env, mod, om = make_models(tmp_path, case, grid, scales, outer_iterlim=3, inner_iterlim=60)
    for t, (m, o) in enumerate(zip(mod.models, om.models)):
        admm_two_level(env, m, None, mode="native")
        oi = o.admm_two_level()
        assert m.info.cumul > 0 and (m.info.outer, m.info.cumul) == (oi.outer, oi.cumul), (t, vars(m.info))
        assert m.info.objval == pytest.approx(oi.objval, rel=1e-8)
        np.testing.assert_allclose(m.solution.u_curr, o.vec("u_curr"), atol=1e-6, err_msg=f"period {t} u")
        # line-limit multipliers and penalties the solo solve leaves behind (a multiplier of a limit that is active
        # only within the AL tolerance may differ in size; the penalty ladder is the same but for a handful of branches)
        np.testing.assert_allclose(m.membuf[24:26], o.membuf()[24:26], rtol=1e-3, atol=2e-3, err_msg=f"period {t} lambda_s")
        assert np.mean(m.membuf[26] != o.membuf()[26]) <= 0.02, f"period {t} mu"
        assert np.abs(o.membuf()[24:26]).max() > 0 or t > 0        # some line limit is active
    mod.close()


def test_single_period_horizon_equals_acopf(tmp_path, case9_grid):
    prefix = write_profile(tmp_path, case9_grid, [1.0])
    env, mod = solve_mpacopf(ea.CASE9, prefix, use_gpu=True, verbose=0, end_period=1, warm_start=False, outer_iterlim=25,
                             outer_eps=2e-5)
    env1, mod1 = solve_acopf(ea.CASE9, use_gpu=True, verbose=0, outer_iterlim=25, outer_eps=2e-5)
    assert (mod.info.outer, mod.info.cumul) == (mod1.info.outer, mod1.info.cumul)
    assert mod.info.objval == pytest.approx(mod1.info.objval, rel=1e-12)
    np.testing.assert_allclose(mod.models[0].solution.u_curr, mod1.solution.u_curr, atol=1e-12)
    mod.close(); mod1.close()


def test_synthetic_grid_six_periods_matches_oracle(tmp_path):
    """A 300-bus synthetic grid, 6 periods with a +-3 % load swing: fused solve vs the oracle."""
    case = synthetic_case(300, 60, 420, seed=300)
    grid = ea.GridData.from_opfdata(case)
    scales = [1.0, 1.01, 1.03, 1.02, 0.99, 0.97]
    prefix = write_profile(tmp_path, grid, scales, name="syn")
    env, mod = solve_mpacopf(case, prefix, use_gpu=True, verbose=0, end_period=6, warm_start=False, outer_iterlim=4,
                             inner_iterlim=150, mode="native")
    par = Parameters(); par.verbose = 0; par.outer_iterlim = 4; par.inner_iterlim = 150
    om = OracleMpModel(grid, par, 4e2, 4e4, np.outer(scales, grid.Pd), np.outer(scales, grid.Qd))
    om.set_threads(8)
    oinfo = om.admm_two_level()
    assert (mod.info.outer, mod.info.cumul) == (oinfo.outer, oinfo.cumul)
    assert mod.info.objval == pytest.approx(oinfo.objval, rel=1e-7)
    assert mod.info.mismatch == pytest.approx(oinfo.mismatch, rel=1e-5)
    _compare_state(mod, om, 1e-6, "final state", beta=env.params.beta)
    mod.close()


def test_error_paths(tmp_path, case9_grid):
    prefix = write_profile(tmp_path, case9_grid, [1.0, 1.0])
    env = AdmmEnv(ea.CASE9, 4e2, 4e4, use_gpu=True, verbose=0, load_prefix=prefix)
    with pytest.raises(ValueError):
        ModelMpacopf(env, start_period=1, end_period=3)            # beyond the profile (mpacopf_model.jl:60)
    env2 = AdmmEnv(ea.CASE9, 4e2, 4e4, use_gpu=True, verbose=0)
    with pytest.raises(ValueError):
        ModelMpacopf(env2, start_period=1, end_period=1)           # no load profile (:59)
    mod = ModelMpacopf(env, start_period=1, end_period=2)
    with pytest.raises(ea.capi.EaError):
        mod.get_ramp_vector(5, "u_curr")
    mod.close()
