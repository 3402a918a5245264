"""One-level ADMM on the SQP sub-problem (`ModelQpsub`) on the CUDA path vs the reference's goldens and the oracle,
through the C ABI (run on the B200 box). Mirrors test/algorithms/qpsub_update_gpu.jl: operator-level known answers on
case9, then the end-to-end solve (Solved / 5107 / 5107 / -21.92744641968529) and what it hands back to the SQP driver."""
import sys
from pathlib import Path

import numpy as np
import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent))
import exaadmm_b200 as ea  # noqa: E402
import qpsub_setup  # noqa: E402
from exaadmm_b200 import operators as ops  # noqa: E402
from exaadmm_b200 import qpsub  # noqa: E402
from exaadmm_b200.capi import EaError  # noqa: E402
from exaadmm_b200.environment import AdmmEnv, Parameters  # noqa: E402
from oracle.oracle import OracleQpModel  # noqa: E402

pytestmark = pytest.mark.gpu

GOLD = qpsub_setup.load_golden()
ITERATE_TOL = 1e-8
FIELDS = ("Hs", "LH_1h", "RH_1h", "LH_1i", "RH_1i", "LH_1j", "RH_1j", "LH_1k", "RH_1k", "ls", "us")
GEN = (("qpsub_pgmax", "pgmax"), ("qpsub_pgmin", "pgmin"), ("qpsub_qgmax", "qgmax"), ("qpsub_qgmin", "qgmin"),
       ("qpsub_c1", "c1"), ("qpsub_c2", "c2"), ("qpsub_Pd", "Pd"), ("qpsub_Qd", "Qd"))


def make_model(case, grid, data, rho_pq, rho_va, **params):
    """env2 / mod2 of the reference's test: construct, copy the QP fields in, init_solution!."""
    env = AdmmEnv(case, rho_pq, rho_va, use_gpu=True, verbose=0)
    mod = ea.ModelQpsub(env, grid=grid)
    for k in FIELDS:
        setattr(mod, k, np.array(getattr(data, k)))
    for k, src in GEN:
        setattr(mod, k, np.array(getattr(data, src)))
    for k, v in params.items():
        setattr(env.params, k, v)
    qpsub.init_solution(mod, mod.solution, env.initial_rho_pq, env.initial_rho_va)
    return env, mod


def oracle_model(grid, data, rho_pq, rho_va, **params):
    par = Parameters(); par.verbose = 0
    for k, v in params.items():
        setattr(par, k, v)
    return OracleQpModel(grid, par, data, rho_pq, rho_va)


@pytest.fixture(scope="module")
def case9_data(case9_grid):
    return qpsub_setup.linearise(case9_grid, GOLD["sqp_point"])


def test_operator_level_goldens_case9(case9_grid, case9_data):
    """qpsub_update_gpu.jl:150-225."""
    p = GOLD["params_one_iteration"]
    atol = p["atol"]
    env, mod = make_model(ea.CASE9, case9_grid, case9_data, p["rho_pq"], p["rho_va"], scale=p["scale"], initial_beta=0, beta=0)
    ops.admm_increment_outer(env, mod); ops.admm_increment_reset_inner(env, mod); ops.admm_increment_inner(env, mod)
    ops.admm_update_x(env, mod)
    assert np.abs(mod.solution.u_curr - GOLD["U_SOL"]).max() <= atol
    ops.admm_update_xbar(env, mod)
    assert np.abs(mod.solution.v_curr - GOLD["V_SOL"]).max() <= atol
    ops.admm_update_l_single(env, mod)
    assert np.abs(mod.solution.l_curr - GOLD["L_SOL"]).max() <= atol
    ops.admm_update_residual(env, mod)
    assert np.abs(mod.solution.rp - GOLD["RP_SOL"]).max() <= atol
    assert np.abs(mod.solution.rd - GOLD["RD_SOL"]).max() <= atol
    assert np.array_equal(mod.solution.Ax_plus_By, mod.solution.rp)
    assert mod.info.mismatch == mod.info.primres > 0


def _lockstep(env, mod, om, n_iter, tol=ITERATE_TOL):
    worst = 0.0
    for it in range(n_iter):
        for m in (mod.info, om):
            m.inner = 1
        ops.admm_update_x(env, mod); om.admm_update_x()
        worst = max(worst, np.abs(mod.solution.u_curr - om.vec("u_curr")).max(),
                    np.abs(mod.sqp_line - om.line_array("sqp_line")).max())
        np.testing.assert_array_equal(mod.qpsub_membuf[4], om.line_array("qpsub_membuf")[4])       # penalty ladder
        np.testing.assert_allclose(mod.qpsub_membuf[2:4], om.line_array("qpsub_membuf")[2:4], rtol=1e-7, atol=1e-7)
        ops.admm_update_xbar(env, mod); om.admm_update_xbar()
        worst = max(worst, np.abs(mod.solution.v_curr - om.vec("v_curr")).max(), np.abs(mod.v_prev - om.vec("v_prev")).max())
        ops.admm_update_l_single(env, mod); om.admm_update_l_single()
        scale_l = max(1.0, np.abs(om.vec("l_curr")).max())
        worst = max(worst, np.abs(mod.solution.l_curr - om.vec("l_curr")).max() / scale_l)
        ops.admm_update_residual(env, mod)
        r = om.admm_update_residual()
        got = np.array([mod.info.primres, mod.info.dualres, mod.info.mismatch, mod.info.objval, mod.info.auglag])
        np.testing.assert_allclose(got, r, rtol=1e-9, atol=1e-9 * max(1.0, np.abs(r).max()))
        assert worst <= tol, (it, worst)
    return worst


def test_iterates_match_oracle_case9(case9_grid, case9_data):
    env, mod = make_model(ea.CASE9, case9_grid, case9_data, 4000.0, 4000.0, scale=1e-4)
    om = oracle_model(case9_grid, case9_data, 4000.0, 4000.0, scale=1e-4)
    _lockstep(env, mod, om, 25)
    c, oc = mod.counters(), om.counters()
    assert (c["line_calls"], c["auglag_iters"], c["tron_evals"]) == (oc["line_calls"], oc["auglag_iters"], oc["tron_evals"])


def test_iterates_match_oracle_binding_line_limits():
    """Synthetic grid with tight limits: several AL iterations per branch, multipliers and penalties move."""
    grid, data = qpsub_setup.synthetic_qpsub(40, 8, 56, seed=11, tight_factor=0.05, spread=0.06)
    env, mod = make_model(ea.CASE9, grid, data, 400.0, 400.0, scale=1e-4)
    om = oracle_model(grid, data, 400.0, 400.0, scale=1e-4)
    _lockstep(env, mod, om, 10, tol=1e-7)
    c, oc = mod.counters(), om.counters()
    assert c["auglag_iters"] == oc["auglag_iters"] > 10 * grid.nline
    assert c["max_auglag_one_call"] == oc["max_auglag_one_call"] >= 3


def _solve_kwargs():
    kw = dict(GOLD["solve"]["kwargs"])
    beta = kw.pop("initial_beta")
    return beta, kw


@pytest.mark.parametrize("mode", ["native", "stepwise"])
def test_solve_qpsub_known_answer(case9_data, mode):
    """qpsub_update_gpu.jl:228-346: Solved / 5107 / 5107 / -21.92744641968529 and everything the SQP driver reads back."""
    s = GOLD["solve"]
    d = case9_data
    beta, kw = _solve_kwargs()
    env, mod = ea.solve_qpsub(ea.CASE9, d.Hs, d.LH_1h, d.RH_1h, d.LH_1i, d.RH_1i, d.LH_1j, d.RH_1j, d.LH_1k, d.RH_1k,
                              d.ls, d.us, d.pgmax, d.pgmin, d.qgmax, d.qgmin, d.c1, d.c2, d.Pd, d.Qd, beta,
                              verbose=0, onelevel=True, use_gpu=True, mode=mode, **kw)
    assert mod.info.status == "Solved"
    assert (mod.info.outer, mod.info.cumul) == (s["outer"], s["cumul"])
    assert abs(mod.info.objval - s["objval"]) <= s["objval_atol"]
    for k in ("dpg_sol", "dqg_sol", "dtheta_sol", "dw_sol", "dual_infeas"):
        assert np.abs(getattr(mod, k) - np.asarray(s[k])).max() <= s["atol"], k
    for k in ("dline_var", "dline_fl"):                      # printed with 6 significant digits in the reference's file
        ref = np.asarray(s[k])
        assert np.abs(getattr(mod, k) - ref).max() <= s["atol"] + 5e-6 * np.abs(ref).max(), k
    lam = np.asarray(s["lambda"])
    assert (np.abs(lam - mod.lambda_) / np.maximum(np.abs(lam), 1)).max() <= 5e-6
    if mode == "native":
        kt = mod.kernel_times()
        assert kt["iterations"] == s["outer"] and kt["graph_replays"] >= s["outer"] // 32


def test_native_loop_equals_stepwise_loop_and_oracle():
    """Synthetic 300-bus grid: the fused loop (CUDA graph, device-side termination) stops on the same iteration as the
    operator-by-operator loop and as the oracle, with the same state."""
    grid, data = qpsub_setup.synthetic_qpsub(300, 40, 420, seed=5, tight_factor=0.2)
    runs = {}
    for mode in ("native", "stepwise"):
        env, mod = make_model(ea.CASE9, grid, data, 4000.0, 4000.0, scale=1e-4, outer_eps=2e-5, outer_iterlim=400)
        ea.admm_one_level(env, mod, None, mode=mode)
        runs[mode] = (mod.info.status, mod.info.outer, mod.info.objval, mod.solution.u_curr, mod.solution.v_curr,
                      mod.solution.l_curr, mod.sqp_line, mod.solution.rp, mod.solution.rd)
    a, b = runs["native"], runs["stepwise"]
    assert a[:2] == b[:2]
    for x, y in zip(a[2:], b[2:]):
        np.testing.assert_array_equal(x, y)                  # same kernels, same order: bitwise
    om = oracle_model(grid, data, 4000.0, 4000.0, scale=1e-4, outer_eps=2e-5, outer_iterlim=400)
    oi = om.admm_one_level()
    assert (ea.capi.STATUS_NAMES[oi.status], oi.outer) == a[:2]
    assert abs(oi.objval - a[2]) <= 1e-8 * max(1.0, abs(oi.objval))
    assert np.abs(om.vec("u_curr") - a[3]).max() <= 1e-7


def test_iteration_limit_and_chunks(case9_grid, case9_data):
    """The loop stops exactly at outer_iterlim, whatever the chunk size, with and without the graph."""
    ref = None
    for chunk, graph in ((32, 1), (7, 1), (5, 0), (1, 0)):
        env, mod = make_model(ea.CASE9, case9_grid, case9_data, 4000.0, 4000.0, scale=1e-4, outer_eps=2e-6, outer_iterlim=45)
        mod.set_option("chunk", chunk); mod.set_option("use_graph", graph)
        ea.admm_one_level(env, mod, None, mode="native")
        assert (mod.info.status, mod.info.outer, mod.info.cumul) == ("IterationLimit", 45, 45)
        u = mod.solution.u_curr
        if ref is None:
            ref = u
        np.testing.assert_array_equal(u, ref)


def test_error_paths(case9_grid, case9_data):
    env = AdmmEnv(ea.CASE9, 20.0, 20.0, use_gpu=True, verbose=0)
    mod = ea.ModelQpsub(env, grid=case9_grid)
    with pytest.raises(RuntimeError):                         # no init_solution! yet
        ops.admm_update_x(env, mod)
    with pytest.raises(EaError):                              # all-zero QP: 1h / 1i cannot be eliminated
        qpsub.init_solution(mod, mod.solution, 20.0, 20.0)
    env, mod = make_model(ea.CASE9, case9_grid, case9_data, 20.0, 20.0)
    with pytest.raises(EaError):
        mod._check(mod.lib.ea_qp_update_x(mod.h, 0, 50, 1e8, 1e-4))     # info.inner must be >= 1
    bad = np.array(case9_data.Hs); bad[0, 1] += 1.0
    mod.Hs = bad
    with pytest.raises(EaError):                              # non-symmetric Hessian block
        qpsub.init_solution(mod, mod.solution, 20.0, 20.0)
    with pytest.raises(TypeError):
        ops.admm_update_l_single(env, object())
    with pytest.raises(NotImplementedError):
        ea.solve_qpsub(ea.CASE9, *([None] * 19), 0.0, use_gpu=False)
