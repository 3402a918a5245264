"""Pin the CPU oracle to the reference's own golden vectors and known answers
(test/algorithms/acopf_update_cpu.jl:17-172). Operator order is the reference's."""
import numpy as np

from exaadmm_b200.environment import Parameters
from oracle.oracle import OracleModel


def _one_iteration(m):
    m.admm_increment_outer()
    m.admm_outer_prestep()
    m.admm_increment_reset_inner()
    m.admm_increment_inner()
    m.admm_inner_prestep()


def test_one_iteration_matches_reference_goldens(case9_grid, golden, golden_params):
    atol = golden["atol"]
    m = OracleModel(case9_grid, golden_params, 4e2, 4e4)
    _one_iteration(m)
    m.admm_update_x()
    u = m.vec("u_curr")
    np.testing.assert_allclose(u[:6], golden["U_GEN"], atol=atol, rtol=0)
    np.testing.assert_allclose(u[6:], golden["U_BR"], atol=atol, rtol=0)
    m.admm_update_xbar()
    v = m.vec("v_curr")
    np.testing.assert_allclose(v[:6], golden["V_GEN"], atol=atol, rtol=0)
    np.testing.assert_allclose(v[6:], golden["V_BR"], atol=atol, rtol=0)
    m.admm_update_z()
    z = m.vec("z_curr")
    np.testing.assert_allclose(z[:6], golden["Z_GEN"], atol=atol, rtol=0)
    np.testing.assert_allclose(z[6:], golden["Z_BR"], atol=atol, rtol=0)
    m.admm_update_l()
    l = m.vec("l_curr")
    np.testing.assert_allclose(l[:6], golden["L_GEN"], atol=atol, rtol=0)
    np.testing.assert_allclose(l[6:], golden["L_BR"], atol=atol, rtol=0)
    # residual definitions (:154-158)
    m.admm_update_residual()
    np.testing.assert_allclose(m.vec("rp"), u - v + z, atol=atol)
    np.testing.assert_allclose(m.vec("rd"), z - m.vec("z_prev"), atol=atol)
    np.testing.assert_allclose(m.vec("Ax_plus_By"), u - v, atol=atol)
    # lz update (:160-166)
    lz_prev = m.vec("lz").copy()
    m.admm_update_lz()
    np.testing.assert_allclose(m.vec("lz"), lz_prev + golden_params.beta * z, atol=atol)


def test_case9_full_solve_known_answer(case9_grid, golden):
    pin = golden["solve_case9"]
    par = Parameters()
    par.verbose = 0
    par.outer_iterlim = pin["kwargs"]["outer_iterlim"]
    par.outer_eps = pin["kwargs"]["outer_eps"]
    m = OracleModel(case9_grid, par, pin["kwargs"]["rho_pq"], pin["kwargs"]["rho_va"])
    info = m.admm_two_level()
    assert info.status == 2                       # :Solved
    assert info.outer == pin["outer"]
    assert info.cumul == pin["cumul"]
    assert abs(info.objval - pin["objval"]) <= pin["objval_atol"]
    c = m.counters()
    assert c["line_calls"] == 9 * 705 and c["max_auglag_hits"] == 0


def test_case9_iteration_trace_regression(case9_grid):
    """Secondary, self-generated pin: inner iterations per outer
    iteration and the beta schedule of the run above."""
    expect = [20, 68, 159, 134, 107, 81, 55, 32, 16, 8, 4, 4, 3, 3, 2, 2, 2, 2, 2, 1]
    par = Parameters()
    par.verbose = 0
    par.outer_iterlim = 25
    par.outer_eps = 2e-5
    m = OracleModel(case9_grid, par, 4e2, 4e4)
    import math
    sqrt_d = math.sqrt(m.nvar)
    counts, betas = [], []
    par.beta = par.initial_beta
    while m.outer < par.outer_iterlim:
        m.admm_increment_outer(); m.admm_outer_prestep(); m.admm_increment_reset_inner()
        betas.append(par.beta)
        while m.inner < par.inner_iterlim:
            res = m.inner_iteration()
            if res[0] <= sqrt_d / (2500 * m.outer):
                break
        counts.append(m.inner)
        if res[3] <= sqrt_d * par.outer_eps:
            break
        m.admm_update_lz()
        if res[2] > par.theta * m.norm_z_prev:
            par.beta = min(par.inc_c * par.beta, 1e24)
    assert counts == expect
    assert betas[:3] == [1e3, 6e3, 3.6e4] and betas[-1] == 4.6656e7
