"""Pin the multi-period CPU oracle to the reference's golden vectors and known answer
(test/algorithms/mpacopf_update_cpu.jl:1-435). Operator order is the reference's."""
import json
from pathlib import Path

import numpy as np
import pytest

from exaadmm_b200.environment import Parameters
from oracle.oracle import OracleMpModel

GOLD = json.loads((Path(__file__).resolve().parent / "golden" / "mpacopf_case9_golden.json").read_text())


def mp_loads(grid, scales):
    Pd = np.array([s * np.asarray(grid.Pd) for s in scales])
    Qd = np.array([s * np.asarray(grid.Qd) for s in scales])
    return Pd, Qd


@pytest.fixture()
def mp_model(case9_grid, golden_params):
    Pd, Qd = mp_loads(case9_grid, GOLD["load_scale"])
    return OracleMpModel(case9_grid, golden_params, 4e2, 4e4, Pd, Qd, ramp_ratio=GOLD["params"]["ramp_ratio"])


def test_one_iteration_matches_reference_goldens(mp_model, golden_params):
    m, atol, T = mp_model, GOLD["atol"], GOLD["len_horizon"]
    assert m.nvar == m.models[0].nvar + m.ngen                    # mpacopf_model.jl:97-102
    m.admm_increment_outer()
    m.admm_outer_prestep()
    m.admm_increment_reset_inner()
    m.admm_increment_inner()
    m.admm_inner_prestep()

    def check(field, gen, br):
        for t in range(T):
            x = m.models[t].vec(field)
            np.testing.assert_allclose(x[:6], GOLD[gen][t], atol=atol, rtol=0, err_msg=f"{gen} t={t}")
            np.testing.assert_allclose(x[6:], GOLD[br][t], atol=atol, rtol=0, err_msg=f"{br} t={t}")

    m.admm_update_x()
    check("u_curr", "U_GEN", "U_BR")
    m.admm_update_xbar()
    check("v_curr", "V_GEN", "V_BR")
    m.admm_update_z()
    check("z_curr", "Z_GEN", "Z_BR")
    m.admm_update_l()
    check("l_curr", "L_GEN", "L_BR")
    # residual definitions (:385-407)
    m.admm_update_residual()
    for t in range(T):
        s = m.models[t]
        np.testing.assert_allclose(s.vec("rp"), s.vec("u_curr") - s.vec("v_curr") + s.vec("z_curr"), atol=atol)
        np.testing.assert_allclose(s.vec("rd"), s.vec("z_curr") - s.vec("z_prev"), atol=atol)
        np.testing.assert_allclose(s.vec("Ax_plus_By"), s.vec("u_curr") - s.vec("v_curr"), atol=atol)
        if t > 0:
            vp = m.models[t - 1].vec("v_curr")[0:2 * m.ngen:2]
            np.testing.assert_allclose(m.ramp(t, "rp"), m.ramp(t, "u_curr") - vp + m.ramp(t, "z_curr"), atol=atol)
            np.testing.assert_allclose(m.ramp(t, "rd"), m.ramp(t, "z_curr") - m.ramp(t, "z_prev"), atol=atol)
            np.testing.assert_allclose(m.ramp(t, "Ax_plus_By"), m.ramp(t, "u_curr") - vp, atol=atol)
    # lz update (:410-428)
    lz_prev = [m.models[t].vec("lz").copy() for t in range(T)]
    lzr_prev = [m.ramp(t, "lz").copy() for t in range(T)]
    m.admm_update_lz()
    for t in range(T):
        np.testing.assert_allclose(m.models[t].vec("lz"), lz_prev[t] + golden_params.beta * m.models[t].vec("z_curr"),
                                   atol=atol)
        if t > 0:
            np.testing.assert_allclose(m.ramp(t, "lz"), lzr_prev[t] + golden_params.beta * m.ramp(t, "z_curr"), atol=atol)


def test_case9_three_periods_known_answer(case9_grid):
    pin = GOLD["solve_case9_T3"]
    par = Parameters()
    par.verbose = 0
    par.outer_iterlim = pin["kwargs"]["outer_iterlim"]
    par.outer_eps = pin["kwargs"]["outer_eps"]
    Pd, Qd = mp_loads(case9_grid, GOLD["load_scale"])
    m = OracleMpModel(case9_grid, par, pin["kwargs"]["rho_pq"], pin["kwargs"]["rho_va"], Pd, Qd)
    info = m.admm_two_level()
    assert info.status == 2                       # :Solved
    assert (info.outer, info.cumul) == (pin["outer"], pin["cumul"])
    assert abs(info.objval - pin["objval"]) <= pin["objval_atol"]
    assert m.err_ramp <= 1e-6
