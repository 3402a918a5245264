"""The dataflow inner loop (csrc/async.cuh: one persistent kernel, branches and buses as a task graph, look-ahead over
a ring of iteration slots) against the synchronous fused loop: same iterates, same stopping iteration, same solve."""
import ctypes as C

import numpy as np
import pytest

import exaadmm_b200 as ea
from exaadmm_b200 import operators as ops
from exaadmm_b200.admm_two_level import admm_two_level
from exaadmm_b200.environment import AdmmEnv
from exaadmm_b200.model import ModelAcopf
from exaadmm_b200.synthetic import synthetic_case

pytestmark = pytest.mark.gpu
FIELDS = ("u_curr", "v_curr", "z_curr", "z_prev", "l_curr", "rp", "rd", "Ax_plus_By")


def _model(case, rho_pq, rho_va, async_mode, tight_factor=1.0, **params):
    env = AdmmEnv(case, rho_pq, rho_va, use_gpu=True, verbose=0, tight_factor=tight_factor)
    for k, v in params.items():
        setattr(env.params, k, v)
    mod = ModelAcopf(env)
    mod.set_option("async", 1.0 if async_mode else 0.0)
    mod.set_option("async_budget_s", 20.0)
    return env, mod


def _run_inner(env, mod, outer, n):
    env.params.inner_iterlim = n
    mod.info.outer = outer
    ops.admm_increment_reset_inner(env, mod)
    ops.admm_outer_prestep(env, mod)
    ops.admm_run_inner(env, mod)
    return mod.info.inner


@pytest.mark.parametrize("case,rho,n", [("case9", (4e2, 4e4), 25), ("syn300", (4e2, 4e4), 40), ("syn300tight", (4e2, 4e4), 40)])
def test_same_iterates_as_the_synchronous_loop(case, rho, n):
    if case == "case9":
        data, tf = ea.CASE9, 1.0
    else:
        data, tf = synthetic_case(300, 40, 420, 5), (0.3 if case.endswith("tight") else 0.99)
    state = {}
    for mode in (0, 1):
        env, mod = _model(data, rho[0], rho[1], mode, tight_factor=tf)
        # outer = 10^9: eps_pri is never met, exactly n iterations; then 3 more from where that run stopped
        assert _run_inner(env, mod, 10 ** 9, n) == n
        first = {f: getattr(mod.solution, f) for f in FIELDS}
        first["membuf"] = mod.membuf[24:27]
        env.params.inner_iterlim = n + 3
        done = C.c_int64()
        out = np.zeros(4)
        mod._check(mod.lib.ea_run_inner_from(mod.h, 10 ** 9, env.params.beta, n, n + 3, env.params.max_auglag, env.params.mu_max,
                                             env.params.scale, 0, C.byref(done), out.ctypes.data_as(C.POINTER(C.c_double))))
        assert done.value == n + 3
        state[mode] = (first, {f: getattr(mod.solution, f) for f in FIELDS}, out)
        mod.close()
    for k in (0, 1):
        for f in state[0][k]:
            np.testing.assert_allclose(state[1][k][f], state[0][k][f], rtol=0, atol=1e-11, err_msg=f"{case} {f} run {k}")
    np.testing.assert_allclose(state[1][2], state[0][2], rtol=1e-11)


def test_device_side_termination_matches():
    """With a real eps_pri the dataflow run stops on the same iteration as the synchronous one (look-ahead work past
    the stopping iteration is discarded), and the state is that iteration's."""
    data = synthetic_case(300, 40, 420, 5)
    got = {}
    for mode in (0, 1):
        env, mod = _model(data, 4e2, 4e4, mode, tight_factor=0.99)
        inner = _run_inner(env, mod, 1, 1000)
        got[mode] = (inner, mod.info.primres, mod.solution.u_curr, mod.solution.z_curr, mod.solution.z_prev, mod.solution.l_curr)
        mod.close()
    assert got[0][0] == got[1][0] and 1 < got[0][0] < 1000
    assert abs(got[0][1] - got[1][1]) <= 1e-11 * max(1.0, got[0][1])
    for a, b in zip(got[0][2:], got[1][2:]):
        np.testing.assert_allclose(b, a, rtol=0, atol=1e-11)


@pytest.mark.parametrize("case", ["case9", "syn1354"])
def test_full_solve_same_answer(case):
    res = {}
    for mode in (0, 1):
        if case == "case9":
            env, mod = _model(ea.CASE9, 4e2, 4e4, mode, outer_iterlim=25, outer_eps=2e-5)
        else:
            env, mod = _model(synthetic_case(1354, 260, 1991, 1354), 4e2, 4e4, mode, tight_factor=0.99, outer_iterlim=6)
        admm_two_level(env, mod, None, mode="native")
        res[mode] = (mod.info.status, mod.info.outer, mod.info.cumul, mod.info.objval, mod.info.mismatch)
        mod.close()
    assert res[0][:3] == res[1][:3]
    assert abs(res[0][3] - res[1][3]) <= 1e-9 * abs(res[0][3])
    if case == "case9":
        assert res[1][:3] == ("Solved", 20, 705) and abs(res[1][3] - 5303.435) <= 1e-3
