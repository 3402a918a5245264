"""The product's device routines (csrc/tron.cuh, branch.cuh) compiled for the host
and driven against the oracle, call by call — no GPU needed. This is what lets
the CPU suite cover the flattened AL/TRON loop and the masked free-set logic."""
import ctypes as C
import math

import numpy as np

import exaadmm_b200 as ea
from exaadmm_b200.environment import Parameters
from exaadmm_b200.synthetic import synthetic_case
from oracle import oracle as orc
from oracle.oracle import OracleModel
from conftest import branch_inputs
from test_oracle_units import _random_problem

pd = C.POINTER(C.c_double)


def P(a):
    return a.ctypes.data_as(pd)


def test_device_eval_matches_oracle(host_harness):
    rng = np.random.default_rng(3)
    for _ in range(300):
        x, _, _, p, Y = _random_problem(rng, binding=True)
        f = C.c_double(); g = np.zeros(6); H = np.zeros(36)
        host_harness.hh_eval(P(x), P(p), P(Y), 1e-4, C.byref(f), P(g), P(H))
        fo = orc.eval_f(x, p, Y, 1e-4); go, Ho = orc.eval_gh(x, p, Y, 1e-4)
        assert abs(f.value - fo) <= 1e-13 * max(1.0, abs(fo))
        np.testing.assert_allclose(g, go, rtol=1e-12, atol=1e-13 * max(1.0, np.abs(go).max()))
        np.testing.assert_allclose(H.reshape(6, 6), Ho, rtol=1e-12, atol=1e-13 * max(1.0, np.abs(Ho).max()))


def _run_lockstep(host_harness, grid, par, rho_pq, rho_va, n_iter, tol):
    """Oracle drives the ADMM iteration; every branch solve is replayed through the
    device code with identical inputs and must give the same u, lambda_s, mu and
    the same number of evaluations."""
    m = OracleModel(grid, par, rho_pq, rho_va)
    nl, ng = grid.nline, grid.ngen
    m.admm_increment_outer(); m.admm_outer_prestep(); m.admm_increment_reset_inner()
    worst = 0.0
    for _ in range(n_iter):
        m.admm_increment_inner(); m.admm_inner_prestep(); m.admm_update_x_gen()
        u = m.vec("u_curr").copy(); v = m.vec("v_curr").copy(); z = m.vec("z_curr").copy()
        l = m.vec("l_curr").copy(); rho = m.vec("rho").copy(); mb = m.membuf().copy()
        m.reset_counters()
        m.admm_update_x_line()
        oc = m.counters()
        u2 = m.vec("u_curr"); mb2 = m.membuf()
        evals = 0
        for I in range(nl):
            x, xl, xu, param, Y = branch_inputs(grid, u, v, z, l, rho, mb, I)
            F = np.zeros(4); work = (C.c_int * 6)()
            host_harness.hh_solve_branch(P(x), P(xl), P(xu), P(param), P(Y), m.inner, par.max_auglag, par.mu_max,
                                         par.scale, P(F), work)
            p = 2 * ng + 8 * I
            uo = np.array([F[0], F[1], F[2], F[3], x[0] * x[0], x[1] * x[1], x[2], x[3]])
            worst = max(worst, np.abs(uo - u2[p:p + 8]).max())
            assert param[26] == mb2[26, I]                                     # mu: exact (powers of ten)
            worst = max(worst, np.abs(param[24:26] - mb2[24:26, I]).max() / max(1.0, param[26]))
            evals += work[1]
        # Lines whose AL loop runs into max_auglag sit at mu = 1e8, where the sub-problem is so
        # ill-conditioned that 1-ulp differences (FMA contraction) change TRON's iteration count;
        # the solutions still agree (checked above). Everywhere else the counts are identical.
        if oc["max_auglag_hits"] == 0:
            assert evals == oc["tron_evals"]
        m.admm_update_xbar(); m.admm_update_z(); m.admm_update_l(); m.admm_update_residual()
    assert worst <= tol, worst


def test_branch_solver_lockstep_case9(host_harness, case9_grid):
    par = Parameters(); par.verbose = 0
    _run_lockstep(host_harness, case9_grid, par, 4e2, 4e4, 40, 1e-11)


def test_branch_solver_lockstep_synthetic_with_binding_limits(host_harness):
    d = synthetic_case(60, 12, 84, seed=60, rate_margin=1.02)    # ratings just above the construction flows -> limits bind
    grid = ea.GridData.from_opfdata(d, tight_factor=0.99)
    par = Parameters(); par.verbose = 0
    _run_lockstep(host_harness, grid, par, 4e2, 4e4, 25, 1e-9)


def test_branch_solver_lockstep_unattainable_limits(host_harness):
    """Ratings below the construction flows: some lines run the AL loop into
    max_auglag at mu = 1e8. Counts may differ there (ill-conditioning), results may not."""
    d = synthetic_case(60, 12, 84, seed=60, rate_margin=0.8)
    grid = ea.GridData.from_opfdata(d, tight_factor=0.99)
    par = Parameters(); par.verbose = 0
    _run_lockstep(host_harness, grid, par, 4e2, 4e4, 25, 1e-7)


def test_tron_logic_is_bit_identical_to_oracle_on_hard_problems(host_harness_nofma):
    """Device AL/TRON loop with the oracle's f/g/H plugged in and FMA contraction off:
    the arithmetic is then the oracle's, so iterates and evaluation counts must agree
    EXACTLY — through Cholesky shifts, negative curvature, rejected steps, Cauchy
    extrapolation and the masked free set."""
    L = orc.lib()
    fn = C.cast(L.orc_eval_f, C.c_void_p); ghn = C.cast(L.orc_eval_gh, C.c_void_p)
    rng = np.random.default_rng(11)
    shifts = rejected = 0
    for k in range(1500):
        x0, xl, xu, p, Y = _random_problem(rng, binding=bool(k % 2))
        if k % 3:
            p[8:16] *= 0.01                               # weak penalties: indefinite Hessians
        if k % 7 == 0:
            xl[2] = xu[2] = x0[2] = 0.0                   # reference-bus end
        pc = p.copy(); xc = x0.copy(); work = (C.c_int * 6)()
        host_harness_nofma.hh_solve_branch_oracle_eval(fn, ghn, P(xc), P(xl), P(xu), P(pc), P(Y), 2, 1, 1e8, 1e-4, work)
        xo, st, minor, nfev = orc.tron_solve(x0, xl, xu, p, Y, 1e-4)
        assert work[1] == nfev
        np.testing.assert_array_equal(xc, xo)
        shifts += work[3]; rejected += work[4]
    assert shifts > 100 and rejected > 1000               # the hard paths were really exercised


def test_parity_build_arithmetic_is_bit_identical_to_oracle(host_harness_parity):
    """The library's PARITY build (csrc/Makefile: -fmad=false -DEA_NO_FMA -DEA_PARITY), compiled for the host: the
    product's own objective (branch::eval_fgh_ref, portable sin / cos) inside the product's AL / TRON state machine
    against the oracle switched to the same sin / cos - every branch of every iteration must give the same BITS
    (u, lambda_s, mu) and the same evaluation count. On the GPU the same source differs only in where it runs
    (tests/test_gpu_baseline_configs.py compares that build with the oracle over whole iterations)."""
    from oracle.oracle import set_portable_sincos
    set_portable_sincos(True)
    try:
        par = Parameters(); par.verbose = 0
        d = synthetic_case(60, 12, 84, seed=60, rate_margin=1.02)         # limits bind: the penalty ladder is exercised
        grid = ea.GridData.from_opfdata(d, tight_factor=0.99)
        _run_lockstep(host_harness_parity, grid, par, 4e2, 4e4, 25, 0.0)
        d = synthetic_case(300, 40, 420, seed=300)
        grid = ea.GridData.from_opfdata(d, tight_factor=0.99)
        _run_lockstep(host_harness_parity, grid, par, 1e1, 1e3, 12, 0.0)   # README rho of the pegase cases: rejected steps
    finally:
        set_portable_sincos(False)


def test_branch_solver_random_problems_fma_build(host_harness):
    """The GPU-arithmetic build (FMA, rsqrt/reciprocal Cholesky, fused compact objective) on
    random cold-start problems. A single TRON solve is only defined up to its own tolerances
    (gtol, CG truncation), so rounding can change the path: require equally good solutions
    everywhere and the identical path on the vast majority."""
    rng = np.random.default_rng(12)
    same_path = total = 0
    for k in range(300):
        x0, xl, xu, p, Y = _random_problem(rng, binding=bool(k % 2))
        pc = p.copy(); xc = x0.copy()
        F = np.zeros(4); work = (C.c_int * 6)()
        host_harness.hh_solve_branch(P(xc), P(xl), P(xu), P(pc), P(Y), 2, 1, 1e8, 1e-4, P(F), work)   # one TRON solve
        xo, st, minor, nfev = orc.tron_solve(x0, xl, xu, p, Y, 1e-4)
        fd, fo = orc.eval_f(xc, p, Y, 1e-4), orc.eval_f(xo, p, Y, 1e-4)
        assert abs(fd - fo) <= 1e-9 * max(1.0, abs(fo))
        np.testing.assert_allclose(xc, xo, rtol=0, atol=1e-4)
        total += 1
        same_path += (work[1] == nfev) and np.allclose(xc, xo, rtol=0, atol=1e-10)
    assert same_path >= 0.95 * total, (same_path, total)


def test_direct_step_follows_the_literal_algorithm(host_harness, host_harness_literal):
    """tron::newton_step (Cauchy search in closed form, one exact Cholesky factor and one CG step per face, projected
    search only where a variable reaches a bound) against the literal algorithm it short-cuts, on the hardest branches
    of a real solve (tests/golden/hard_branches.npz: penalty ladders, trust-region-limited solves at mu = 1e8, Cholesky
    shifts, rejected steps) plus ordinary ones: same AL iterations everywhere; the same evaluation and CG counts and the
    same solution to 1e-6 on all but the few solves whose path is rounding-sensitive."""
    from pathlib import Path
    d = np.load(Path(__file__).resolve().parent / "golden" / "hard_branches.npz")
    prob, meta = d["prob"], d["meta"]
    n = prob.shape[0]
    res = {}
    for name, H in (("direct", host_harness), ("literal", host_harness_literal)):
        xs = np.zeros((n, 10)); ws = np.zeros((n, 6), dtype=np.int64)
        for i in range(n):
            p = prob[i]
            param = np.zeros(31); param[0:24] = p[0:24]; param[24:27] = p[50:53]
            Y = p[24:32].copy(); xl = p[32:38].copy(); xu = p[38:44].copy(); x = p[44:50].copy()
            F = np.zeros(4); work = (C.c_int * 6)()
            # major = 2: mu comes from param[26] (the fixture stores the mu the solve started with)
            H.hh_solve_branch(P(x), P(xl), P(xu), P(param), P(Y), 2, int(meta[2]), float(meta[1]), float(meta[0]), P(F), work)
            xs[i, :6] = x; xs[i, 6:] = F; ws[i] = list(work)
        res[name] = (xs, ws)
    (xd, wd), (xl_, wl) = res["direct"], res["literal"]
    assert wl[:, 1].max() >= 150 and (wl[:, 0] >= 20).sum() >= 50
    np.testing.assert_array_equal(wd[:, 0], wl[:, 0])                         # AL iterations
    same = (wd[:, 1] == wl[:, 1]) & (wd[:, 2] == wl[:, 2])                    # evaluations and CG iterations
    assert same.mean() >= 0.97, same.mean()
    np.testing.assert_allclose(xd[same], xl_[same], rtol=0, atol=1e-6)
    assert np.abs(xd - xl_).max() <= 1e-4


# ---------------------------------------------------------------------------
# multi-period model: the n = 3 generator sub-problem (genramp.cuh) against the oracle
# ---------------------------------------------------------------------------
def _gen_problems(n, seed):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        pmax = rng.uniform(0.5, 4.0)
        pmin = rng.uniform(0.0, 0.4) * pmax
        ramp = rng.choice([0.02, 0.1, 0.5]) * pmax
        xl = np.array([pmin, pmin, -ramp]); xu = np.array([pmax, pmax, ramp])
        x0 = np.array([rng.uniform(pmin, pmax), rng.uniform(pmin, pmax), rng.uniform(-ramp, ramp)])
        rho = rng.choice([10.0, 400.0, 3e4])
        param = np.array([rng.normal() * 50, rng.normal() * 50, rho, rho, rng.uniform(pmin, pmax) + rng.normal() * 0.05,
                          rng.uniform(pmin, pmax) + rng.normal() * 0.05, rng.normal() * 5.0, rng.choice([10.0, 100.0, 1e4])])
        cost = (rng.uniform(0.01, 0.2), rng.uniform(5, 40), rng.uniform(0, 500))
        out.append((x0, xl, xu, param, cost))
    return out


def _run_gen(fn, x0, xl, xu, param, cost, max_auglag=50, xi_max=1e8, i32=False):
    import ctypes as C
    x, p = x0.copy(), param.copy()
    work = (C.c_int32 * 3)() if i32 else (C.c_int * 3)()
    pd = C.POINTER(C.c_double)
    fn(x.ctypes.data_as(pd), xl.ctypes.data_as(pd), xu.ctypes.data_as(pd), p.ctypes.data_as(pd), cost[0], cost[1], cost[2],
       100.0, 1.0, max_auglag, xi_max, work)
    return x, p, list(work)


def test_generator_ramp_solve_bit_exact_logic(host_harness_nofma):
    """-DEA_NO_FMA build: same arithmetic as the oracle => identical iterates and counts."""
    from oracle.oracle import lib
    L = lib()
    for prob in _gen_problems(300, 11):
        xo, po, wo = _run_gen(L.orc_gen_ramp_solve, *prob, i32=True)
        xh, ph, wh = _run_gen(host_harness_nofma.hh_solve_gen, *prob)
        assert wo == wh, (wo, wh)
        np.testing.assert_array_equal(xo, xh)
        np.testing.assert_array_equal(po, ph)


def test_generator_ramp_solve_fma_build_close(host_harness):
    from oracle.oracle import lib
    L = lib()
    for prob in _gen_problems(300, 12):
        xo, po, wo = _run_gen(L.orc_gen_ramp_solve, *prob, i32=True)
        xh, ph, wh = _run_gen(host_harness.hh_solve_gen, *prob)
        np.testing.assert_allclose(xh, xo, atol=1e-9, rtol=0)
        np.testing.assert_allclose(ph[6:], po[6:], atol=1e-6, rtol=1e-9)
        # the ramp equality holds to the AL tolerance
        assert abs(xh[0] - xh[1] - xh[2]) <= 1e-6 or wh[0] == 50


# ---- QP sub-problem model: the device branch solver (csrc/qpsub.cuh) against the oracle, branch by branch ------------
def _qp_lockstep(harness, grid, data, par, rho_pq, rho_va, n_iter):
    """The oracle drives the one-level ADMM; every branch solve is replayed through the device code on identical
    inputs and must give the same u, sqp_line, multipliers, AL state and AL iteration count."""
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent))
    m = orc.OracleQpModel(grid, par, data, rho_pq, rho_va)
    nl, ng = grid.nline, grid.ngen
    worst, al_total, al_max = 0.0, 0, 0
    res = np.zeros((nl, 4)) if data.line_res is None else np.asarray(data.line_res).reshape(4, nl).T.copy()
    for it in range(n_iter):
        m.inner = 1
        v = m.vec("v_curr").copy(); z = m.vec("z_curr").copy(); l = m.vec("l_curr").copy(); rho = m.vec("rho").copy()
        sq0 = m.line_array("sqp_line").copy(); mb0 = m.line_array("qpsub_membuf").copy()
        c0 = m.counters()
        m.admm_update_x()
        c1 = m.counters()
        u2 = m.vec("u_curr"); sq2 = m.line_array("sqp_line"); mb2 = m.line_array("qpsub_membuf"); lam2 = m.line_array("lambda")
        al = 0
        for I in range(nl):
            p = 2 * ng + 8 * I
            H = np.ascontiguousarray(data.Hs[6 * I:6 * I + 6])
            Y = np.array([grid.YffR[I], grid.YffI[I], grid.YftR[I], grid.YftI[I], grid.YttR[I], grid.YttI[I],
                          grid.YtfR[I], grid.YtfI[I]])
            lin = np.concatenate([data.LH_1h[I], [data.RH_1h[I]], data.LH_1i[I], [data.RH_1i[I]], data.LH_1j[I],
                                  [data.RH_1j[I]], data.LH_1k[I], [data.RH_1k[I]]])
            sq = np.ascontiguousarray(sq0[:, I]); mb = np.ascontiguousarray(mb0[:, I])
            u = np.zeros(8); lam = np.zeros(4); work = (C.c_int * 3)()
            harness.hh_solve_qp_branch(P(H), P(np.ascontiguousarray(l[p:p + 8])), P(np.ascontiguousarray(rho[p:p + 8])),
                                       P(np.ascontiguousarray(v[p:p + 8])), P(np.ascontiguousarray(z[p:p + 8])), P(Y),
                                       P(np.ascontiguousarray(res[I])), P(lin), P(np.ascontiguousarray(data.ls[I])),
                                       P(np.ascontiguousarray(data.us[I])), P(sq), P(mb), 1, par.max_auglag, par.mu_max,
                                       par.scale, P(u), P(lam), work)
            al += work[0]
            al_max = max(al_max, work[0])
            worst = max(worst, np.abs(u - u2[p:p + 8]).max(), np.abs(sq - sq2[:, I]).max())
            assert mb[4] == mb2[4, I], (it, I, mb, mb2[:, I])
            np.testing.assert_allclose(mb[2:4], mb2[2:4, I], rtol=1e-7, atol=1e-7)
            np.testing.assert_allclose(lam, lam2[:, I], rtol=1e-6, atol=1e-6 * max(1.0, np.abs(lam2[:, I]).max()))
        assert al == c1["auglag_iters"] - c0["auglag_iters"]
        al_total += al
        m.admm_update_xbar(); m.admm_update_l_single()
    return worst, al_total, al_max


def _qp_harness(h):
    h.hh_solve_qp_branch.argtypes = [pd] * 12 + [C.c_longlong, C.c_int, C.c_double, C.c_double, pd, pd, C.POINTER(C.c_int)]
    return h


def test_qp_branch_solver_on_host_case9(host_harness, case9_grid):
    import qpsub_setup
    g = qpsub_setup.load_golden()
    data = qpsub_setup.linearise(case9_grid, g["sqp_point"])
    par = Parameters(); par.verbose = 0; par.scale = 1e-4
    worst, _, _ = _qp_lockstep(_qp_harness(host_harness), case9_grid, data, par, 4000.0, 4000.0, 30)
    assert worst <= 1e-10


def test_qp_branch_solver_on_host_binding_limits(host_harness, host_harness_nofma):
    """Synthetic grid with tight line limits: the AL loop runs several TRON solves per branch (multiplier and penalty
    updates), the slack variables leave their bounds, and the device code follows the oracle through all of it."""
    import qpsub_setup
    grid, data = qpsub_setup.synthetic_qpsub(40, 8, 56, seed=11, tight_factor=0.05, spread=0.06)
    par = Parameters(); par.verbose = 0; par.scale = 1e-4
    for h in (host_harness, host_harness_nofma):
        worst, al_total, al_max = _qp_lockstep(_qp_harness(h), grid, data, par, 400.0, 400.0, 12)
        assert al_max >= 3 and al_total > 12 * grid.nline        # the limits do bind
        assert worst <= 1e-7      # penalties up to 1e7 on a problem scaled by 1e-4: solutions are defined to ~1e-8
