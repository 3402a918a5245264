"""Unit-level checks of the oracle's branch objective and TRON restatement."""
import numpy as np
import pytest

from oracle import oracle as orc


def _random_problem(rng, binding=False):
    r, x, b = rng.uniform(0.001, 0.04), rng.uniform(0.01, 0.2), rng.uniform(0, 0.4)
    ys = 1 / (r + 1j * x)
    tap = rng.uniform(0.9, 1.1) * np.exp(1j * rng.uniform(-0.2, 0.2))
    ytt = ys + 0.5j * b
    yff = ytt / (tap * np.conj(tap)); yft = -ys / np.conj(tap); ytf = -ys / tap
    Y = np.array([yff.real, yff.imag, yft.real, yft.imag, ytt.real, ytt.imag, ytf.real, ytf.imag])
    param = np.zeros(31)
    param[0:8] = rng.normal(0, 5, 8)
    param[8:12] = rng.choice([10.0, 400.0]); param[12:16] = param[8] * 100
    param[16:20] = rng.normal(0, 0.5, 4); param[20:22] = rng.uniform(0.9, 1.1, 2); param[22:24] = rng.normal(0, 0.2, 2)
    param[24:26] = rng.uniform(0, 2, 2) if binding else 0.0
    param[26] = 10.0 ** rng.integers(1, 5)
    rate = (0.05 if binding else 4.0)
    param[28] = rate
    x0 = np.array([rng.uniform(0.95, 1.05), rng.uniform(0.95, 1.05), rng.normal(0, 0.1), rng.normal(0, 0.1),
                   -rng.uniform(0, rate), -rng.uniform(0, rate)])
    xl = np.array([0.9, 0.9, -2 * np.pi, -2 * np.pi, -rate, -rate])
    xu = np.array([1.1, 1.1, 2 * np.pi, 2 * np.pi, 0.0, 0.0])
    return x0, xl, xu, param, Y


def test_compact_objective_equals_reference_term_order():
    rng = np.random.default_rng(0)
    for _ in range(200):
        x, _, _, p, Y = _random_problem(rng, binding=True)
        f1 = orc.eval_f(x, p, Y, 1e-4)
        f2 = orc.eval_f(x, p, Y, 1e-4, reforder=True)
        assert abs(f1 - f2) <= 1e-12 * max(1.0, abs(f2))


def test_gradient_and_hessian_against_finite_differences():
    rng = np.random.default_rng(1)
    for _ in range(40):
        x, _, _, p, Y = _random_problem(rng, binding=True)
        g, H = orc.eval_gh(x, p, Y, 1.0)
        np.testing.assert_allclose(H, H.T, rtol=0, atol=0)
        h = 1e-6
        gfd = np.zeros(6); Hfd = np.zeros((6, 6))
        for i in range(6):
            e = np.zeros(6); e[i] = h
            gfd[i] = (orc.eval_f(x + e, p, Y, 1.0, reforder=True) - orc.eval_f(x - e, p, Y, 1.0, reforder=True)) / (2 * h)
            gp, _ = orc.eval_gh(x + e, p, Y, 1.0); gm, _ = orc.eval_gh(x - e, p, Y, 1.0)
            Hfd[:, i] = (gp - gm) / (2 * h)
        sc = max(1.0, np.abs(g).max())
        np.testing.assert_allclose(g, gfd, atol=2e-6 * sc, rtol=1e-6)
        np.testing.assert_allclose(H, Hfd, atol=2e-5 * max(1.0, np.abs(H).max()), rtol=1e-5)
        # spot checks quoted in SURVEY App. A.4 (eval_cpu.jl:593-600)
        assert H[4, 4] == p[26] and H[5, 5] == p[26] and H[4, 5] == 0.0


@pytest.mark.parametrize("binding", [False, True])
def test_tron_reaches_a_projected_stationary_point(binding):
    rng = np.random.default_rng(2 + binding)
    for _ in range(60):
        x0, xl, xu, p, Y = _random_problem(rng, binding)
        x, status, minor, nfev = orc.tron_solve(x0, xl, xu, p, Y, 1e-4)
        assert status == 0 and minor < 200
        assert np.all(x >= xl) and np.all(x <= xu)
        g, _ = orc.eval_gh(x, p, Y, 1e-4)
        pg = np.where(x == xl, np.minimum(g, 0), np.where(x == xu, np.maximum(g, 0), g))
        # converged either on the projected gradient (gtol) or on the frtol test
        assert np.abs(pg).max() <= 1e-6 or nfev > 1
        assert orc.eval_f(x, p, Y, 1e-4) <= orc.eval_f(x0, p, Y, 1e-4) + 1e-15


def test_tron_fixed_reference_angle_stays_put():
    rng = np.random.default_rng(5)
    x0, xl, xu, p, Y = _random_problem(rng)
    xl[2] = xu[2] = 0.0; x0[2] = 0.0          # reference-bus end: (0, 0) bound (opfdata.jl:702-713)
    x, status, _, _ = orc.tron_solve(x0, xl, xu, p, Y, 1e-4)
    assert status == 0 and x[2] == 0.0
