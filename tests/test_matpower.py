"""Host logic: MATPOWER ingestion and GridData construction
(reference: src/utils/parse_matpower.jl, src/utils/opfdata.jl)."""
import numpy as np
import pytest

import exaadmm_b200 as ea
from exaadmm_b200.matpower import parse_matpower_text, MatpowerFormatError
from exaadmm_b200.synthetic import synthetic_case, synthetic_matrices

CASE = open(ea.CASE9).read()


def test_case9_counts_and_pu():
    d = ea.parse_matpower(ea.CASE9)
    assert (d.nbus, d.ngen, d.nline) == (9, 3, 9)
    assert d.baseMVA == 100.0 and d.bus_ref == 0
    np.testing.assert_allclose(d.Pmax, [2.5, 3.0, 2.7])          # p.u. (parse_matpower.jl:139-141)
    np.testing.assert_allclose(d.Pd[[4, 6, 8]], [90, 100, 125])  # MW, not p.u.
    np.testing.assert_allclose(d.coeff[0], [0.11, 5, 150])


def test_case9_grid_arrays_match_reference_conventions():
    g = ea.GridData.from_opfdata(ea.parse_matpower(ea.CASE9))
    # 1-based CSR, file order inside a bus (opfdata.jl:613-618)
    assert g.FrStart.tolist() == [1, 2, 2, 3, 4, 5, 6, 7, 9, 10]
    assert g.FrIdx.tolist() == [1, 4, 2, 3, 5, 6, 7, 8, 9]
    assert g.ToIdx.tolist() == [7, 1, 9, 2, 3, 4, 5, 6, 8]
    assert g.GenStart.tolist() == [1, 2, 3, 4, 4, 4, 4, 4, 4, 4]
    # reference bus angle fixed to 0 (opfdata.jl:702-713): line 1 starts at bus 1
    assert g.FrVaBound[:2].tolist() == [0.0, 0.0]
    np.testing.assert_allclose(g.ToVaBound[:2], [-2 * np.pi, 2 * np.pi])
    # rateA: squared p.u. rating (opfdata.jl:714)
    np.testing.assert_allclose(g.rateA, [6.25, 6.25, 2.25, 9.0, 2.25, 6.25, 6.25, 6.25, 6.25])
    # branch 1: r = 0, x = 0.0576, b = 0 -> Yff = -j/x
    np.testing.assert_allclose([g.YffR[0], g.YffI[0], g.YftI[0]], [0.0, -1 / 0.0576, 1 / 0.0576])
    # branch 2 has charging: Ytt = Ys + j b/2
    ys = 1 / (0.017 + 0.092j)
    np.testing.assert_allclose([g.YttR[1], g.YttI[1]], [ys.real, ys.imag + 0.079])


def test_tight_factor_and_unlimited_rating():
    txt = CASE.replace("1\t4\t0\t0.0576\t0\t250\t250\t250", "1\t4\t0\t0.0576\t0\t0\t250\t250")
    g = ea.GridData.from_opfdata(parse_matpower_text(txt), tight_factor=0.99)
    assert g.rateA[0] == 1e3                       # rateA == 0 -> 1e3
    np.testing.assert_allclose(g.rateA[1], 0.99 * 2.5 ** 2)


def test_status_filter_and_inf_limits():
    txt = CASE.replace("2\t163\t6.54\t300\t-300\t1.025\t100\t1\t300", "2\t163\t6.54\tInf\t-Inf\t1.025\t100\t1\t300")
    txt = txt.replace("3\t85\t-10.95\t300\t-300\t1.025\t100\t1", "3\t85\t-10.95\t300\t-300\t1.025\t100\t0")
    txt = txt.replace("9\t4\t0.01\t0.085\t0.176\t250\t250\t250\t0\t0\t1", "9\t4\t0.01\t0.085\t0.176\t250\t250\t250\t0\t0\t0")
    d = parse_matpower_text(txt)
    assert d.ngen == 2 and d.nline == 8 and d.n_gen_total == 3
    assert d.Qmax[1] == 999.99 and d.Qmin[1] == -999.99      # opfdata.jl:199-206
    np.testing.assert_allclose(d.coeff[:, 0], [0.11, 0.085])  # gencost rows follow the gen rows


def test_reference_bus_errors():
    with pytest.raises(MatpowerFormatError, match="More than one reference bus"):
        parse_matpower_text(CASE.replace("2\t2\t0\t0\t0\t0\t1\t1\t0\t345", "2\t3\t0\t0\t0\t0\t1\t1\t0\t345"))
    with pytest.raises(MatpowerFormatError, match="reference bus was not found"):
        parse_matpower_text(CASE.replace("1\t3\t0\t0\t0\t0\t1\t1\t0\t345", "1\t2\t0\t0\t0\t0\t1\t1\t0\t345"))


def test_transformer_tap_and_shift():
    txt = CASE.replace("4\t5\t0.017\t0.092\t0.158\t250\t250\t250\t0\t0", "4\t5\t0.017\t0.092\t0.158\t250\t250\t250\t0.95\t10")
    g = ea.GridData.from_opfdata(parse_matpower_text(txt))
    ys = 1 / (0.017 + 0.092j)
    tap = 0.95 * np.exp(1j * np.deg2rad(10))
    yft = -ys / np.conj(tap)
    np.testing.assert_allclose([g.YftR[1], g.YftI[1]], [yft.real, yft.imag], rtol=1e-14)
    np.testing.assert_allclose(g.YffR[1], ((ys + 0.079j) / 0.95 ** 2).real, rtol=1e-14)


def test_synthetic_is_deterministic_connected_and_round_trips(tmp_path):
    a, _, _ = synthetic_matrices(300, 40, 420, seed=7)
    b, _, _ = synthetic_matrices(300, 40, 420, seed=7)
    for k in ("bus", "gen", "branch", "gencost"):
        np.testing.assert_array_equal(a[k], b[k])
    d = synthetic_case(300, 40, 420, seed=7, path=tmp_path / "syn.m")
    assert (d.nbus, d.ngen, d.nline) == (300, 40, 420)
    d2 = synthetic_case(300, 40, 420, seed=7)                       # in-memory path parses identically
    np.testing.assert_array_equal(d.br_x, d2.br_x)
    np.testing.assert_array_equal(d.Pd, d2.Pd)
    # connected
    import scipy.sparse as sp
    import scipy.sparse.csgraph as cg
    A = sp.coo_matrix((np.ones(d.nline), (d.br_from, d.br_to)), shape=(d.nbus, d.nbus))
    assert cg.connected_components(A, directed=False)[0] == 1
    # the construction point is AC-feasible: bus balance closes by construction
    assert np.isfinite(d.extras["feasible_cost"])
