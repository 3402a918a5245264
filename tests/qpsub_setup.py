"""The SQP driver's side of the QPsub tests: linearise ACOPF at an SQP iterate into the fields of ``ModelQpsub``.

Mirrors the set-up block of the reference's tests (``test/algorithms/qpsub_update_cpu.jl:29-140``), including its
literal ``line_var[1,2]`` in ``cons_3`` (:82). Test infrastructure, not product code: in production these fields are
filled in by the SQP driver (the SQOPF package), outside ExaAdmm.
"""
from __future__ import annotations

import json
from pathlib import Path
from types import SimpleNamespace

import numpy as np

GOLDEN = Path(__file__).resolve().parent / "golden" / "qpsub_case9_golden.json"


def load_golden():
    return json.loads(GOLDEN.read_text())


def linearise(grid, pt, pi_14=None, check=True):
    """``grid``: GridData; ``pt``: dict with pg, qg, line_var (6 x nline), line_fl (4 x nline), pgb, pft, ptf, qgb, qft,
    qtf, bus_w. Returns a namespace with Hs (6 nline, 6), LH_* / RH_*, ls / us (nline, 6), the shifted generator bounds
    and costs, and qpsub_Pd / qpsub_Qd - the reference's field shapes."""
    nline, nbus = grid.nline, grid.nbus
    pg, qg = np.asarray(pt["pg"], float), np.asarray(pt["qg"], float)
    lv, lf = np.asarray(pt["line_var"], float), np.asarray(pt["line_fl"], float)
    if pi_14 is None:
        pi_14 = -np.ones((4, nline))
    d = SimpleNamespace()
    d.pgmax, d.pgmin = grid.pgmax - pg, grid.pgmin - pg
    d.qgmax, d.qgmin = grid.qgmax - qg, grid.qgmin - qg
    d.c1 = grid.c1 + 2 * grid.c2 * pg
    d.c2 = np.array(grid.c2, float)
    d.ls, d.us = np.zeros((nline, 6)), np.zeros((nline, 6))
    for l in range(nline):
        vv = 2 * grid.FrVmBound[2 * l + 1] * grid.ToVmBound[2 * l + 1]
        d.ls[l, 0:2], d.us[l, 0:2] = -vv, vv
        d.ls[l, 2] = grid.FrVmBound[2 * l] ** 2 - lv[2, l]
        d.us[l, 2] = grid.FrVmBound[2 * l + 1] ** 2 - lv[2, l]
        d.ls[l, 3] = grid.ToVmBound[2 * l] ** 2 - lv[3, l]
        d.us[l, 3] = grid.ToVmBound[2 * l + 1] ** 2 - lv[3, l]
        d.ls[l, 4] = grid.FrVaBound[2 * l] - lv[4, l]
        d.us[l, 4] = grid.FrVaBound[2 * l + 1] - lv[4, l]
        d.ls[l, 5] = grid.ToVaBound[2 * l] - lv[5, l]
        d.us[l, 5] = grid.ToVaBound[2 * l + 1] - lv[5, l]
    B = grid.baseMVA
    pgb, pft, ptf = (np.asarray(pt[k], float) for k in ("pgb", "pft", "ptf"))
    qgb, qft, qtf = (np.asarray(pt[k], float) for k in ("qgb", "qft", "qtf"))
    bus_w = np.asarray(pt["bus_w"], float)
    d.Pd = B * (grid.Pd / B - (pgb - pft - ptf - grid.YshR * bus_w))
    d.Qd = B * (grid.Qd / B - (qgb - qft - qtf + grid.YshI * bus_w))
    d.Hs = np.zeros((6 * nline, 6))
    d.LH_1h, d.RH_1h = np.zeros((nline, 4)), np.zeros(nline)
    d.LH_1i, d.RH_1i = np.zeros((nline, 4)), np.zeros(nline)
    d.LH_1j, d.RH_1j = np.zeros((nline, 2)), np.zeros(nline)
    d.LH_1k, d.RH_1k = np.zeros((nline, 2)), np.zeros(nline)
    d.line_res = None
    for l in range(nline):
        H14h = np.zeros((6, 6))
        H14h[0, 0] = H14h[1, 1] = 2 * pi_14[0, l]
        H14h[2, 3] = H14h[3, 2] = -pi_14[0, l]
        dth = lv[4, l] - lv[5, l]
        c1 = pi_14[1, l] * np.cos(dth)
        c2 = pi_14[1, l] * np.sin(dth)
        c3 = pi_14[1, l] * (-lv[0, l] * np.sin(dth) + lv[0, 1] * np.cos(dth))     # line_var[1,2]: as in the reference
        H14i = np.zeros((6, 6))
        H14i[0, 4] = H14i[4, 0] = c1
        H14i[0, 5] = H14i[5, 0] = -c1
        H14i[1, 4] = H14i[4, 1] = c2
        H14i[1, 5] = H14i[5, 1] = -c2
        H14i[4, 4] = H14i[5, 5] = c3
        H14i[4, 5] = H14i[5, 4] = -c3
        supY = np.array([[grid.YftR[l], grid.YftI[l], grid.YffR[l], 0, 0, 0],
                         [-grid.YftI[l], grid.YftR[l], -grid.YffI[l], 0, 0, 0],
                         [grid.YtfR[l], -grid.YtfI[l], 0, grid.YttR[l], 0, 0],
                         [-grid.YtfI[l], -grid.YtfR[l], 0, -grid.YttI[l], 0, 0]])
        H14j = -2 * pi_14[2, l] * (np.outer(supY[0], supY[0]) + np.outer(supY[1], supY[1]))
        H14k = -2 * pi_14[3, l] * (np.outer(supY[2], supY[2]) + np.outer(supY[3], supY[3]))
        Hs = H14h + H14i + H14j + H14k + 4 * np.eye(6)
        if check:
            assert np.max(np.abs(Hs - Hs.T)) <= 1e-6
            assert np.linalg.eigvalsh(Hs).min() >= 0.0
        d.Hs[6 * l:6 * l + 6, :] = Hs
        d.LH_1h[l] = [2 * lv[0, l], 2 * lv[1, l], -lv[3, l], -lv[2, l]]
        d.RH_1h[l] = -lv[0, l] ** 2 - lv[1, l] ** 2 + lv[2, l] * lv[3, l]
        t = lv[0, l] * np.cos(dth) + lv[1, l] * np.sin(dth)
        d.LH_1i[l] = [np.sin(dth), -np.cos(dth), t, -t]
        d.RH_1i[l] = -lv[0, l] * np.sin(dth) + lv[1, l] * np.cos(dth)
        d.LH_1j[l] = [2 * lf[0, l], 2 * lf[1, l]]
        d.RH_1j[l] = -(lf[0, l] ** 2 + lf[1, l] ** 2 - grid.rateA[l])
        d.LH_1k[l] = [2 * lf[2, l], 2 * lf[3, l]]
        d.RH_1k[l] = -(lf[2, l] ** 2 + lf[3, l] ** 2 - grid.rateA[l])
    return d


def sqp_point(grid, V, theta, pg=None, qg=None):
    """An SQP iterate for any grid: bus voltages (V, theta) -> the dict ``linearise`` takes (flows from the admittances,
    generators at the middle of their bounds unless given). Used for synthetic grids, where no reference fixture exists."""
    nline, nbus = grid.nline, grid.nbus
    f = np.asarray(grid.brBusIdx[0::2]) - 1
    t = np.asarray(grid.brBusIdx[1::2]) - 1
    wi, wj = V[f] ** 2, V[t] ** 2
    wR = V[f] * V[t] * np.cos(theta[f] - theta[t])
    wI = V[f] * V[t] * np.sin(theta[f] - theta[t])
    pij = grid.YffR * wi + grid.YftR * wR + grid.YftI * wI
    qij = -grid.YffI * wi - grid.YftI * wR + grid.YftR * wI
    pji = grid.YttR * wj + grid.YtfR * wR - grid.YtfI * wI
    qji = -grid.YttI * wj - grid.YtfI * wR - grid.YtfR * wI
    pg = 0.5 * (grid.pgmin + grid.pgmax) if pg is None else np.asarray(pg, float)
    qg = 0.5 * (grid.qgmin + grid.qgmax) if qg is None else np.asarray(qg, float)
    pgb, qgb = np.zeros(nbus), np.zeros(nbus)
    for b in range(nbus):
        for k in range(grid.GenStart[b] - 1, grid.GenStart[b + 1] - 1):
            pgb[b] += pg[grid.GenIdx[k] - 1]
            qgb[b] += qg[grid.GenIdx[k] - 1]
    pft, ptf, qft, qtf = np.zeros(nbus), np.zeros(nbus), np.zeros(nbus), np.zeros(nbus)
    np.add.at(pft, f, pij); np.add.at(qft, f, qij)
    np.add.at(ptf, t, pji); np.add.at(qtf, t, qji)
    return {"pg": pg, "qg": qg, "line_var": np.vstack([wR, wI, wi, wj, theta[f], theta[t]]),
            "line_fl": np.vstack([pij, qij, pji, qji]), "pgb": pgb, "pft": pft, "ptf": ptf, "qgb": qgb, "qft": qft,
            "qtf": qtf, "bus_w": V ** 2}


def synthetic_qpsub(nbus, ngen, nline, seed, tight_factor=1.0, spread=0.03):
    """(grid, data) for a synthetic grid linearised at a random near-flat voltage profile; a small ``tight_factor``
    makes line limits bind in the QP."""
    import exaadmm_b200 as ea
    from exaadmm_b200.synthetic import synthetic_case
    rng = np.random.default_rng(seed)
    grid = ea.GridData.from_opfdata(synthetic_case(nbus, ngen, nline, seed), tight_factor=tight_factor)
    V = rng.uniform(0.97, 1.04, grid.nbus)
    theta = rng.normal(0.0, spread, grid.nbus)
    fr = np.asarray(grid.FrVaBound).reshape(-1, 2)
    ref_from = np.nonzero(fr[:, 1] == 0.0)[0]
    if len(ref_from):
        theta[np.asarray(grid.brBusIdx[0::2])[ref_from[0]] - 1] = 0.0
    pt = sqp_point(grid, V, theta)
    return grid, linearise(grid, pt, check=False)
