"""Deterministic synthetic transmission grids with prescribed bus / generator /
branch counts.

The reference's benchmark cases (pegase, ACTIVSg) come from a lazy artifact that
needs network access (``/root/reference/test/Artifacts.toml``); BASELINE.json
allows "synthetic grids with the named bus/branch/generator counts" instead.
Recipe: connected graph = random spanning tree with index
locality + chords (a few % parallel circuits and tapped transformers), a feasible
operating point by construction (voltages drawn, angles from a DC solve, exact AC
flows at that point, loads absorb the remainder), quadratic costs, ``rateA`` =
1.3 x the flow at the construction point (some unlimited). The case is returned
as MATPOWER matrices so it goes through the same parser as a real file.
"""
from __future__ import annotations

from pathlib import Path

import numpy as np

from .matpower import OPFData, parse_matpower_text, write_matpower

NAMED = {
    # name: (nbus, ngen, nline) — MATPOWER nominal totals (recalled)
    "case9": (9, 3, 9),
    "case1354pegase": (1354, 260, 1991),
    "case2869pegase": (2869, 510, 4582),
    "case9241pegase": (9241, 1445, 16049),
    "case13659pegase": (13659, 4092, 20467),
    "ACTIVSg25k": (25000, 4834, 32230),
    "ACTIVSg70k": (70000, 10390, 88207),
}


def synthetic_matrices(nbus: int, ngen: int, nline: int, seed: int, *, base_mva: float = 100.0,
                       mean_load_pu: float = 0.05, window: int = 24, rate_margin: float = 1.3,
                       frac_unlimited: float = 0.1, frac_transformer: float = 0.05,
                       frac_parallel: float = 0.03):
    if nline < nbus - 1:
        raise ValueError("need nline >= nbus - 1 for a connected grid")
    rng = np.random.default_rng(seed)

    # ---- topology ------------------------------------------------------------------
    child = np.arange(1, nbus)
    lo = np.maximum(0, child - window)
    parent = lo + (rng.random(nbus - 1) * (child - lo)).astype(np.int64)
    fr = [parent]
    to = [child]
    extra = nline - (nbus - 1)
    n_par = min(int(round(frac_parallel * nline)), extra)
    n_chord = extra - n_par
    if n_chord > 0:
        a = rng.integers(0, nbus, size=n_chord)
        off = rng.integers(1, max(2, min(window * 4, nbus)), size=n_chord)
        b = (a + off) % nbus
        fr.append(a); to.append(b)
    if n_par > 0:
        pick = rng.integers(0, nbus - 1, size=n_par)
        fr.append(parent[pick]); to.append(child[pick])
    fr = np.concatenate(fr); to = np.concatenate(to)
    flip = rng.random(nline) < 0.5
    fr, to = np.where(flip, to, fr), np.where(flip, fr, to)
    perm = np.argsort(np.minimum(fr, to) + rng.random(nline), kind="stable")   # roughly bus-ordered, like real files
    fr, to = fr[perm], to[perm]

    r = rng.uniform(0.001, 0.04, nline)
    x = rng.uniform(0.01, 0.2, nline)
    bch = rng.uniform(0.0, 0.4, nline)
    is_tr = rng.random(nline) < frac_transformer
    ratio = np.where(is_tr, rng.uniform(0.9, 1.1, nline), 0.0)
    bch = np.where(is_tr, 0.0, bch)
    r = np.where(is_tr, r * 0.2, r)

    # ---- injections and operating point -------------------------------------------------
    has_load = rng.random(nbus) < 0.7
    pd0 = np.where(has_load, rng.uniform(0.0, 2.0 * mean_load_pu / 0.7, nbus), 0.0)
    gen_bus = np.sort(rng.choice(nbus, size=ngen, replace=ngen > nbus))
    share = rng.uniform(0.5, 1.5, ngen)
    pg_op = share / share.sum() * pd0.sum() * 1.02
    p_inj = -pd0.copy()
    np.add.at(p_inj, gen_bus, pg_op)
    p_inj -= p_inj.mean()

    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    bsus = 1.0 / x
    rows = np.concatenate([fr, to, fr, to]); cols = np.concatenate([fr, to, to, fr])
    vals = np.concatenate([bsus, bsus, -bsus, -bsus])
    B = sp.csr_matrix((vals, (rows, cols)), shape=(nbus, nbus))
    keep = np.arange(1, nbus)
    theta = np.zeros(nbus)
    theta[1:] = spla.spsolve(B[keep][:, keep].tocsc(), p_inj[1:])
    dth = np.abs(theta[fr] - theta[to]).max()
    if dth > 0.35:                       # keep angle differences moderate
        theta *= 0.35 / dth
    vm = rng.uniform(0.98, 1.04, nbus)

    # exact AC flows at (vm, theta)
    ys = 1.0 / (r + 1j * x)
    tap = np.where(ratio == 0.0, 1.0, ratio).astype(complex)
    ytt = ys + 1j * bch / 2
    yff = ytt / (tap * np.conj(tap))
    yft = -ys / np.conj(tap)
    ytf = -ys / tap
    V = vm * np.exp(1j * theta)
    s_f = V[fr] * np.conj(yff * V[fr] + yft * V[to])
    s_t = V[to] * np.conj(ytf * V[fr] + ytt * V[to])
    s_bus = np.zeros(nbus, dtype=complex)
    np.add.at(s_bus, fr, s_f)
    np.add.at(s_bus, to, s_t)

    # generators: dispatch so that loads stay close to pd0; loads absorb the remainder
    pg_bus = np.zeros(nbus); np.add.at(pg_bus, gen_bus, pg_op)
    pd = pg_bus - s_bus.real
    neg = pd < 0                           # push negative loads into the local generator if there is one
    qd0 = 0.3 * np.maximum(pd, 0.0)
    is_gen_bus = np.zeros(nbus, dtype=bool); is_gen_bus[gen_bus] = True
    qg_bus = np.where(is_gen_bus, qd0 + s_bus.imag, 0.0)
    qd = np.where(is_gen_bus, qd0, -s_bus.imag)
    cnt = np.bincount(gen_bus, minlength=nbus).astype(float)
    qg_op = qg_bus[gen_bus] / cnt[gen_bus]
    pmax = pg_op / rng.uniform(0.3, 0.8, ngen)
    pmin = 0.1 * pmax
    qmax = np.maximum(0.6 * pmax, qg_op + 0.3 * pmax)
    qmin = np.minimum(-0.6 * pmax, qg_op - 0.3 * pmax)
    c2 = rng.uniform(0.01, 0.1, ngen)
    c1 = rng.uniform(10.0, 40.0, ngen)
    c0 = rng.uniform(0.0, 100.0, ngen)

    smax = np.maximum(np.abs(s_f), np.abs(s_t)) * base_mva
    rate = np.maximum(rate_margin * smax, 5.0)
    rate = np.where(rng.random(nline) < frac_unlimited, 0.0, rate)

    btype = np.ones(nbus); btype[gen_bus] = 2; btype[0] = 3
    bus = np.column_stack([np.arange(1, nbus + 1), btype, pd * base_mva, qd * base_mva, np.zeros(nbus),
                           np.zeros(nbus), np.ones(nbus), vm, theta * 180 / np.pi, np.full(nbus, 345.0),
                           np.ones(nbus), np.full(nbus, 1.1), np.full(nbus, 0.9)])
    gen = np.zeros((ngen, 21))
    gen[:, 0] = gen_bus + 1; gen[:, 1] = pg_op * base_mva; gen[:, 2] = qg_op * base_mva
    gen[:, 3] = qmax * base_mva; gen[:, 4] = qmin * base_mva; gen[:, 5] = vm[gen_bus]; gen[:, 6] = base_mva
    gen[:, 7] = 1; gen[:, 8] = pmax * base_mva; gen[:, 9] = pmin * base_mva
    branch = np.column_stack([fr + 1, to + 1, r, x, bch, rate, rate, rate, ratio, np.zeros(nline),
                              np.ones(nline), np.full(nline, -360.0), np.full(nline, 360.0)])
    gencost = np.column_stack([np.full(ngen, 2.0), np.zeros(ngen), np.zeros(ngen), np.full(ngen, 3.0), c2, c1, c0])
    feasible_cost = float(np.sum(c2 * (pg_op * base_mva) ** 2 + c1 * (pg_op * base_mva) + c0))
    return dict(baseMVA=base_mva, bus=bus, gen=gen, branch=branch, gencost=gencost), feasible_cost, int(neg.sum())


def synthetic_case(nbus: int, ngen: int, nline: int, seed: int, path=None, name: str = "synthetic", **kw) -> OPFData:
    """Build the grid; optionally write it as a MATPOWER file; return the parsed OPFData."""
    mats, feas_cost, _ = synthetic_matrices(nbus, ngen, nline, seed, **kw)
    if path is not None:
        write_matpower(path, name=name, **mats)
        from .matpower import parse_matpower
        data = parse_matpower(path)
    else:
        data = _from_matrices(mats, name)
    data.extras["feasible_cost"] = feas_cost
    return data


def _from_matrices(mats, name) -> OPFData:
    """Same result as writing + parsing, without the text round trip (values are
    written with repr(), which round-trips binary64 exactly)."""
    import io
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        p = Path(td) / f"{name}.m"
        write_matpower(p, name=name, **mats)
        return parse_matpower_text(p.read_text(), case=name)


def named_case(name: str, seed: int | None = None, path=None, **kw) -> OPFData:
    """``case1354pegase``-like etc.: synthetic grid with the named counts; seed
    defaults to the bus count."""
    nbus, ngen, nline = NAMED[name]
    return synthetic_case(nbus, ngen, nline, seed if seed is not None else nbus, path=path,
                          name=f"{name}_like", **kw)
