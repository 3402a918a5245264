"""``GridData``: the flat arrays the ADMM hot path consumes.

Restates ``/root/reference/src/utils/grid_data.jl:61-83`` and the getters it
calls (``/root/reference/src/utils/opfdata.jl:417-483`` admittances,
``:542-618`` generator and bus data, ``:686-714`` branch bounds and ``rateA``,
``:790-805`` ``brBusIdx``). Integer index arrays keep the reference's
convention — ``int64`` and **1-based** — because that is what the Julia side
holds and what the C ABI (``include/exaadmm_b200.h``) accepts; the library
converts to its own layout internally.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .matpower import OPFData


def compute_admittances(d: OPFData):
    """``computeAdmitances`` (``opfdata.jl:417-483``)."""
    ys = 1.0 / (d.br_r + 1j * d.br_x)
    ratio = np.where(d.br_ratio == 0.0, 1.0, d.br_ratio)
    tap = ratio * np.exp(1j * (d.br_angle * np.pi / 180.0))
    ytt = ys + 1j * (d.br_b / 2.0)
    yff = ytt / (tap * np.conj(tap))
    yft = -ys / np.conj(tap)
    ytf = -ys / tap
    for name, y in (("Yff", yff), ("Ytt", ytt), ("Yft", yft), ("Ytf", ytf)):
        if not np.all(np.isfinite(y)):
            raise ValueError(f"non-finite admittance in {name} (r = x = 0 branch?)")  # opfdata.jl:459-468
    return yff, yft, ytt, ytf


def _csr_1based(owner: np.ndarray, nbus: int):
    """Group item ids by owning bus, file order inside a bus
    (``opfdata.jl:613-618``): returns (Start[nbus+1], Idx[nitems]), 1-based."""
    order = np.argsort(owner, kind="stable")
    counts = np.bincount(owner, minlength=nbus)
    start = np.concatenate(([1], 1 + np.cumsum(counts))).astype(np.int64)
    return start, (order + 1).astype(np.int64)


@dataclass
class GridData:
    baseMVA: float
    ngen: int
    nline: int
    nbus: int
    pgmin: np.ndarray
    pgmax: np.ndarray
    qgmin: np.ndarray
    qgmax: np.ndarray
    c2: np.ndarray
    c1: np.ndarray
    c0: np.ndarray
    YshR: np.ndarray
    YshI: np.ndarray
    YffR: np.ndarray
    YffI: np.ndarray
    YftR: np.ndarray
    YftI: np.ndarray
    YttR: np.ndarray
    YttI: np.ndarray
    YtfR: np.ndarray
    YtfI: np.ndarray
    FrVmBound: np.ndarray
    ToVmBound: np.ndarray
    FrVaBound: np.ndarray
    ToVaBound: np.ndarray
    rateA: np.ndarray
    FrStart: np.ndarray
    FrIdx: np.ndarray
    ToStart: np.ndarray
    ToIdx: np.ndarray
    GenStart: np.ndarray
    GenIdx: np.ndarray
    Pd: np.ndarray
    Qd: np.ndarray
    Vmin: np.ndarray
    Vmax: np.ndarray
    brBusIdx: np.ndarray
    ramp_rate: np.ndarray | None = None

    @classmethod
    def from_opfdata(cls, d: OPFData, tight_factor: float = 1.0) -> "GridData":
        nbus, nline, ngen = d.nbus, d.nline, d.ngen
        yff, yft, ytt, ytf = compute_admittances(d)
        f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)

        fr_vm = np.empty(2 * nline); fr_vm[0::2] = d.Vmin[d.br_from]; fr_vm[1::2] = d.Vmax[d.br_from]
        to_vm = np.empty(2 * nline); to_vm[0::2] = d.Vmin[d.br_to]; to_vm[1::2] = d.Vmax[d.br_to]
        fr_va = np.empty(2 * nline); fr_va[0::2] = -2 * np.pi; fr_va[1::2] = 2 * np.pi
        to_va = fr_va.copy()
        at_ref = d.br_from == d.bus_ref                      # opfdata.jl:704-707
        fr_va[0::2][at_ref] = 0.0; fr_va[1::2][at_ref] = 0.0
        at_ref = d.br_to == d.bus_ref                        # opfdata.jl:708-711
        to_va[0::2][at_ref] = 0.0; to_va[1::2][at_ref] = 0.0
        # opfdata.jl:714 — squared p.u. rating, 0 means "unlimited"
        rate = np.where(d.br_rateA == 0.0, 1e3, tight_factor * (d.br_rateA / d.baseMVA) ** 2)

        fr_start, fr_idx = _csr_1based(d.br_from, nbus)
        to_start, to_idx = _csr_1based(d.br_to, nbus)
        gen_start, gen_idx = _csr_1based(d.gen_bus, nbus)
        br_bus = np.empty(2 * nline, dtype=np.int64)
        br_bus[0::2] = d.br_from + 1
        br_bus[1::2] = d.br_to + 1

        return cls(
            baseMVA=d.baseMVA, ngen=ngen, nline=nline, nbus=nbus,
            pgmin=f64(d.Pmin), pgmax=f64(d.Pmax), qgmin=f64(d.Qmin), qgmax=f64(d.Qmax),
            c2=f64(d.coeff[:, 0]), c1=f64(d.coeff[:, 1]), c0=f64(d.coeff[:, 2]),
            YshR=f64(d.Gs / d.baseMVA), YshI=f64(d.Bs / d.baseMVA),
            YffR=f64(yff.real), YffI=f64(yff.imag), YftR=f64(yft.real), YftI=f64(yft.imag),
            YttR=f64(ytt.real), YttI=f64(ytt.imag), YtfR=f64(ytf.real), YtfI=f64(ytf.imag),
            FrVmBound=f64(fr_vm), ToVmBound=f64(to_vm), FrVaBound=f64(fr_va), ToVaBound=f64(to_va),
            rateA=f64(rate),
            FrStart=fr_start, FrIdx=fr_idx, ToStart=to_start, ToIdx=to_idx,
            GenStart=gen_start, GenIdx=gen_idx,
            Pd=f64(d.Pd), Qd=f64(d.Qd), Vmin=f64(d.Vmin), Vmax=f64(d.Vmax),
            brBusIdx=br_bus,
        )
