"""Extract the golden vectors of the reference's multi-period ACOPF test into a JSON fixture.

Run in the build container (needs /root/reference, which does not exist on the GPU box):
    python tests/golden/make_mpacopf_golden.py
Source: /root/reference/test/algorithms/mpacopf_update_cpu.jl:28-395 (u, v, z, l of the three
periods after one inner iteration on case9, atol 1e-6) and :431-435 (end-to-end pins).

The test's load profile (ExaData artifact `mp_demand/case9_onehour_60.{Pd,Qd}`) is NOT part of
/root/reference. It is recovered here from the goldens: at the first iteration u_p = 0, lambda = z
= lz = 0 for every branch end, so l_new = -beta*rho/(beta+rho) * xbar_p, and the xbar_p of the ends at
a load bus without generators or shunts sum to -Pd/baseMVA (mpacopf_bus_kernel_cpu.jl:60-104). The
L_BR goldens carry 9 significant digits, which pins the period scale factor r_t = Pd_t / Pd_base to
~1e-8; the three load buses give the SAME r_t to that precision, and the (coarser) q-side xbar goldens
agree with it, so the profile is one scale factor per period applied to (Pd, Qd) of case9:
r = (1, 0.99961029, 0.99930348).
"""
import json
import re
from pathlib import Path

import numpy as np

SRC = Path("/root/reference/test/algorithms/mpacopf_update_cpu.jl")
OUT = Path(__file__).resolve().parent / "mpacopf_case9_golden.json"

# case9.m: load buses (1-based) with Pd, Qd in MW / MVAr, and the lines (1-based, file order) at each
LOADS = {5: (90.0, 30.0), 7: (100.0, 35.0), 9: (125.0, 50.0)}
BRANCHES = [(1, 4), (4, 5), (5, 6), (3, 6), (6, 7), (7, 8), (8, 2), (8, 9), (9, 4)]


def nested(txt, name):
    i = txt.index(name + " = [") + len(name) + 3
    depth, j = 0, i
    while True:
        c = txt[j]
        depth += (c == "[") - (c == "]")
        if depth == 0:
            break
        j += 1
    body = txt[i + 1:j]
    periods = re.findall(r"\[(.*?)\]", body, re.S)
    return [[float(t) for t in re.findall(r"-?\d+\.\d+(?:[eE][-+]?\d+)?", p)] for p in periods]


def main():
    txt = SRC.read_text()
    out = {"source": "test/algorithms/mpacopf_update_cpu.jl", "atol": 1e-6, "len_horizon": 3,
           "params": {"rho_pq": 4e2, "rho_va": 4e4, "scale": 1e-4, "initial_beta": 1e3, "beta": 1e3,
                      "ramp_ratio": 0.02}}
    for name in ("U_GEN", "U_BR", "V_GEN", "V_BR", "Z_GEN", "Z_BR", "L_GEN", "L_BR"):
        out[name] = nested(txt, name)
        assert len(out[name]) == 3 and all(len(p) == (6 if "GEN" in name else 72) for p in out[name]), name
    # period scale factors of the load profile
    par = out["params"]
    k = (par["beta"] + par["rho_pq"]) / (par["beta"] * par["rho_pq"])       # xbar_p = -k * l_new
    scales = []
    for t in range(3):
        est = []
        for bus, (pd, _qd) in LOADS.items():
            sl = 0.0
            for l, (f, to) in enumerate(BRANCHES):
                rec = out["L_BR"][t][8 * l:8 * l + 8]
                sl += (rec[0] if f == bus else 0.0) + (rec[2] if to == bus else 0.0)
            est.append(sl * k * 100.0 / pd)
        assert max(est) - min(est) < 2e-8, est
        scales.append(round(float(np.mean(est)), 8))
        # cross-check with the q-side xbar goldens (6 decimals)
        for bus, (_pd, qd) in LOADS.items():
            sq = 0.0
            for l, (f, to) in enumerate(BRANCHES):
                rec = out["V_BR"][t][8 * l:8 * l + 8]
                sq += (rec[1] if f == bus else 0.0) + (rec[3] if to == bus else 0.0)
            assert abs(-sq * 100.0 / qd - scales[-1]) < 1e-5, (t, bus)
    out["load_scale"] = scales
    out["load_scale_note"] = "recovered from V_BR, see the module docstring; period 1 is the base case"
    assert scales[0] == 1.0
    out["solve_case9_T3"] = {"kwargs": {"end_period": 3, "warm_start": False, "outer_iterlim": 25, "rho_pq": 4e2,
                                        "rho_va": 4e4, "outer_eps": 2e-5},
                             "status": "Solved", "outer": 20, "cumul": 729, "objval": 15901.48, "objval_atol": 1e-2}
    OUT.write_text(json.dumps(out, indent=1))
    print("wrote", OUT, "load scales", scales)


if __name__ == "__main__":
    main()
