"""Extract the golden vectors of the reference's own ACOPF test into a JSON fixture.

Run in the build container (needs /root/reference, which does not exist on the
GPU box):  python tests/golden/make_case9_golden.py
Source: /root/reference/test/algorithms/acopf_update_cpu.jl:28-151 (u, v, z, l
after one inner iteration on case9, atol 1e-6) and :168-179 (end-to-end pins).
"""
import json
import re
from pathlib import Path

SRC = Path("/root/reference/test/algorithms/acopf_update_cpu.jl")
OUT = Path(__file__).resolve().parent / "case9_reference_golden.json"


def main():
    txt = SRC.read_text()
    out = {"source": "test/algorithms/acopf_update_cpu.jl", "atol": 1e-6,
           "params": {"rho_pq": 4e2, "rho_va": 4e4, "scale": 1e-4, "initial_beta": 1e3, "beta": 1e3}}
    for name in ("U_GEN", "U_BR", "V_GEN", "V_BR", "Z_GEN", "Z_BR", "L_GEN", "L_BR"):
        m = re.search(rf"{name}\s*=\s*\[(.*?)\]", txt, re.S)
        out[name] = [float(t) for t in re.findall(r"-?\d+\.\d+(?:[eE][-+]?\d+)?|-?\d+", m.group(1))]
    assert len(out["U_GEN"]) == 6 and len(out["U_BR"]) == 72
    out["solve_case9"] = {"kwargs": {"outer_iterlim": 25, "rho_pq": 4e2, "rho_va": 4e4, "outer_eps": 2e-5},
                          "status": "Solved", "outer": 20, "cumul": 705, "objval": 5303.435, "objval_atol": 1e-3}
    out["solve_case118"] = {"status": "Solved", "outer": 20, "cumul": 1232, "objval": 129645.676,
                            "objval_rtol": 1e-6, "note": "case118.m is not available offline"}
    OUT.write_text(json.dumps(out, indent=1))
    print("wrote", OUT)


if __name__ == "__main__":
    main()
