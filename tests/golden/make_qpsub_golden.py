"""Extract the inputs and golden vectors of the reference's own QPsub tests into a JSON fixture.

Run in the build container (needs /root/reference, which does not exist on the
GPU box):  python tests/golden/make_qpsub_golden.py
Sources:
  /root/reference/test/algorithms/qpsub_update_cpu.jl:15-27   the SQP iterate the QP is linearised at
                                                 :160-197  u, v, l, rp, rd after one ADMM iteration (atol 2e-6)
                                                 :224-237  Solved / 5107 / 5107 / objval -21.92744641968529
  /root/reference/test/algorithms/qpsub_update_gpu.jl:228-346 step, KKT error and multipliers handed back to the
                                                 SQP driver after the full solve (atol 1e-6, lambda 2e-6 rel.)
"""
import json
import re
from pathlib import Path

CPU = Path("/root/reference/test/algorithms/qpsub_update_cpu.jl")
GPU = Path("/root/reference/test/algorithms/qpsub_update_gpu.jl")
OUT = Path(__file__).resolve().parent / "qpsub_case9_golden.json"

NUM = r"-?\d+\.?\d*(?:[eE][-+]?\d+)?"


def grab(txt, name):
    m = re.search(rf"(?<![\w.]){name}\s*=\s*\[(.*?)\]", txt, re.S)
    assert m, name
    return m.group(1)


def vec(txt, name):
    return [float(t) for t in re.findall(NUM, grab(txt, name))]


def mat(txt, name, rows):
    body = grab(txt, name)
    parts = [p for p in (body.split(";") if ";" in body else body.strip().split("\n")) if p.strip()]
    out = [[float(t) for t in re.findall(NUM, p)] for p in parts]
    assert len(out) == rows and len({len(r) for r in out}) == 1, (name, len(out))
    return out


def main():
    c = CPU.read_text()
    g = GPU.read_text()
    out = {"source": "test/algorithms/qpsub_update_{cpu,gpu}.jl",
           "params_one_iteration": {"rho_pq": 20.0, "rho_va": 20.0, "scale": 1e-4, "atol": 2e-6},
           "sqp_point": {k: vec(c, k) for k in ("pg", "qg", "pgb", "pft", "ptf", "qgb", "qft", "qtf", "bus_w")}}
    out["sqp_point"]["line_var"] = mat(c, "line_var", 6)
    out["sqp_point"]["line_fl"] = mat(c, "line_fl", 4)
    for k in ("U_SOL", "V_SOL", "L_SOL", "RP_SOL", "RD_SOL"):
        out[k] = vec(c, k)
        assert len(out[k]) == 78, (k, len(out[k]))
    out["solve"] = {"kwargs": {"initial_beta": 100000.0, "outer_iterlim": 10000, "inner_iterlim": 1, "scale": 1e-4,
                               "obj_scale": 1, "rho_pq": 4000.0, "rho_va": 4000.0, "outer_eps": 2e-6},
                    "status": "Solved", "outer": 5107, "cumul": 5107, "objval": -21.92744641968529,
                    "objval_atol": 1e-6, "atol": 1e-6, "lambda_rtol": 2e-6,
                    "dpg_sol": vec(g, "dpg_sol_cpu"), "dqg_sol": vec(g, "dqg_sol_cpu"),
                    "dline_var": mat(g, "dline_var_cpu", 6), "dline_fl": mat(g, "dline_fl_cpu", 4),
                    "dtheta_sol": vec(g, "dtheta_sol_cpu"), "dw_sol": vec(g, "dw_sol_cpu"),
                    "lambda": mat(g, "lambda_cpu", 4)}
    di = []
    for i in range(1, 6):
        di += vec(g, f"dual_infeas_{i}_cpu")
    assert len(di) == 3 + 6 * 9, len(di)
    out["solve"]["dual_infeas"] = di
    assert len(out["solve"]["dpg_sol"]) == 3 and len(out["solve"]["dw_sol"]) == 9
    OUT.write_text(json.dumps(out, indent=1))
    print("wrote", OUT)


if __name__ == "__main__":
    main()
