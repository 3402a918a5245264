"""Time the one-level ADMM on the SQP sub-problem (`ModelQpsub`) on the GPU against the CPU oracle.

    python tools/time_qpsub.py [workload|case9] [iterations] [--oracle] [--no-graph]

case9: the reference's own test problem (Solved after 5107 iterations). Otherwise a name from
exaadmm_b200.synthetic.NAMED (default case2869pegase), linearised at a random near-flat voltage profile
(tests/qpsub_setup.py), run for a fixed number of iterations. Prints one JSON line.
"""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import exaadmm_b200 as ea                                   # noqa: E402
import qpsub_setup                                          # noqa: E402
from exaadmm_b200 import qpsub                              # noqa: E402
from exaadmm_b200.environment import AdmmEnv, Parameters     # noqa: E402
from exaadmm_b200.synthetic import named_case                # noqa: E402

FIELDS = ("Hs", "LH_1h", "RH_1h", "LH_1i", "RH_1i", "LH_1j", "RH_1j", "LH_1k", "RH_1k", "ls", "us")
GEN = (("qpsub_pgmax", "pgmax"), ("qpsub_pgmin", "pgmin"), ("qpsub_qgmax", "qgmax"), ("qpsub_qgmin", "qgmin"),
       ("qpsub_c1", "c1"), ("qpsub_c2", "c2"), ("qpsub_Pd", "Pd"), ("qpsub_Qd", "Qd"))


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    workload = args[0] if args else "case2869pegase"
    with_oracle = "--oracle" in sys.argv
    use_graph = "--no-graph" not in sys.argv
    if workload == "case9":
        case = ea.CASE9
        grid = ea.GridData.from_opfdata(ea.parse_matpower(case))
        data = qpsub_setup.linearise(grid, qpsub_setup.load_golden()["sqp_point"])
        iters, eps = 10000, 2e-6
    else:
        case = named_case(workload)
        grid = ea.GridData.from_opfdata(case, tight_factor=0.6)
        rng = np.random.default_rng(7)
        pt = qpsub_setup.sqp_point(grid, rng.uniform(0.98, 1.03, grid.nbus), rng.normal(0.0, 0.02, grid.nbus))
        data = qpsub_setup.linearise(grid, pt, check=False)
        iters, eps = (int(args[1]) if len(args) > 1 else 2000), 1e-12
    rho = 4000.0
    out = {"workload": workload, "nline": grid.nline, "ngen": grid.ngen, "nbus": grid.nbus, "graph": use_graph}

    for rep in range(2):                                     # first pass warms the context / memory pool
        env = AdmmEnv(case, rho, rho, use_gpu=True, verbose=0)
        mod = ea.ModelQpsub(env, grid=grid)
        for k in FIELDS:
            setattr(mod, k, np.array(getattr(data, k)))
        for k, src in GEN:
            setattr(mod, k, np.array(getattr(data, src)))
        env.params.scale, env.params.outer_eps, env.params.outer_iterlim = 1e-4, eps, iters
        qpsub.init_solution(mod, mod.solution, rho, rho)
        mod.set_option("use_graph", 1 if use_graph else 0)
        t0 = time.perf_counter()
        ea.admm_one_level(env, mod, None, mode="native")
        wall = time.perf_counter() - t0
        kt, c = mod.kernel_times(), mod.counters()
        out.update(status=mod.info.status, iterations=int(mod.info.outer), objval=mod.info.objval,
                   mismatch=mod.info.mismatch, gpu_wall_s=wall, gpu_device_s=kt["device_s"],
                   gpu_it_per_s=mod.info.outer / kt["device_s"], us_per_iteration=1e6 * kt["device_s"] / max(mod.info.outer, 1),
                   launches=kt["launches"], auglag_per_branch=c["auglag_iters"] / max(c["line_calls"], 1),
                   evals_per_branch=c["tron_evals"] / max(c["line_calls"], 1), max_auglag=c["max_auglag_one_call"])
        mod.close()

    if with_oracle:
        from oracle.oracle import OracleQpModel
        for threads in (1, os.cpu_count() or 1):
            par = Parameters(); par.verbose = 0
            par.scale, par.outer_eps, par.outer_iterlim = 1e-4, eps, (iters if workload == "case9" else min(iters, 200))
            om = OracleQpModel(grid, par, data, rho, rho)
            om.set_threads(threads)
            info = om.admm_one_level()
            out[f"oracle_{threads}t_it_per_s"] = info.outer / info.time_overall
            out[f"oracle_{threads}t_objval"] = info.objval
            out[f"oracle_{threads}t_iterations"] = int(info.outer)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
