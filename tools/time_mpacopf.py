"""Time the multi-period ACOPF solve on the GPU against the CPU oracle on the same case.

    python tools/time_mpacopf.py [workload] [T] [--oracle]

workload: a name from exaadmm_b200.synthetic.NAMED (default case1354pegase). Loads follow a smooth +-3 % daily
shape over the T periods. Prints one JSON line: iterations, seconds, inner-iterations/s (all periods advance together,
so one iteration = T period-iterations), launches per iteration, and the oracle's figures when --oracle is given.
"""
import json
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import exaadmm_b200 as ea                                   # noqa: E402
from exaadmm_b200.environment import AdmmEnv, Parameters     # noqa: E402
from exaadmm_b200.mpacopf import ModelMpacopf                # noqa: E402
from exaadmm_b200.admm_two_level import admm_two_level       # noqa: E402
from exaadmm_b200.synthetic import named_case                # noqa: E402


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    workload = args[0] if args else "case1354pegase"
    T = int(args[1]) if len(args) > 1 else 6
    with_oracle = "--oracle" in sys.argv
    data = named_case(workload)
    grid = ea.GridData.from_opfdata(data)
    scales = 1.0 + 0.03 * np.sin(2 * np.pi * np.arange(T) / max(T, 2))
    tmp = Path(tempfile.mkdtemp())
    np.savetxt(tmp / "p.Pd", np.outer(grid.Pd, scales), fmt="%.17g")
    np.savetxt(tmp / "p.Qd", np.outer(grid.Qd, scales), fmt="%.17g")
    rho_pq, rho_va = 4e2, 4e4
    out = {"workload": workload, "T": T, "nline": grid.nline, "ngen": grid.ngen}

    def params(p):
        p.verbose = 0; p.outer_iterlim = 3; p.inner_iterlim = 200; p.scale = 1e-4

    for rep in range(2):                                     # first pass warms the context / memory pool
        env = AdmmEnv(data, rho_pq, rho_va, use_gpu=True, verbose=0, load_prefix=str(tmp / "p"))
        params(env.params)
        mod = ModelMpacopf(env, start_period=1, end_period=T)
        t0 = time.perf_counter()
        admm_two_level(env, mod, None, mode="native")
        wall = time.perf_counter() - t0
        kt = np.zeros(4)
        mod.lib.ea_mp_get_kernel_times(mod.h, kt.ctypes.data_as(__import__("ctypes").POINTER(__import__("ctypes").c_double)))
        out.update(gpu_wall_s=wall, gpu_device_s=float(kt[0]), launches=int(kt[1]), iterations=int(mod.info.cumul),
                   outer=int(mod.info.outer), objval=mod.info.objval, mismatch=mod.info.mismatch,
                   gpu_it_per_s=mod.info.cumul / wall, gpu_period_it_per_s=T * mod.info.cumul / wall)
        mod.close()

    if with_oracle:
        from oracle.oracle import OracleMpModel
        import os
        par = Parameters(); params(par)
        om = OracleMpModel(grid, par, rho_pq, rho_va, np.outer(scales, grid.Pd), np.outer(scales, grid.Qd))
        om.set_threads(os.cpu_count())
        t0 = time.perf_counter()
        oi = om.admm_two_level()
        ow = time.perf_counter() - t0
        out.update(oracle_wall_s=ow, oracle_threads=os.cpu_count(), oracle_iterations=int(oi.cumul), oracle_objval=oi.objval,
                   oracle_it_per_s=oi.cumul / ow, speedup=ow / out["gpu_wall_s"] * out["iterations"] / max(oi.cumul, 1))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
