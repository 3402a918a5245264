"""Why does tron::newton_step decline? CPU analysis tool.

Drives the CPU oracle on a synthetic grid with BASELINE.json's parameters, replays every branch of selected inner
iterations through the product's device code compiled for the host with outcome counters (tests/harness,
-DEA_STATS -DEA_FAST_EVALS=1: the direct step is attempted on every TRON step, also the first of a branch) and
prints, separately for the first step of a branch and for later steps, how the attempts ended.

Usage: python tools/step_stats.py [--workload case13659pegase] [--iters 2,8,50]
"""
import argparse
import ctypes as C
import subprocess
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import exaadmm_b200 as ea  # noqa: E402
from exaadmm_b200.synthetic import synthetic_case, NAMED  # noqa: E402
from exaadmm_b200.environment import Parameters  # noqa: E402
from oracle.oracle import OracleModel  # noqa: E402
from conftest import branch_inputs  # noqa: E402

pd = C.POINTER(C.c_double)
REASONS = ["declined: no variable can move", "(trials beyond the first break point)", "(Cauchy point beyond the first break point)",
           "declined: no free variable", "declined: Cholesky pivot <= 0", "declined: zero residual",
           "(faces with a projected search)", "declined: unclipped step leaves a residual", "ok: one face",
           "ok: several faces", "(clipped by vi)", "(by vj)", "(by ti)", "(by tj) / declined: more than 3 faces", "(by s_ij)", "(by s_ji)"]


def P(a):
    return a.ctypes.data_as(pd)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="case13659pegase")
    ap.add_argument("--iters", default="2,8,50")
    ap.add_argument("--flags", default="")
    args = ap.parse_args()
    hdir = ROOT / "tests" / "harness"
    so = hdir / "_build_host_harness_stats.so"
    subprocess.run(["nvcc", "-O2", "-std=c++17", "-ccbin", "/usr/bin/g++", "-Wno-deprecated-gpu-targets", "-Xcompiler",
                    "-fPIC,-ffp-contract=off", "-shared", "-DEA_STATS", "-DEA_FAST_EVALS=1", *args.flags.split(),
                    "-o", str(so), str(hdir / "host_harness.cu")], check=True)
    H = C.CDLL(str(so))
    H.hh_solve_branch.argtypes = [pd, pd, pd, pd, pd, C.c_longlong, C.c_int, C.c_double, C.c_double, pd, C.POINTER(C.c_int)]
    nbus, ngen, nline = NAMED[args.workload]
    case = synthetic_case(nbus, ngen, nline, seed=nbus)
    grid = ea.GridData.from_opfdata(case, tight_factor=0.99)
    par = Parameters(); par.verbose = 0
    rho_pq, rho_va, par.scale = {"ACTIVSg70k": (3e4, 3e5, 1e-5), "case13659pegase": (5e1, 5e3, 1e-4),
                                 "case2869pegase": (1e1, 1e3, 1e-4), "case1354pegase": (1e1, 1e3, 1e-4)}[args.workload]
    m = OracleModel(grid, par, rho_pq, rho_va)
    m.set_threads(8)
    want = sorted(int(s) for s in args.iters.split(","))
    m.admm_increment_outer(); m.admm_outer_prestep(); m.admm_increment_reset_inner()
    t0 = time.time()
    stats = (C.c_longlong * 32)()
    for it in range(1, want[-1] + 1):
        m.admm_increment_inner(); m.admm_inner_prestep(); m.admm_update_x_gen()
        if it in want:
            u = m.vec("u_curr").copy(); v = m.vec("v_curr").copy(); z = m.vec("z_curr").copy()
            l = m.vec("l_curr").copy(); rho = m.vec("rho").copy(); mb = m.membuf().copy()
            H.hh_stats(stats, 1)
            ev = np.zeros(grid.nline, dtype=np.int64)
            for I in range(grid.nline):
                x, xl, xu, param, Y = branch_inputs(grid, u, v, z, l, rho, mb, I)
                F = np.zeros(4); work = (C.c_int * 6)()
                H.hh_solve_branch(P(x), P(xl), P(xu), P(param), P(Y), m.inner, par.max_auglag, par.mu_max, par.scale,
                                  P(F), work)
                ev[I] = work[1]
            H.hh_stats(stats, 1)
            s = np.array(list(stats)).reshape(2, 16)
            print(f"iteration {it}: {grid.nline} branches, evaluations mean {ev.mean():.2f} max {ev.max()} "
                  f"[{time.time() - t0:.1f}s]")
            for f, name in ((1, "first step of a branch"), (0, "later steps")):
                tot = s[f, :10].sum()
                print(f"  {name}: {tot} attempts")
                for k in range(16):
                    if s[f, k]:
                        print(f"    {REASONS[k]:42s} {s[f, k]:8d}  {100.0 * s[f, k] / max(tot, 1):6.2f} %")
        m.admm_update_x_line(); m.admm_update_xbar(); m.admm_update_z(); m.admm_update_l(); m.admm_update_residual()


if __name__ == "__main__":
    main()
