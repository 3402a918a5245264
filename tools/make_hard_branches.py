"""Collect HARD branch sub-problems from a real solve (test / tuning fixture).

The x-update of an inner iteration lasts as long as its slowest branch: a branch whose line limit has just become
active walks the augmented-Lagrangian penalty ladder (20-30 TRON solves in a row, 50-200 objective evaluations)
while the average branch needs 3 evaluations. This script drives the CPU oracle on the ACTIVSg70k-like grid with
BASELINE.json's parameters, replays every branch of selected inner iterations through the product's device code
compiled for the host (tests/harness), and keeps the branches that need the most evaluations together with a sample
of ordinary ones.

Output: tests/golden/hard_branches.npz
    prob   n x 59   lam8 rho8 xt8 Y8 | xl6 xu6 | x0(6) | ls0 ls1 mu | major | rateA | (pad)
    sol    n x 13   x(6) F(4) ls0 ls1 mu          (host build of the device code, FMA contraction on)
    work   n x 6    auglag, evals, cg, shifts, rejected, hit_max
    meta   scale, mu_max, max_auglag

Usage: python tools/make_hard_branches.py [--iters 3,12,40] [--keep 192] [--workload ACTIVSg70k]
"""
import argparse
import ctypes as C
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
from conftest import branch_inputs, _load_harness  # noqa: E402

pd = C.POINTER(C.c_double)


def P(a):
    return a.ctypes.data_as(pd)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="ACTIVSg70k")
    ap.add_argument("--iters", default="2,6,15,40")
    ap.add_argument("--keep", type=int, default=160)
    ap.add_argument("--easy", type=int, default=96)
    ap.add_argument("--out", default=str(ROOT / "tests" / "golden" / "hard_branches.npz"))
    args = ap.parse_args()
    nbus, ngen, nline = NAMED[args.workload]
    case = synthetic_case(nbus, ngen, nline, seed=nbus)
    grid = ea.GridData.from_opfdata(case, tight_factor=0.99)
    par = Parameters(); par.verbose = 0
    rho_pq, rho_va, par.scale = {"ACTIVSg70k": (3e4, 3e5, 1e-5), "case13659pegase": (5e1, 5e3, 1e-4),
                                 "case2869pegase": (1e1, 1e3, 1e-4), "case1354pegase": (1e1, 1e3, 1e-4)}[args.workload]
    H = _load_harness("_build_host_harness.so")
    m = OracleModel(grid, par, rho_pq, rho_va)
    m.set_threads(8)
    want = sorted(int(s) for s in args.iters.split(","))
    m.admm_increment_outer(); m.admm_outer_prestep(); m.admm_increment_reset_inner()
    probs, sols, works = [], [], []
    rng = np.random.default_rng(1)
    t0 = time.time()
    for it in range(1, want[-1] + 1):
        m.admm_increment_inner(); m.admm_inner_prestep(); m.admm_update_x_gen()
        if it in want:
            u = m.vec("u_curr").copy(); v = m.vec("v_curr").copy(); z = m.vec("z_curr").copy()
            l = m.vec("l_curr").copy(); rho = m.vec("rho").copy(); mb = m.membuf().copy()
            rows = []
            for I in range(grid.nline):
                x, xl, xu, param, Y = branch_inputs(grid, u, v, z, l, rho, mb, I)
                x0 = x.copy(); p0 = param.copy()
                F = np.zeros(4); work = (C.c_int * 6)()
                H.hh_solve_branch(P(x), P(xl), P(xu), P(param), P(Y), m.inner, par.max_auglag, par.mu_max, par.scale,
                                  P(F), work)
                mu0 = 10.0 if m.inner == 1 else p0[26]
                prob = np.concatenate([p0[0:24], Y, xl, xu, x0, [p0[24], p0[25], mu0, float(m.inner), grid.rateA[I], 0.0]])
                sol = np.concatenate([x, F, param[24:27]])
                rows.append((work[1], prob, sol, np.array(list(work), dtype=np.int64)))
            ev = np.array([r[0] for r in rows])
            order = np.argsort(-ev)
            per = args.keep // len(want)
            pick = list(order[:per]) + list(rng.choice(grid.nline, args.easy // len(want), replace=False))
            for I in pick:
                probs.append(rows[I][1]); sols.append(rows[I][2]); works.append(rows[I][3])
            print(f"iter {it}: evals mean {ev.mean():.2f} max {ev.max()} >16: {(ev > 16).sum()} >32: {(ev > 32).sum()} "
                  f"[{time.time() - t0:.1f}s]", flush=True)
        m.admm_update_x_line(); m.admm_update_xbar(); m.admm_update_z(); m.admm_update_l(); m.admm_update_residual()
    np.savez_compressed(args.out, prob=np.array(probs), sol=np.array(sols), work=np.array(works),
                        meta=np.array([par.scale, par.mu_max, float(par.max_auglag)]))
    print("wrote", args.out, np.array(probs).shape)


if __name__ == "__main__":
    main()
