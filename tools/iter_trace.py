#!/usr/bin/env python
"""Per-iteration trace of a whole two-level solve: for every inner iteration the wall time of the fused iteration and
the work of its x-update (largest evaluation count of one branch, total evaluations).
usage: python tools/iter_trace.py [workload] [max_outer]   -> one line per iteration on stdout, a summary at the end"""
import ctypes as C
import math
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import bench
from exaadmm_b200 import capi
from exaadmm_b200.capi import dptr, make_grid_struct, EaCounters

wl = sys.argv[1] if len(sys.argv) > 1 else "ACTIVSg70k"
max_outer = int(sys.argv[2]) if len(sys.argv) > 2 else 20
lib = capi.load_library()
grid, _ = bench.make_grid(wl)
par, rho_pq, rho_va = bench.default_params(wl)
gs, keep = make_grid_struct(grid)
h = C.c_void_p()
assert lib.ea_create(C.byref(gs), 0, C.byref(h)) == 0
assert lib.ea_init_solution(h, rho_pq, rho_va) == 0
nvar = 2 * grid.ngen + 8 * grid.nline
sqrt_d = math.sqrt(nvar)
outer_tol = sqrt_d * par.outer_eps
beta = par.initial_beta
res = np.zeros(4); got = C.c_int64(); nz = C.c_double(); c = EaCounters()
rows = []
outer = 0
while outer < min(par.outer_iterlim, max_outer):
    outer += 1
    lib.ea_outer_prestep(h, C.byref(nz))
    inner = 0
    eps_pri = sqrt_d / (2500 * outer)
    while inner < par.inner_iterlim:
        lib.ea_reset_counters(h)
        t = time.perf_counter()
        lib.ea_run_inner_from(h, outer, beta, inner, inner + 1, par.max_auglag, par.mu_max, par.scale, 1, C.byref(got), dptr(res))
        dt = time.perf_counter() - t
        inner += 1
        lib.ea_get_counters(h, C.byref(c))
        rows.append((outer, inner, 1e6 * dt, c.max_evals_lane, c.tron_evals))
        if res[0] <= eps_pri:
            break
    if res[3] <= outer_tol:
        break
    lib.ea_update_lz(h, beta, par.MAX_MULTIPLIER)
    if res[2] > par.theta * nz.value:
        beta = min(par.inc_c * beta, 1e24)
lib.ea_destroy(h)
a = np.array(rows)
for r in rows:
    print("%3d %4d %8.1f us  max evals %4d  evals %8d" % r)
tot = a[:, 2].sum()
print(f"# {len(rows)} iterations, {tot/1e3:.1f} ms (one host call per iteration: ~25 us each of overhead)", file=sys.stderr)
for lo, hi in ((1, 5), (6, 10), (11, 20), (21, 40), (41, 10**9)):
    m = (a[:, 1] >= lo) & (a[:, 1] <= hi)
    if m.any():
        print(f"# inner {lo:>3}-{min(hi, int(a[:,1].max())):<4}: {m.sum():4d} iterations, {a[m, 2].sum()/1e3:7.1f} ms "
              f"({100*a[m,2].sum()/tot:4.1f} %), mean {a[m,2].mean():7.1f} us, mean max-evals {a[m,3].mean():6.1f}", file=sys.stderr)
