#!/usr/bin/env python
"""Run W + K fused inner iterations of a workload (for ncu / quick timing).
usage: python tools/profile_iter.py [workload] [warmup] [steps] [chunk]"""
import ctypes as C
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
W = int(sys.argv[2]) if len(sys.argv) > 2 else 40
K = int(sys.argv[3]) if len(sys.argv) > 3 else 20
chunk = int(sys.argv[4]) if len(sys.argv) > 4 else 16
lib = capi.load_library()
grid, _ = bench.make_grid(wl)
par, rho_pq, rho_va = bench.default_params(wl)
gs, keep = make_grid_struct(grid)
h = C.c_void_p()
assert lib.ea_create(C.byref(gs), 0, C.byref(h)) == 0, lib.ea_last_error(None)
assert lib.ea_init_solution(h, rho_pq, rho_va) == 0
lib.ea_set_option(h, b"chunk", float(chunk))
lib.ea_set_option(h, b"count_work", float(os.environ.get("EA_COUNT_WORK", "1")))
import os
lib.ea_set_option(h, b"use_graph", float(os.environ.get("EA_USE_GRAPH", "1")))
res = np.zeros(4); got = C.c_int64(); nz = C.c_double()
lib.ea_outer_prestep(h, C.byref(nz))
# eps_pri is never met with outer = huge -> run exactly the requested number of iterations
lib.ea_run_inner_from(h, 10**9, par.initial_beta, 0, W, par.max_auglag, par.mu_max, par.scale, chunk, C.byref(got), dptr(res))
lib.ea_reset_counters(h)
lib.ea_set_option(h, b"kernel_timing", float(os.environ.get("EA_KERNEL_TIMING", "1")))
t = time.perf_counter()
lib.ea_run_inner_from(h, 10**9, par.initial_beta, W, W + K, par.max_auglag, par.mu_max, par.scale, chunk, C.byref(got), dptr(res))
dt = time.perf_counter() - t
kt = (C.c_double * 8)(); lib.ea_get_kernel_times(h, kt)
c = EaCounters(); lib.ea_get_counters(h, C.byref(c))
print(f"{wl}: {K} iterations after {W}: wall {1e3*dt/K:.3f} ms/iter, span {1e3*kt[0]/K:.3f} ms/iter, "
      f"x-update {1e6*kt[2]/max(kt[1],1):.1f} us, bus {1e6*kt[4]/max(kt[3],1):.1f} us; res {res}")
print(c.as_dict())
# phase split of ONE x-update (queue-empty time vs end), diagnostics
lib.ea_set_option(h, b"count_work", 2.0)
for rep in range(3):
    lib.ea_reset_counters(h)
    lib.ea_run_inner_from(h, 10**9, par.initial_beta, W + K + rep, W + K + rep + 1, par.max_auglag, par.mu_max, par.scale, 1, C.byref(got), dptr(res))
    lib.ea_get_kernel_times(h, kt); lib.ea_get_counters(h, C.byref(c))
    print(f"  x-update phase split: queue empty after {1e6*kt[6]:.1f} us, kernel end {1e6*kt[7]:.1f} us, max evals/line {c.max_evals_lane}")
lib.ea_destroy(h)
