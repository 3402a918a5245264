#!/usr/bin/env python
"""The timed window of bench.py (iterations W+1 .. W+K of the 70k solve, kernels bracketed by CUDA events) under the
three cache regimes: no flush (working set L2-resident), 256 MB write before every iteration (dirty L2), write + read
of a second buffer (cold, clean L2).  usage: python tools/flush_modes.py [workload] [warmup] [steps]"""
import ctypes as C
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import bench
from exaadmm_b200 import capi
from exaadmm_b200.capi import dptr, make_grid_struct

wl = sys.argv[1] if len(sys.argv) > 1 else "ACTIVSg70k"
W = int(sys.argv[2]) if len(sys.argv) > 2 else 5
K = int(sys.argv[3]) if len(sys.argv) > 3 else 20
lib = capi.load_library()
grid, _ = bench.make_grid(wl)
par, rho_pq, rho_va = bench.default_params(wl)
gs, keep = make_grid_struct(grid)
for name, mb, clean in (("no flush", 0, 0), ("write 256 MB", 256, 0), ("write 256 MB + read 256 MB", 256, 1)):
    h = C.c_void_p()
    assert lib.ea_create(C.byref(gs), 0, C.byref(h)) == 0, lib.ea_last_error(None)
    assert lib.ea_init_solution(h, rho_pq, rho_va) == 0
    res = np.zeros(4); got = C.c_int64(); nz = C.c_double()
    lib.ea_outer_prestep(h, C.byref(nz))
    lib.ea_run_inner_from(h, 10**9, par.initial_beta, 0, W, par.max_auglag, par.mu_max, par.scale, 16, C.byref(got), dptr(res))
    lib.ea_reset_counters(h)
    lib.ea_set_option(h, b"kernel_timing", 1.0)
    lib.ea_set_option(h, b"l2_flush_mb", float(mb))
    lib.ea_set_option(h, b"l2_flush_clean", float(clean))
    lib.ea_run_inner_from(h, 10**9, par.initial_beta, W, W + K, par.max_auglag, par.mu_max, par.scale, 16, C.byref(got), dptr(res))
    kt = (C.c_double * 8)(); lib.ea_get_kernel_times(h, kt)
    print(f"{wl} iterations {W + 1}-{W + K}, {name:28s}: x-update {1e6 * kt[2] / max(kt[1], 1):7.1f} us, bus {1e6 * kt[4] / max(kt[3], 1):6.1f} us")
    lib.ea_destroy(h)
