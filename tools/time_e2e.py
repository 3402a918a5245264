import ctypes as C, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import bench
from exaadmm_b200 import capi
from exaadmm_b200.capi import EaInfo, dptr, make_grid_struct, params_struct
lib = capi.load_library()
grid, _ = bench.make_grid("ACTIVSg70k"); par, rho_pq, rho_va = bench.default_params("ACTIVSg70k")
gs, keep = make_grid_struct(grid); nvar = 2*grid.ngen + 8*grid.nline
for rep in range(3):
    t=[time.perf_counter()]
    h = C.c_void_p(); lib.ea_create(C.byref(gs), 0, C.byref(h)); t.append(time.perf_counter())
    lib.ea_init_solution(h, rho_pq, rho_va); t.append(time.perf_counter())
    lib.ea_set_option(h, b"count_work", 0.0)
    info = EaInfo(); p = params_struct(par); lib.ea_admm_two_level(h, C.byref(p), C.byref(info)); t.append(time.perf_counter())
    u = np.empty(nvar); lib.ea_get_vector(h, 0, dptr(u), nvar); t.append(time.perf_counter())
    lib.ea_destroy(h); t.append(time.perf_counter())
    print("create %.1f ms, init %.1f ms, two_level %.1f ms (solver %.1f), get_vector %.1f ms, destroy %.1f ms" % (
        1e3*(t[1]-t[0]), 1e3*(t[2]-t[1]), 1e3*(t[3]-t[2]), 1e3*info.time_overall, 1e3*(t[4]-t[3]), 1e3*(t[5]-t[4])))
