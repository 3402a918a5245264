import sys, time, ctypes as C
sys.path.insert(0, '/root/repo')
import bench
from exaadmm_b200 import capi
from exaadmm_b200.capi import make_grid_struct
lib = capi.load_library()
grid, _ = bench.make_grid("ACTIVSg70k")
gs, keep = make_grid_struct(grid)
for rep in range(3):
    h = C.c_void_p()
    t = time.perf_counter()
    assert lib.ea_create(C.byref(gs), 0, C.byref(h)) == 0
    print(f"ea_create {1e3*(time.perf_counter()-t):.2f} ms", file=sys.stderr)
    t = time.perf_counter(); lib.ea_destroy(h); print(f"ea_destroy {1e3*(time.perf_counter()-t):.2f} ms", file=sys.stderr)
