#!/usr/bin/env python
"""torchrun entry: time K fused inner iterations of a bus-partitioned case (per-iteration cost of the exchange)."""
import ctypes as C, os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch, torch.distributed as dist
import bench, exaadmm_b200 as ea
from exaadmm_b200.capi import dptr
from exaadmm_b200.environment import AdmmEnv
from exaadmm_b200.partition import partition_buses
from exaadmm_b200.partitioned import make_partitioned_model, init_comm, init_peer_exchange

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
wl = sys.argv[1] if len(sys.argv) > 1 else "ACTIVSg70k"
W, K = 40, 100
torch.cuda.set_device(local); dist.init_process_group("nccl", device_id=torch.device("cuda", local))
_, data = bench.make_grid(wl); par, rho_pq, rho_va = bench.default_params(wl)
env = AdmmEnv(data, rho_pq, rho_va, use_gpu=True, tight_factor=0.99, gpu_no=local, verbose=0)
grid = ea.GridData.from_opfdata(data, tight_factor=0.99)
mod, lg = make_partitioned_model(env, grid, partition_buses(grid, world), rank)
init_comm(mod, rank)
if len(sys.argv) > 2 and sys.argv[2] == 'peer':
    init_peer_exchange(mod)
lib, h = mod.lib, mod.h
lib.ea_set_option(h, b"count_work", 0.0)
res = np.zeros(4); got = C.c_int64(); nz = C.c_double()
lib.ea_outer_prestep(h, C.byref(nz))
lib.ea_run_inner_from(h, 10**9, par.initial_beta, 0, W, par.max_auglag, par.mu_max, par.scale, 16, C.byref(got), dptr(res))
lib.ea_reset_counters(h); lib.ea_set_option(h, b"kernel_timing", 1.0)
dist.barrier(); t = time.perf_counter()
lib.ea_run_inner_from(h, 10**9, par.initial_beta, W, W + K, par.max_auglag, par.mu_max, par.scale, 16, C.byref(got), dptr(res))
dt = time.perf_counter() - t
kt = (C.c_double * 8)(); lib.ea_get_kernel_times(h, kt)
print(f"rank {rank}/{world} {wl}: wall {1e3*dt/K:.3f} ms/iter, x-update {1e6*kt[2]/max(kt[1],1):.1f} us, "
      f"bus+exchange+finish {1e6*kt[4]/max(kt[3],1):.1f} us, lines {lg.grid.nline} (cut {lg.stats['cut_lines']}), res {res[:2]}")
dist.barrier(); mod.close(); dist.destroy_process_group()
