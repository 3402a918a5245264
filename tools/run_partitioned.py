#!/usr/bin/env python
"""torchrun entry: solve one case bus-partitioned over WORLD_SIZE GPUs and compare with the
single-GPU solve (rank 0).   torchrun --nproc-per-node N tools/run_partitioned.py [workload]"""
import json
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
import torch.distributed as dist

import bench
from exaadmm_b200.partition import partition_buses, cut_statistics
from exaadmm_b200.partitioned import solve_acopf_partitioned
from exaadmm_b200.solve_acopf import solve_acopf
import exaadmm_b200 as ea

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
wl = sys.argv[1] if len(sys.argv) > 1 else "case13659pegase"
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
_, data = bench.make_grid(wl)
par, rho_pq, rho_va = bench.default_params(wl)
if len(sys.argv) > 3:                      # optional rho override: workload rho_pq rho_va
    rho_pq, rho_va = float(sys.argv[2]), float(sys.argv[3])
kw = dict(rho_pq=rho_pq, rho_va=rho_va, scale=par.scale, tight_factor=0.99, outer_iterlim=20, inner_iterlim=1000)
grid = ea.GridData.from_opfdata(data, tight_factor=0.99)
part = partition_buses(grid, world)
results = {}
for exchange in ("nccl", "peer"):
  for rep in range(2):
    dist.barrier(); t0 = time.perf_counter()
    env, mod, lg = solve_acopf_partitioned(data, rank, world, part=part, exchange=exchange, **kw)
    torch.cuda.synchronize(); dist.barrier(); dt = time.perf_counter() - t0
    info = mod.info
    res = dict(status=info.status, outer=info.outer, cumul=info.cumul, objval=info.objval, mismatch=info.mismatch,
               solver_s=info.time_overall, wall_s=dt)
    results[exchange] = res
    mod.close()
if rank == 0:
    out = {"workload": wl, "n_gpus": world, "partition": cut_statistics(grid, part), "partitioned": res,
           "partitioned_nccl_allgather": results["nccl"], "partitioned_peer_stores": results["peer"]}
    t0 = time.perf_counter()
    env1, mod1 = solve_acopf(data, use_gpu=True, verbose=0, mode="native", **kw)
    env1, mod1 = solve_acopf(data, use_gpu=True, verbose=0, mode="native", **kw)
    i1 = mod1.info
    out["single_gpu"] = dict(status=i1.status, outer=i1.outer, cumul=i1.cumul, objval=i1.objval, mismatch=i1.mismatch,
                             solver_s=i1.time_overall)
    out["match"] = all(bool(i1.status == r["status"] and i1.outer == r["outer"] and i1.cumul == r["cumul"]
                            and abs(i1.objval - r["objval"]) <= 1e-9 * abs(i1.objval)) for r in results.values())
    print(json.dumps(out))
dist.barrier()
dist.destroy_process_group()
