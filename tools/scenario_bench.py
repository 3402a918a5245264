#!/usr/bin/env python
"""Aggregate inner iterations/s of S concurrent load scenarios per GPU (BASELINE config 5).
usage: [torchrun ...] python tools/scenario_bench.py [workload] [S ...]"""
import json, os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench
from exaadmm_b200.scenarios import solve_scenarios

wl = sys.argv[1] if len(sys.argv) > 1 else "case2869pegase"
counts = [int(a) for a in sys.argv[2:]] or [1, 2, 4, 8]
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
_, data = bench.make_grid(wl)
par, rho_pq, rho_va = bench.default_params(wl)
if len(sys.argv) > 1 and wl.startswith("case"):
    rho_pq, rho_va = 4e2, 4e4          # the README's 1e1/1e3 stalls on the synthetic stand-ins
kw = dict(rho_pq=rho_pq, rho_va=rho_va, scale=par.scale, tight_factor=0.99, outer_iterlim=20, inner_iterlim=1000, gpu_no=local)
solve_scenarios(data, [0], **kw)                                         # warm-up
for S in counts:
    ids = [rank + world * k for k in range(S)]
    res, wall = solve_scenarios(data, ids, **kw)
    its = sum(m.info.cumul for _, m in res)
    print(json.dumps({"workload": wl, "rank": rank, "scenarios_per_gpu": S, "wall_s": wall, "iterations": its,
                      "iterations_per_s": its / wall, "solved": sum(m.info.status == "Solved" for _, m in res)}))
    for _, m in res:
        m.close()
