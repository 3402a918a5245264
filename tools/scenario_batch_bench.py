#!/usr/bin/env python
"""Config 5 exploration: S load scenarios of a case on one GPU - batched (ea_batch_*), one after the other, and as
concurrent stand-alone solves on separate streams. usage: scenario_batch_bench.py [workload] [S] [rho_pq rho_va] [spread]"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import bench
from exaadmm_b200.scenarios import ScenarioBatch, solve_scenarios
wl = sys.argv[1] if len(sys.argv) > 1 else "case2869pegase"
S = int(sys.argv[2]) if len(sys.argv) > 2 else 8
_, data = bench.make_grid(wl); par, rho_pq, rho_va = bench.default_params(wl)
if len(sys.argv) > 4:
    rho_pq, rho_va = float(sys.argv[3]), float(sys.argv[4])
spread = float(sys.argv[5]) if len(sys.argv) > 5 else 0.05
kw = dict(scale=par.scale, outer_iterlim=20, inner_iterlim=1000)
ids = list(range(S))
for rep in range(2):
    b = ScenarioBatch(data, ids, rho_pq=rho_pq, rho_va=rho_va, tight_factor=0.99, spread=spread)
    wall = b.solve(**kw)
    cum = [m.info.cumul for m in b.models]; st = [m.info.status for m in b.models]
    b.close()
print(f"{wl} x {S} (rho {rho_pq}/{rho_va}, spread {spread}): batched {wall:.3f} s, {sum(cum)} iterations = {sum(cum)/wall:.0f} it/s; "
      f"solved {st.count('Solved')}/{S}; cumul min/max {min(cum)}/{max(cum)}")
for conc, name in ((1, "one after the other"), (S, "concurrent streams")):
    res, wall2 = solve_scenarios(data, ids, rho_pq=rho_pq, rho_va=rho_va, tight_factor=0.99, max_concurrent=conc, spread=spread, **kw)
    cum2 = [m.info.cumul for _, m in res]
    for _, m in res: m.close()
    print(f"   {name}: {wall2:.3f} s = {sum(cum2)/wall2:.0f} it/s (same counts: {cum2 == cum})")
