"""Independent load scenarios of one case solved concurrently (BASELINE config 5;
SURVEY.md §8e "scenario batch": independent replicas, no data-path collective).

One handle (own CUDA stream, own HBM state) per scenario; the handles are driven from a
thread pool — ctypes releases the GIL for the duration of `ea_admm_two_level`, so the
streams run concurrently on the device. This is what keeps a B200 busy: a single case
leaves most SMs idle while the slowest branch finishes its augmented-Lagrangian loop
(DESIGN.md section 6); other scenarios' kernels fill those SMs. Across GPUs, scenarios
are sharded (rank r takes scenarios r, r+world, ...).
"""
from __future__ import annotations

import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from .admm_two_level import admm_two_level
from .environment import AdmmEnv
from .grid_data import GridData
from .model import ModelAcopf


def scenario_loads(grid: GridData, scenario: int, seed: int | None = None, spread: float = 0.05):
    """Loads of scenario s: every (Pd, Qd) scaled by iid U[1-spread, 1+spread], seed = nbus + s
    (SURVEY.md §8d)."""
    rng = np.random.default_rng((grid.nbus if seed is None else seed) + scenario)
    f = rng.uniform(1.0 - spread, 1.0 + spread, grid.nbus)
    return grid.Pd * f, grid.Qd * f


def solve_scenarios(case, scenarios, *, rho_pq=400.0, rho_va=40000.0, scale=1e-4, obj_scale=1.0, outer_iterlim=20,
                    inner_iterlim=1000, tight_factor=1.0, outer_eps=2e-4, gpu_no=0, max_concurrent=None, spread=0.05):
    """Solve the given scenario ids of `case` concurrently on one GPU.
    Returns (list of (env, mod) in the order of `scenarios`, wall seconds of the concurrent solve)."""
    scenarios = list(scenarios)
    envs, mods = [], []
    for s in scenarios:
        env = AdmmEnv(case if not envs else envs[0].data, rho_pq, rho_va, use_gpu=True, tight_factor=tight_factor,
                      gpu_no=gpu_no, verbose=0)
        mod = ModelAcopf(env)
        Pd, Qd = scenario_loads(mod.grid_data, s, spread=spread)
        mod.set_load(Pd, Qd)
        mod.set_option("count_work", 0)
        p = env.params
        p.scale, p.obj_scale, p.outer_eps, p.outer_iterlim, p.inner_iterlim, p.verbose = \
            scale, obj_scale, outer_eps, outer_iterlim, inner_iterlim, 0
        envs.append(env); mods.append(mod)
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=max_concurrent or len(scenarios)) as ex:
        list(ex.map(lambda em: admm_two_level(em[0], em[1], None, mode="native"), zip(envs, mods)))
    wall = time.perf_counter() - t0
    return list(zip(envs, mods)), wall


class ScenarioBatch:
    """S load scenarios of one case solved TOGETHER on one GPU (`ea_batch_*`, include/exaadmm_b200.h): one branch
    kernel per inner iteration over all (scenario, branch) pairs of the scenarios still iterating - S times the bulk,
    one tail -, one bus kernel with a termination test per scenario. `models[s]` are ordinary `ModelAcopf` objects on
    borrowed handles: `models[s].solution.u_curr`, `.membuf`, `.info` as after a stand-alone `solve_acopf` of that
    scenario (same iterates, bit for bit)."""

    def __init__(self, case, scenarios, *, rho_pq=400.0, rho_va=40000.0, tight_factor=1.0, gpu_no=0, spread=0.05):
        import copy
        import ctypes as C
        from . import capi
        self.lib = capi.load_library()
        self.scenarios = list(scenarios)
        self.env = AdmmEnv(case, rho_pq, rho_va, use_gpu=True, tight_factor=tight_factor, gpu_no=gpu_no, verbose=0)
        self.grid = g = GridData.from_opfdata(self.env.data, tight_factor=tight_factor)
        loads = [scenario_loads(g, s, spread=spread) for s in self.scenarios]
        self.Pd = np.ascontiguousarray([l[0] for l in loads])
        self.Qd = np.ascontiguousarray([l[1] for l in loads])
        gs, keep = capi.make_grid_struct(g)
        h = C.c_void_p()
        rc = self.lib.ea_batch_create(C.byref(gs), int(gpu_no), len(self.scenarios), capi.dptr(self.Pd), capi.dptr(self.Qd),
                                      C.byref(h))
        if rc != 0:
            raise capi.EaError(rc, (self.lib.ea_batch_last_error(None) or b"").decode())
        self.h = h
        self._check(self.lib.ea_batch_init_solution(self.h, rho_pq, rho_va))
        self.models = []
        for k in range(len(self.scenarios)):
            gk = copy.copy(g)
            gk.Pd, gk.Qd = self.Pd[k].copy(), self.Qd[k].copy()
            self.models.append(ModelAcopf.borrowed(self.env, C.c_void_p(self.lib.ea_batch_scenario(self.h, k)), gk))

    def _check(self, rc):
        if rc != 0:
            from . import capi
            raise capi.EaError(rc, (self.lib.ea_batch_last_error(self.h) or b"").decode())

    def solve(self, *, scale=1e-4, obj_scale=1.0, outer_iterlim=20, inner_iterlim=1000, outer_eps=2e-4):
        """admm_two_level of every scenario; fills `models[s].info`. Returns the wall seconds of the batch."""
        import ctypes as C
        from .capi import EaInfo, STATUS_NAMES, params_struct
        p = self.env.params
        p.scale, p.obj_scale, p.outer_eps, p.outer_iterlim, p.inner_iterlim, p.verbose = \
            scale, obj_scale, outer_eps, outer_iterlim, inner_iterlim, 0
        ps = params_struct(p)
        infos = (EaInfo * len(self.models))()
        t0 = time.perf_counter()
        self._check(self.lib.ea_batch_admm_two_level(self.h, C.byref(ps), infos))
        wall = time.perf_counter() - t0
        for m, out in zip(self.models, infos):
            m.info.status = STATUS_NAMES[out.status]
            for k in ("inner", "outer", "cumul", "objval", "primres", "dualres", "mismatch", "eps_pri", "norm_z_curr",
                      "norm_z_prev", "time_overall"):
                setattr(m.info, k, getattr(out, k))
        return wall

    def close(self):
        if getattr(self, "h", None):
            self.lib.ea_batch_destroy(self.h)
            self.h = None
            for m in self.models:
                m.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
