"""Rolling-horizon ACOPF: warm-started re-solves over a load profile with ramp-limited
generator bounds. Mirrors ``/root/reference/src/interface/solve_acopf_rolling.jl:1-30`` and
``/root/reference/src/models/acopf/acopf_admm_rolling_gpu.jl:1-77``; the hot path is reused
unchanged (``admm_two_level`` never resets the Solution or membuf, so every period starts
from the previous iterate — ``admm_two_level.jl:10-13``).
"""
from __future__ import annotations

from dataclasses import dataclass
from pathlib import Path

import numpy as np

from .admm_two_level import admm_two_level
from .environment import AdmmEnv
from .model import ModelAcopf


@dataclass
class Load:
    """``Load{TM}`` (opfdata.jl:102-119): nbus x nperiods matrices, MW / MVAr."""
    pd: np.ndarray
    qd: np.ndarray


def get_load(prefix, load_scale: float = 1.0) -> Load:
    """``get_load(name)`` (opfdata.jl:121-130): reads ``<prefix>.Pd`` and ``<prefix>.Qd``
    (whitespace-delimited, one row per bus, one column per period)."""
    pd = np.loadtxt(str(prefix) + ".Pd", ndmin=2)           # a single-period file is one column, not one row
    qd = np.loadtxt(str(prefix) + ".Qd", ndmin=2)
    if pd.shape != qd.shape:
        raise ValueError("Pd and Qd profiles differ in shape")
    return Load(pd * load_scale, qd * load_scale)


def update_real_power_current_bounds(mod: ModelAcopf) -> None:
    """``update_real_power_current_bounds`` (acopf_admm_rolling_gpu.jl:1-14): ramp limits around
    the dispatch just computed, clipped to the original bounds."""
    g = mod.grid_data
    pg = mod.solution.u_curr[0:2 * g.ngen:2]
    mod.set_pg_bounds(np.maximum(g.pgmin, pg - g.ramp_rate), np.minimum(g.pgmax, pg + g.ramp_rate))


def admm_restart_rolling(env: AdmmEnv, mod: ModelAcopf, start_period: int = 1, end_period: int = 6,
                         result_file: str | None = None, mode: str = "native"):
    """``admm_restart_rolling`` (acopf_admm_rolling_gpu.jl:16-77). Periods are 1-based like the
    reference. Returns the per-period statistics."""
    if not env.load_specified:
        raise ValueError("a load profile is required (load_prefix)")
    if not (start_period >= 1 and end_period <= env.load.pd.shape[1]):
        raise ValueError("period range outside the load profile")
    stats = []
    io = open(f"{result_file}_tight-factor{env.tight_factor}.txt", "w") if result_file else None
    for t in range(start_period, end_period + 1):
        mod.set_load(env.load.pd[:, t - 1], env.load.qd[:, t - 1])
        admm_two_level(env, mod, None, mode=mode)
        i = mod.info
        rec = dict(period=t, status=i.status, objval=i.objval, mismatch=i.mismatch, cumul=i.cumul,
                   time=i.time_overall + i.time_projection)
        stats.append(rec)
        text = (f" ** Statistics of time period {t}\n"
                f"Status  . . . . . . . . . . . . . . . . . {i.status}\n"
                f"Objective value . . . . . . . . . . . . . {i.objval:.6e}\n"
                f"Residual  . . . . . . . . . . . . . . . . {i.mismatch:.6e}\n"
                f"Cumulative iterations . . . . . . . . . . {i.cumul:5d}\n"
                f"Time (secs) . . . . . . . . . . . . . . . {rec['time']:5.3f}\n")
        if env.params.verbose > 0:
            print(text, end="")
        if io:
            io.write(text); io.flush()
        update_real_power_current_bounds(mod)
    if io:
        io.close()
    return stats


def solve_acopf_rolling(case, load_prefix, *, case_format="matpower", outer_iterlim=20, inner_iterlim=1000,
                        rho_pq=400.0, rho_va=40000.0, obj_scale=1.0, scale=1e-4, use_gpu=False, use_linelimit=True,
                        use_projection=False, tight_factor=0.99, outer_eps=2e-4, gpu_no=0, verbose=1, ramp_ratio=0.02,
                        start_period=1, end_period=6, result_file=None):
    """Same keywords and defaults as the reference (``solve_acopf_rolling.jl:1-8``)."""
    if not use_gpu:
        raise NotImplementedError("use_gpu=false: no CPU fallback on this path; pass use_gpu=True")
    env = AdmmEnv(case, rho_pq, rho_va, case_format=case_format, use_gpu=True, use_linelimit=use_linelimit,
                  use_projection=use_projection, load_prefix=str(load_prefix), tight_factor=tight_factor, gpu_no=gpu_no,
                  verbose=verbose)
    mod = ModelAcopf(env, ramp_ratio=ramp_ratio)
    p = env.params
    p.scale, p.obj_scale, p.outer_eps, p.outer_iterlim, p.inner_iterlim = scale, obj_scale, outer_eps, outer_iterlim, inner_iterlim
    mod.rolling_stats = admm_restart_rolling(env, mod, start_period, end_period, result_file)
    return env, mod
