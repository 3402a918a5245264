"""Multi-period ACOPF: ``ModelMpacopf``, ``SolutionRamping`` and ``solve_mpacopf``.

Host mirrors of ``/root/reference/src/models/mpacopf/mpacopf_model.jl:1-109`` and
``/root/reference/src/interface/solve_mpacopf.jl:1-37``. The periods are ordinary
``ModelAcopf`` objects whose handles are owned by the multi-period handle
(``ea_mp_period``), so ``mod.models[i].solution.u_curr`` etc. work as in the reference's
tests (``test/algorithms/mpacopf_update_gpu.jl``); the ramp-coupling vectors are
``mod.solution[i].u_curr`` ... (``SolutionRamping``). Periods are 0-based here.
Only the ``use_gpu=true`` path exists (no CPU fallback).
"""
from __future__ import annotations

import copy
import ctypes as C

import numpy as np

from . import capi
from .capi import EaError, EaInfo, RAMP_FIELDS, STATUS_NAMES, dptr, params_struct
from .environment import AdmmEnv, IterationInformation
from .grid_data import GridData
from .model import ModelAcopf


class SolutionRamping:
    """``SolutionRamping{T,TD}`` (mpacopf_model.jl:1-16): attribute access = D2H copy of ngen doubles
    in the reference's generator order; assignment uploads."""

    def __init__(self, model: "ModelMpacopf", t: int):
        object.__setattr__(self, "_m", model)
        object.__setattr__(self, "t", t + 1)                       # 1-based like the reference
        object.__setattr__(self, "len_horizon", model.len_horizon)

    def __getattr__(self, name):
        if name in RAMP_FIELDS:
            return self._m.get_ramp_vector(self.t - 1, name)
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if name not in RAMP_FIELDS:
            raise AttributeError(name)
        self._m.set_ramp_vector(self.t - 1, name, value)


class ModelMpacopf:
    """``ModelMpacopf{T,TD,TI,TM}(env; start_period, end_period, ramp_ratio)`` on one B200."""

    is_multiperiod = True

    def __init__(self, env: AdmmEnv, start_period: int = 1, end_period: int = 1, ramp_ratio: float = 0.02,
                 grid: GridData | None = None):
        if not env.use_gpu:
            raise NotImplementedError("exaadmm_b200 implements the use_gpu=true path only (no CPU fallback)")
        if not env.load_specified:
            raise ValueError("a load profile is required (load_prefix)")                      # mpacopf_model.jl:59
        if not (1 <= start_period <= end_period <= env.load.pd.shape[1]):
            raise ValueError("period range outside the load profile")                         # :60
        self.lib = capi.load_library()
        self.env = env
        self.len_horizon = end_period - start_period + 1
        self.ramp_ratio = ramp_ratio
        g = grid if grid is not None else GridData.from_opfdata(env.data, tight_factor=env.tight_factor)
        Pd = np.ascontiguousarray(env.load.pd[:, start_period - 1:end_period].T, dtype=np.float64)   # T x nbus
        Qd = np.ascontiguousarray(env.load.qd[:, start_period - 1:end_period].T, dtype=np.float64)
        if Pd.shape[1] != g.nbus:
            raise ValueError("load profile rows do not match the number of buses")
        gs, keep = capi.make_grid_struct(g)
        h = C.c_void_p()
        rc = self.lib.ea_mp_create(C.byref(gs), int(env.gpu_no), self.len_horizon, dptr(Pd), dptr(Qd), float(ramp_ratio),
                                   env.initial_rho_pq, env.initial_rho_va, C.byref(h))
        if rc != 0:
            raise EaError(rc, (self.lib.ea_mp_last_error(None) or b"").decode())
        self.h = h
        del keep
        self.ngen = g.ngen
        self.models = []
        for t in range(self.len_horizon):
            gt = copy.copy(g)
            m = ModelAcopf.borrowed(env, C.c_void_p(self.lib.ea_mp_period(self.h, t)), gt, ramp_ratio=ramp_ratio)
            m.grid_data.Pd, m.grid_data.Qd = Pd[t].copy(), Qd[t].copy()
            self.models.append(m)
        self.solution = [SolutionRamping(self, t) for t in range(self.len_horizon)]
        self.nvar = int(self.lib.ea_mp_nvar(self.h))                                           # :97-102
        self.info = IterationInformation()
        self.kernel_seconds = 0.0

    # -- plumbing ---------------------------------------------------------------
    def _check(self, rc: int):
        if rc != 0:
            raise EaError(rc, (self.lib.ea_mp_last_error(self.h) or b"").decode())

    def close(self):
        if getattr(self, "h", None):
            for m in self.models:
                m.h = None                       # borrowed: freed with the multi-period handle
            self.lib.ea_mp_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def get_ramp_vector(self, t: int, name: str) -> np.ndarray:
        out = np.empty(self.ngen, dtype=np.float64)
        self._check(self.lib.ea_mp_get_ramp_vector(self.h, t, RAMP_FIELDS[name], dptr(out), self.ngen))
        return out

    def set_ramp_vector(self, t: int, name: str, value) -> None:
        a = np.ascontiguousarray(value, dtype=np.float64)
        if a.shape != (self.ngen,):
            raise ValueError(f"{name}: expected shape ({self.ngen},), got {a.shape}")
        self._check(self.lib.ea_mp_set_ramp_vector(self.h, t, RAMP_FIELDS[name], dptr(a), self.ngen))

    def gen_membuf(self, t: int) -> np.ndarray:
        """Rows 7 (multiplier) and 8 (penalty) of ``mod.models[t].gen_membuf`` as a (2, ngen) array."""
        out = np.zeros((2, self.ngen))
        for i, row in enumerate((7, 8)):
            r = np.empty(self.ngen)
            self._check(self.lib.ea_mp_get_gen_membuf(self.h, t, row, dptr(r), self.ngen))
            out[i] = r
        return out

    def init_solution(self, rho_pq: float, rho_va: float) -> None:
        """``init_solution!(mod, mod.solution, rho_pq, rho_va)`` (mpacopf_init_solution_gpu.jl:16-35)."""
        self._check(self.lib.ea_mp_init_solution(self.h, rho_pq, rho_va))


# -- the operators for mod::ModelMpacopf (dispatched to from operators.py) -------------------
def admm_increment_outer(env, mod):
    mod.info.outer += 1
    for m in mod.models:
        m.info.outer += 1


def admm_increment_reset_inner(env, mod):
    mod.info.inner = 0
    for m in mod.models:
        m.info.inner = 0


def admm_increment_inner(env, mod):
    mod.info.inner += 1
    mod.info.cumul += 1
    for m in mod.models:
        m.info.inner += 1
        m.info.cumul += 1


def admm_outer_prestep(env, mod):
    out = C.c_double()
    mod._check(mod.lib.ea_mp_outer_prestep(mod.h, C.byref(out)))
    mod.info.norm_z_prev = out.value


def admm_inner_prestep(env, mod):
    mod._check(mod.lib.ea_mp_inner_prestep(mod.h))


def admm_update_x(env, mod):
    par = env.params
    mod._check(mod.lib.ea_mp_update_x(mod.h, mod.info.inner, par.max_auglag, par.mu_max, par.scale))


def admm_update_xbar(env, mod):
    mod._check(mod.lib.ea_mp_update_xbar(mod.h))


def admm_update_z(env, mod):
    mod._check(mod.lib.ea_mp_update_z(mod.h, env.params.beta))


def admm_update_l(env, mod):
    mod._check(mod.lib.ea_mp_update_l(mod.h, env.params.beta))


def admm_update_lz(env, mod):
    mod._check(mod.lib.ea_mp_update_lz(mod.h, env.params.beta, env.params.MAX_MULTIPLIER))


def admm_update_residual(env, mod):
    out = np.zeros(4)
    mod._check(mod.lib.ea_mp_update_residual(mod.h, dptr(out)))
    info = mod.info
    info.primres, info.dualres, info.norm_z_curr, info.mismatch = (float(x) for x in out)


def admm_poststep(env, mod):
    obj, err = C.c_double(), C.c_double()
    mod._check(mod.lib.ea_mp_poststep(mod.h, C.byref(obj), C.byref(err)))
    mod.info.objval = obj.value
    mod.info.user.err_ramp = err.value


def admm_run_inner(env, mod, chunk: int = 0):
    par = env.params
    out = np.zeros(4)
    done = C.c_int64()
    mod._check(mod.lib.ea_mp_run_inner(mod.h, mod.info.outer, par.beta, par.inner_iterlim, par.max_auglag, par.mu_max,
                                       par.scale, int(chunk), C.byref(done), dptr(out)))
    info = mod.info
    info.inner = int(done.value)
    info.cumul += int(done.value)
    if done.value > 0:
        info.primres, info.dualres, info.norm_z_curr, info.mismatch = (float(x) for x in out)


def admm_two_level_native(env, mod):
    par = env.params
    p = params_struct(par)
    out = EaInfo()
    err = C.c_double()
    mod._check(mod.lib.ea_mp_admm_two_level(mod.h, C.byref(p), C.byref(out), C.byref(err)))
    info = mod.info
    info.status = STATUS_NAMES[out.status]
    for k in ("inner", "outer", "cumul", "objval", "primres", "dualres", "mismatch", "auglag", "eps_pri",
              "norm_z_curr", "norm_z_prev", "time_overall"):
        setattr(info, k, getattr(out, k))
    info.user.err_ramp = err.value
    par.beta = out.beta


def solve_mpacopf(case, load_prefix, *, case_format="matpower", start_period=1, end_period=1, outer_iterlim=20,
                  inner_iterlim=1000, rho_pq=400.0, rho_va=40000.0, obj_scale=1.0, scale=1e-4, use_gpu=False,
                  use_linelimit=True, tight_factor=1.0, outer_eps=2e-4, gpu_no=0, verbose=1, ramp_ratio=0.02,
                  warm_start=True, multiperiod_tight=True, mode="fused"):
    """``solve_mpacopf(case, load_prefix; kwargs...) -> (env, mod)`` (solve_mpacopf.jl:1-37)."""
    from .admm_two_level import admm_two_level
    if not use_gpu:
        raise NotImplementedError("use_gpu=false: the reference's CPU path is not part of this package; pass use_gpu=True")
    env = AdmmEnv(case, rho_pq, rho_va, case_format=case_format, use_gpu=use_gpu, use_linelimit=use_linelimit,
                  load_prefix=load_prefix, tight_factor=tight_factor, gpu_no=gpu_no, verbose=verbose)
    mod = ModelMpacopf(env, start_period=start_period, end_period=end_period, ramp_ratio=ramp_ratio)
    env.params.scale = scale
    env.params.obj_scale = obj_scale
    env.params.outer_eps = outer_eps
    env.params.outer_iterlim = outer_iterlim
    env.params.inner_iterlim = inner_iterlim
    if warm_start:                                    # solve_mpacopf.jl:27-32: every period alone, then re-initialise
        for m in mod.models:
            admm_two_level(env, m, None, mode="native" if mode != "stepwise" else mode)
        mod.init_solution(rho_pq, rho_va)
    admm_two_level(env, mod, None, mode=mode)
    return env, mod
