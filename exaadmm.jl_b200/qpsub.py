"""``ModelQpsub``, its operator functions, ``admm_one_level`` and ``solve_qpsub`` — the one-level ADMM on the QP
sub-problem of an SQP iteration, on the B200 path.

Host mirrors of ``/root/reference/src/models/qpsub/qpsub_model.jl:33-310``, ``src/algorithms/admm_one_level.jl:1-81``
and ``src/interface/solve_qpsub.jl:1-126``; every operator body is one call through the C ABI (``ea_qp_*`` in
``include/exaadmm_b200.h``). As in the reference, the SQP driver constructs the model, fills in the QP fields (``Hs``,
``LH_1h`` ... ``qpsub_Qd``: plain host arrays here) and calls ``init_solution!`` — which is where the fields go to HBM.
"""
from __future__ import annotations

import ctypes as C
import time

import numpy as np

from . import capi
from .capi import EaError, EaInfo, FIELDS, QP_ARRAYS, STATUS_NAMES, dptr
from .environment import AdmmEnv, IterationInformation
from .grid_data import GridData
from .model import Solution


class _QpFields:
    """What ``make_qpsub_struct`` reads: the reference's field names without the ``qpsub_`` prefix."""

    def __init__(self, m: "ModelQpsub"):
        for n in ("Hs", "LH_1h", "RH_1h", "LH_1i", "RH_1i", "LH_1j", "RH_1j", "LH_1k", "RH_1k", "ls", "us", "line_res"):
            setattr(self, n, getattr(m, n))
        self.pgmax, self.pgmin, self.qgmax, self.qgmin = m.qpsub_pgmax, m.qpsub_pgmin, m.qpsub_qgmax, m.qpsub_qgmin
        self.c1, self.c2, self.Pd, self.Qd = m.qpsub_c1, m.qpsub_c2, m.qpsub_Pd, m.qpsub_Qd


class ModelQpsub:
    """``ModelQpsub{T,TD,TI,TM}(env)`` (qpsub_model.jl:133-310) on one B200."""

    is_qpsub = True

    def __init__(self, env: AdmmEnv, ramp_ratio: float = 0.02, grid: GridData | None = None):
        if not env.use_gpu:
            raise NotImplementedError("exaadmm_b200 implements the use_gpu=true path only; there is no CPU fallback")
        if not env.use_linelimit:
            raise NotImplementedError("ModelQpsub without line limits (n = 4) has no kernels in the reference either")
        self.lib = capi.load_library()
        self.env = env
        self.grid_data = g = grid if grid is not None else GridData.from_opfdata(env.data, tight_factor=env.tight_factor)
        nl, ng, nb = g.nline, g.ngen, g.nbus
        self.n = 6
        self.nline_padded = nl
        self.nvar = self.nvar_padded = self.nvar_u = self.nvar_u_padded = 2 * ng + 8 * nl
        self.gen_start, self.line_start = 1, 2 * ng + 1
        self.pgmin_curr, self.pgmax_curr = g.pgmin.copy(), g.pgmax.copy()
        g.ramp_rate = ramp_ratio * g.pgmax
        self.nvar_v = 2 * ng + 4 * nl + 2 * nb
        self.bus_start = 2 * ng + 4 * nl + 1
        # the QP of the SQP iteration, zero until the driver fills it in (qpsub_model.jl:190-262)
        self.Hs = np.zeros((6 * nl, 6))
        self.LH_1h, self.RH_1h = np.zeros((nl, 4)), np.zeros(nl)
        self.LH_1i, self.RH_1i = np.zeros((nl, 4)), np.zeros(nl)
        self.LH_1j, self.RH_1j = np.zeros((nl, 2)), np.zeros(nl)
        self.LH_1k, self.RH_1k = np.zeros((nl, 2)), np.zeros(nl)
        self.ls, self.us = np.zeros((nl, 6)), np.zeros((nl, 6))
        self.line_res = np.zeros((4, nl))
        self.is_HS_sym, self.is_HS_PSD = np.ones(nl, bool), np.ones(nl, bool)
        self.qpsub_c1, self.qpsub_c2 = np.zeros(ng), np.zeros(ng)
        self.qpsub_pgmax, self.qpsub_pgmin = np.zeros(ng), np.zeros(ng)
        self.qpsub_qgmax, self.qpsub_qgmin = np.zeros(ng), np.zeros(ng)
        self.qpsub_Pd, self.qpsub_Qd = np.zeros(nb), np.zeros(nb)
        # what admm_poststep hands back to the SQP driver (:264-292)
        self.dpg_sol, self.dqg_sol = np.zeros(ng), np.zeros(ng)
        self.dline_var, self.dline_fl = np.zeros((6, nl)), np.zeros((4, nl))
        self.dtheta_sol, self.dw_sol = np.zeros(nb), np.zeros(nb)
        self.dual_infeas = np.full(ng + 6 * nl, 1000.0)
        self.h = None
        self.solution = Solution(self)
        self.gen_solution = None
        self.info = IterationInformation()

    # -- plumbing ---------------------------------------------------------------
    def _check(self, rc: int):
        if rc != 0:
            raise EaError(rc, (self.lib.ea_qp_last_error(self.h) or b"").decode())

    def _need(self):
        if self.h is None:
            raise RuntimeError("ModelQpsub: call init_solution!(mod, mod.solution, rho_pq, rho_va) after filling in the QP")
        return self.h

    def upload(self):
        """(Re)create the device model from the host fields (the reference's fields are device arrays the driver
        assigns before ``init_solution!``)."""
        self.close()
        g = self.grid_data
        gs, keep = capi.make_grid_struct(g)
        ds, keep2 = capi.make_qpsub_struct(_QpFields(self), g.nline, g.ngen, g.nbus)
        h = C.c_void_p()
        rc = self.lib.ea_qp_create(C.byref(gs), C.byref(ds), int(self.env.gpu_no), C.byref(h))
        if rc != 0:
            raise EaError(rc, (self.lib.ea_qp_last_error(None) or b"").decode())
        self.h = h
        del keep, keep2

    def close(self):
        if getattr(self, "h", None):
            self.lib.ea_qp_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def get_vector(self, name: str) -> np.ndarray:
        out = np.empty(self.nvar)
        self._check(self.lib.ea_qp_get_vector(self._need(), FIELDS[name], dptr(out), self.nvar))
        return out

    def set_vector(self, name: str, value) -> None:
        a = np.ascontiguousarray(value, dtype=np.float64)
        if a.shape != (self.nvar,):
            raise ValueError(f"{name}: expected shape ({self.nvar},), got {a.shape}")
        self._check(self.lib.ea_qp_set_vector(self._need(), FIELDS[name], dptr(a), self.nvar))

    def _line_array(self, name: str) -> np.ndarray:
        which, rows = QP_ARRAYS[name]
        nl = self.grid_data.nline
        out = np.empty(rows * nl)
        self._check(self.lib.ea_qp_get_line_array(self._need(), which, dptr(out), rows * nl))
        return out.reshape(nl, rows).T.copy()

    def _set_line_array(self, name: str, value) -> None:
        which, rows = QP_ARRAYS[name]
        nl = self.grid_data.nline
        a = np.ascontiguousarray(np.asarray(value, dtype=np.float64).reshape(rows, nl).T)
        self._check(self.lib.ea_qp_set_line_array(self._need(), which, dptr(a), rows * nl))

    @property
    def v_prev(self) -> np.ndarray:
        return self.get_vector("v_prev")

    sqp_line = property(lambda self: self._line_array("sqp_line"), lambda self, v: self._set_line_array("sqp_line", v))
    qpsub_membuf = property(lambda self: self._line_array("qpsub_membuf"),
                            lambda self, v: self._set_line_array("qpsub_membuf", v))
    # `lambda` is a Python keyword: mod.lambda_ is the reference's mod.lambda
    lambda_ = property(lambda self: self._line_array("lambda"), lambda self, v: self._set_line_array("lambda", v))

    def counters(self) -> dict:
        c = (C.c_int64 * 4)()
        self._check(self.lib.ea_qp_get_counters(self._need(), c))
        return dict(zip(("line_calls", "auglag_iters", "tron_evals", "max_auglag_one_call"), map(int, c)))

    def kernel_times(self) -> dict:
        t = np.zeros(4)
        self._check(self.lib.ea_qp_get_kernel_times(self._need(), dptr(t)))
        return {"device_s": t[0], "iterations": int(t[1]), "launches": int(t[2]), "graph_replays": int(t[3])}

    def set_option(self, name: str, value: float) -> None:
        self._check(self.lib.ea_qp_set_option(self._need(), name.encode(), float(value)))


# -- operator functions (docs/src/dev.md:65-155; qpsub methods) ---------------------------------------------------
def init_solution(mod: ModelQpsub, sol, rho_pq: float, rho_va: float, device=None):
    """``init_solution!(model::ModelQpsub, sol, rho_pq, rho_va)`` (qpsub_init_solution_gpu.jl:49-98)."""
    mod.upload()
    mod._check(mod.lib.ea_qp_init_solution(mod.h, float(rho_pq), float(rho_va)))


def admm_increment_outer(env, mod, device=None):
    mod.info.outer += 1


def admm_increment_reset_inner(env, mod, device=None):
    mod.info.inner = 0


def admm_increment_inner(env, mod, device=None):
    mod.info.inner += 1
    mod.info.cumul += 1


def admm_outer_prestep(env, mod, device=None):
    """Generic method the CPU test of the reference calls (not on the one-level path): ``norm_z_prev = ||z_curr||``."""
    mod.info.norm_z_prev = float(np.linalg.norm(mod.solution.z_curr))


def admm_inner_prestep(env, mod, device=None):
    mod.solution.z_prev = mod.solution.z_curr


def admm_update_x(env, mod, device=None):
    par = env.params
    mod._check(mod.lib.ea_qp_update_x(mod._need(), mod.info.inner, par.max_auglag, par.mu_max, par.scale))


def admm_update_xbar(env, mod, device=None):
    mod._check(mod.lib.ea_qp_update_xbar(mod._need()))


def admm_update_l_single(env, mod, device=None):
    mod._check(mod.lib.ea_qp_update_l_single(mod._need()))


def admm_update_residual(env, mod, device=None):
    out = np.zeros(5)
    mod._check(mod.lib.ea_qp_update_residual(mod._need(), dptr(out)))
    i = mod.info
    i.primres, i.dualres, i.mismatch, i.objval, i.auglag = (float(x) for x in out)


def admm_poststep(env, mod, device=None):
    """qpsub_admm_prepoststep_gpu.jl:9-106: objective / augmented Lagrangian and the step, KKT error and multipliers the
    SQP driver reads from the model."""
    if env.use_projection:
        raise NotImplementedError("use_projection=true: pf_projection has no ModelQpsub method in the reference")
    g = mod.grid_data
    o, a = C.c_double(), C.c_double()
    dw, dt, di = np.zeros(g.nbus), np.zeros(g.nbus), np.zeros(g.ngen + 6 * g.nline)
    mod._check(mod.lib.ea_qp_poststep(mod._need(), C.byref(o), C.byref(a), dptr(dw), dptr(dt), dptr(di)))
    mod.info.objval, mod.info.auglag = o.value, a.value
    mod.dw_sol, mod.dtheta_sol, mod.dual_infeas = dw, dt, di
    u = mod.solution.u_curr
    ls = 2 * g.ngen
    mod.dpg_sol, mod.dqg_sol = u[0:ls:2].copy(), u[1:ls:2].copy()
    mod.dline_var = mod.sqp_line
    mod.dline_fl = u[ls:].reshape(g.nline, 8)[:, :4].T.copy()


def admm_one_level(env, mod, device=None, mode: str = "native"):
    """``admm_one_level(env, mod, device)`` (admm_one_level.jl:1-81).

    ``mode="native"``: the whole loop in one ABI call (3 launches per iteration replayed from a CUDA graph, termination
    test on the device); ``mode="stepwise"``: the reference's loop, one operator call per step."""
    par, info = env.params, mod.info
    sqrt_d = np.sqrt(mod.nvar)
    OUTER_TOL = sqrt_d * par.outer_eps
    info.fill(0)
    info.mismatch = float("inf")
    info.norm_z_prev = info.norm_z_curr = 0
    par.initial_beta = 0
    par.beta = 0
    par.inner_iterlim = 1
    if mode == "native":
        ps = capi.params_struct(par)
        ei = EaInfo()
        mod._check(mod.lib.ea_qp_admm_one_level(mod._need(), C.byref(ps), C.byref(ei)))
        info.status = STATUS_NAMES[ei.status]
        for k in ("inner", "outer", "cumul", "objval", "primres", "dualres", "mismatch", "auglag", "time_overall"):
            setattr(info, k, getattr(ei, k))
        admm_poststep(env, mod, device)
        return
    if mode != "stepwise":
        raise ValueError(f"admm_one_level: unknown mode {mode!r}")
    z = np.zeros(mod.nvar)
    mod.solution.lz = z
    mod.solution.z_curr = z
    mod.solution.z_prev = z
    if par.verbose > 0:
        admm_update_residual(env, mod, device)
    info.status = "IterationLimit"
    norm_rho = float(np.linalg.norm(mod.solution.rho))
    t0 = time.perf_counter()
    while info.outer < par.outer_iterlim:
        admm_increment_outer(env, mod, device)
        admm_increment_reset_inner(env, mod)
        while info.inner < par.inner_iterlim:
            admm_increment_inner(env, mod, device)
            admm_update_x(env, mod, device)
            admm_update_xbar(env, mod, device)
            admm_update_l_single(env, mod, device)
            admm_update_residual(env, mod, device)
            if par.verbose > 0:
                if info.cumul % 50 == 0:
                    print("%8s  %10s  %10s  %10s  %10s  %10s  %10s" % ("Iter", "Objval", "Auglag", "PrimRes", "PrimTol",
                                                                      "DualRes", "DualTol"))
                print("%8d  %10.3e  %10.3e  %10.3e  %10.3e %10.3e  %10.3e" % (
                    info.outer, info.objval, info.auglag, info.mismatch, OUTER_TOL, info.dualres,
                    OUTER_TOL * norm_rho / sqrt_d))
        if info.mismatch <= OUTER_TOL and info.dualres <= OUTER_TOL * norm_rho / sqrt_d:
            info.status = "Solved"
            break
    info.time_overall = time.perf_counter() - t0
    admm_poststep(env, mod, device)


def solve_qpsub(case, Hs, LH_1h, RH_1h, LH_1i, RH_1i, LH_1j, RH_1j, LH_1k, RH_1k, ls, us, pgmax, pgmin, qgmax, qgmin,
                c1, c2, Pd, Qd, initial_beta, *, case_format="matpower", outer_iterlim=20, inner_iterlim=1000,
                rho_pq=400.0, rho_va=40000.0, obj_scale=1.0, scale=1e-4, storage_ratio=0.0, storage_charge_max=1.0,
                use_gpu=False, ka_device=None, use_linelimit=True, use_projection=False, tight_factor=1.0,
                outer_eps=2e-4, gpu_no=0, verbose=1, onelevel=True, mode="native"):
    """``solve_qpsub(case, Hs, ..., Qd, initial_beta; kwargs...) -> (env, mod)`` (solve_qpsub.jl:1-126), same positional
    arguments, keywords, defaults and return."""
    if ka_device is not None:
        raise NotImplementedError("KernelAbstractions devices are out of scope: no KA/ROCm dispatch on this path")
    if not use_gpu:
        raise NotImplementedError("use_gpu=false: the reference's CPU path is not part of this package; pass use_gpu=True")
    env = AdmmEnv(case, rho_pq, rho_va, case_format=case_format, use_gpu=use_gpu, ka_device=ka_device,
                  use_linelimit=use_linelimit, use_projection=use_projection, tight_factor=tight_factor, gpu_no=gpu_no,
                  storage_ratio=storage_ratio, storage_charge_max=storage_charge_max, verbose=verbose)
    mod = ModelQpsub(env)
    f64 = lambda a: np.array(a, dtype=np.float64)             # copy(...) in the reference
    mod.Hs, mod.LH_1h, mod.RH_1h, mod.LH_1i, mod.RH_1i = f64(Hs), f64(LH_1h), f64(RH_1h), f64(LH_1i), f64(RH_1i)
    mod.LH_1j, mod.RH_1j, mod.LH_1k, mod.RH_1k = f64(LH_1j), f64(RH_1j), f64(LH_1k), f64(RH_1k)
    mod.ls, mod.us = f64(ls), f64(us)
    mod.qpsub_pgmax, mod.qpsub_pgmin, mod.qpsub_qgmax, mod.qpsub_qgmin = f64(pgmax), f64(pgmin), f64(qgmax), f64(qgmin)
    mod.qpsub_c1, mod.qpsub_c2, mod.qpsub_Pd, mod.qpsub_Qd = f64(c1), f64(c2), f64(Pd), f64(Qd)
    env.params.scale = scale
    env.params.obj_scale = obj_scale          # stored; the constructor's scaling step has already run (inert, as in the reference)
    env.params.outer_eps = outer_eps
    env.params.outer_iterlim = outer_iterlim
    env.params.inner_iterlim = inner_iterlim
    env.params.shmem_size = 8 * (16 * mod.n + 4 * mod.n ** 2 + 178) + 8 * (4 * mod.n)     # solve_qpsub.jl:110 (informational)
    env.params.initial_beta = initial_beta
    init_solution(mod, mod.solution, env.initial_rho_pq, env.initial_rho_va)
    if onelevel:
        admm_one_level(env, mod, None, mode=mode)
    else:
        import warnings
        warnings.warn("two-level ADMM is not implemented in QPsub")     # solve_qpsub.jl:123
    return env, mod
