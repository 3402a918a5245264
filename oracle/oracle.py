"""ctypes wrapper of the CPU ORACLE (``oracle/acopf_oracle.c``).

TEST INFRASTRUCTURE ONLY. Importable from ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs; the product
package never imports this module. Pinning status: see ``acopf_oracle.h``.

``OracleModel`` offers the reference's operator names (``admm_update_x`` …) so
that the parity tests read like ``test/algorithms/acopf_update_cpu.jl``.
"""
from __future__ import annotations

import ctypes as C
import subprocess
import sys
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(_HERE.parent))
from exaadmm_b200.capi import (EaGrid, EaParams, EaInfo, EaCounters, EaQpsubData, FIELDS, QP_ARRAYS,  # noqa: E402
                               make_grid_struct, make_qpsub_struct, params_struct)

LIB_PATH = _HERE / "_build" / "libacopf_oracle.so"
_pd = C.POINTER(C.c_double)


def build(force: bool = False) -> Path:
    newest = max(f.stat().st_mtime for f in list(_HERE.glob("*.c")) + list(_HERE.glob("*.h")))
    if force or not LIB_PATH.exists() or LIB_PATH.stat().st_mtime < newest:
        subprocess.run(["make", "-C", str(_HERE)] + (["-B"] if force else []), check=True,
                       stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(LIB_PATH))
        H = C.c_void_p
        L.orc_create.argtypes = [C.POINTER(EaGrid), C.POINTER(H)]
        L.orc_create.restype = C.c_int
        L.orc_destroy.argtypes = [H]
        L.orc_set_threads.argtypes = [H, C.c_int]
        L.orc_set_portable_sincos.argtypes = [C.c_int]
        L.orc_set_portable_sincos.restype = None
        L.orc_get_threads.argtypes = [H]
        L.orc_get_threads.restype = C.c_int
        L.orc_init_solution.argtypes = [H, C.c_double, C.c_double]
        L.orc_outer_prestep.argtypes = [H]
        L.orc_outer_prestep.restype = C.c_double
        L.orc_inner_prestep.argtypes = [H]
        L.orc_update_x_gen.argtypes = [H]
        L.orc_update_x_line.argtypes = [H, C.c_int64, C.c_int32, C.c_double, C.c_double]
        L.orc_update_xbar.argtypes = [H]
        L.orc_update_z.argtypes = [H, C.c_double]
        L.orc_update_l.argtypes = [H, C.c_double]
        L.orc_update_lz.argtypes = [H, C.c_double, C.c_double]
        L.orc_update_residual.argtypes = [H, _pd]
        L.orc_poststep.argtypes = [H]
        L.orc_poststep.restype = C.c_double
        L.orc_admm_two_level.argtypes = [H, C.POINTER(EaParams), C.POINTER(EaInfo)]
        L.orc_admm_two_level.restype = C.c_int
        L.orc_nvar.argtypes = [H]
        L.orc_nvar.restype = C.c_int64
        L.orc_vector.argtypes = [H, C.c_int]
        L.orc_vector.restype = _pd
        L.orc_membuf.argtypes = [H]
        L.orc_membuf.restype = _pd
        L.orc_set_load.argtypes = [H, _pd, _pd]
        L.orc_set_pg_bounds.argtypes = [H, _pd, _pd]
        L.orc_get_counters.argtypes = [H, C.POINTER(EaCounters)]
        L.orc_reset_counters.argtypes = [H]
        L.orc_eval_f.argtypes = [_pd, _pd, _pd, C.c_double]
        L.orc_eval_f.restype = C.c_double
        L.orc_eval_f_reforder.argtypes = [_pd, _pd, _pd, C.c_double]
        L.orc_eval_f_reforder.restype = C.c_double
        L.orc_eval_gh.argtypes = [_pd, _pd, _pd, C.c_double, _pd, _pd]
        L.orc_tron_solve.argtypes = [_pd, _pd, _pd, _pd, _pd, C.c_double, C.c_int, C.c_int, C.c_double,
                                     C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.orc_tron_solve.restype = C.c_int
        # multi-period model (mpacopf_oracle.h)
        L.orc_mp_create.argtypes = [C.POINTER(EaGrid), C.c_int32, _pd, _pd, C.c_double, C.POINTER(H)]
        L.orc_mp_create.restype = C.c_int
        L.orc_mp_destroy.argtypes = [H]
        L.orc_mp_set_threads.argtypes = [H, C.c_int]
        L.orc_mp_period.argtypes = [H, C.c_int32]
        L.orc_mp_period.restype = H
        L.orc_mp_ramp_vector.argtypes = [H, C.c_int32, C.c_int]
        L.orc_mp_ramp_vector.restype = _pd
        L.orc_mp_gen_membuf.argtypes = [H, C.c_int32]
        L.orc_mp_gen_membuf.restype = _pd
        L.orc_mp_nvar.argtypes = [H]
        L.orc_mp_nvar.restype = C.c_int64
        L.orc_mp_init_solution.argtypes = [H, C.c_double, C.c_double]
        L.orc_mp_outer_prestep.argtypes = [H]
        L.orc_mp_outer_prestep.restype = C.c_double
        L.orc_mp_inner_prestep.argtypes = [H]
        L.orc_mp_update_x.argtypes = [H, C.c_int64, C.c_int32, C.c_double, C.c_double]
        L.orc_mp_update_xbar.argtypes = [H]
        L.orc_mp_update_z.argtypes = [H, C.c_double]
        L.orc_mp_update_l.argtypes = [H, C.c_double]
        L.orc_mp_update_lz.argtypes = [H, C.c_double, C.c_double]
        L.orc_mp_update_residual.argtypes = [H, _pd]
        L.orc_mp_poststep.argtypes = [H, _pd]
        L.orc_mp_poststep.restype = C.c_double
        L.orc_mp_admm_two_level.argtypes = [H, C.POINTER(EaParams), C.POINTER(EaInfo), _pd]
        L.orc_mp_admm_two_level.restype = C.c_int
        L.orc_gen_ramp_solve.argtypes = [_pd, _pd, _pd, _pd, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                         C.c_int32, C.c_double, C.POINTER(C.c_int32)]
        # one-level ADMM on the SQP sub-problem (qpsub_oracle.h)
        L.orc_qp_create.argtypes = [C.POINTER(EaGrid), C.POINTER(EaQpsubData), C.POINTER(H)]
        L.orc_qp_create.restype = C.c_int
        L.orc_qp_destroy.argtypes = [H]
        L.orc_qp_set_threads.argtypes = [H, C.c_int]
        L.orc_qp_nvar.argtypes = [H]
        L.orc_qp_nvar.restype = C.c_int64
        L.orc_qp_vector.argtypes = [H, C.c_int]
        L.orc_qp_vector.restype = _pd
        L.orc_qp_line_array.argtypes = [H, C.c_int]
        L.orc_qp_line_array.restype = _pd
        L.orc_qp_init_solution.argtypes = [H, C.c_double, C.c_double]
        L.orc_qp_update_x.argtypes = [H, C.c_int64, C.c_int32, C.c_double, C.c_double]
        L.orc_qp_update_xbar.argtypes = [H]
        L.orc_qp_update_l_single.argtypes = [H]
        L.orc_qp_update_residual.argtypes = [H, _pd]
        L.orc_qp_poststep.argtypes = [H, _pd, _pd, _pd, _pd, _pd]
        L.orc_qp_admm_one_level.argtypes = [H, C.POINTER(EaParams), C.POINTER(EaInfo)]
        L.orc_qp_admm_one_level.restype = C.c_int
        L.orc_qp_counters.argtypes = [H, C.POINTER(C.c_int64)]
        L.orc_qp_branch_qp.argtypes = [_pd] * 8 + [C.c_double, _pd, C.c_double, _pd, C.c_double, _pd, C.c_double, _pd,
                                                   C.c_double, _pd, _pd, _pd, _pd]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(_pd)


def set_portable_sincos(on: bool) -> None:
    """Process-wide: the oracle's branch objective uses the portable sin / cos shared with the PARITY build of the CUDA
    library (oracle/portable_sincos.h) instead of libm. For bit-for-bit comparisons with that build only."""
    lib().orc_set_portable_sincos(1 if on else 0)


class OracleModel:
    """CPU-oracle twin of ``ModelAcopf`` + the L3 operator functions."""

    def __init__(self, grid, params, rho_pq: float, rho_va: float):
        self.L = lib()
        self.grid = grid
        self.par = params
        gs, self._keep = make_grid_struct(grid)
        h = C.c_void_p()
        rc = self.L.orc_create(C.byref(gs), C.byref(h))
        if rc != 0:
            raise RuntimeError(f"orc_create failed: {rc}")
        self.h = h
        self.nvar = int(self.L.orc_nvar(h))
        self.nline = grid.nline
        self.L.orc_init_solution(h, rho_pq, rho_va)
        self.inner = self.outer = self.cumul = 0
        self.res = np.zeros(4)

    @classmethod
    def borrowed(cls, handle, grid, params):
        """View of a model owned by someone else (a period of ``OracleMpModel``): never destroyed here."""
        self = cls.__new__(cls)
        self.L = lib()
        self.grid, self.par = grid, params
        self.h = C.c_void_p(handle) if not isinstance(handle, C.c_void_p) else handle
        self._borrowed = True
        self.nvar = int(self.L.orc_nvar(self.h))
        self.nline = grid.nline
        self.inner = self.outer = self.cumul = 0
        self.res = np.zeros(4)
        return self

    def __del__(self):
        if getattr(self, "h", None) and not getattr(self, "_borrowed", False):
            self.L.orc_destroy(self.h)
            self.h = None

    def set_threads(self, n):
        self.L.orc_set_threads(self.h, int(n))

    # -- Solution access (views into the oracle's memory) -------------------
    def vec(self, name: str) -> np.ndarray:
        ptr = self.L.orc_vector(self.h, FIELDS[name])
        return np.ctypeslib.as_array(ptr, shape=(self.nvar,))

    def membuf(self) -> np.ndarray:
        """31 x nline, as in the reference (rows 0-based here)."""
        ptr = self.L.orc_membuf(self.h)
        return np.ctypeslib.as_array(ptr, shape=(self.nline, 31)).T

    # -- operator API (names of docs/src/dev.md:65-155) -----------------------
    def admm_increment_outer(self): self.outer += 1
    def admm_increment_reset_inner(self): self.inner = 0
    def admm_increment_inner(self): self.inner += 1; self.cumul += 1
    def admm_outer_prestep(self): self.norm_z_prev = self.L.orc_outer_prestep(self.h)
    def admm_inner_prestep(self): self.L.orc_inner_prestep(self.h)

    def admm_update_x(self):
        self.L.orc_update_x_gen(self.h)
        self.L.orc_update_x_line(self.h, self.inner, self.par.max_auglag, self.par.mu_max, self.par.scale)

    def admm_update_x_gen(self): self.L.orc_update_x_gen(self.h)

    def admm_update_x_line(self):
        self.L.orc_update_x_line(self.h, self.inner, self.par.max_auglag, self.par.mu_max, self.par.scale)

    def admm_update_xbar(self): self.L.orc_update_xbar(self.h)
    def admm_update_z(self): self.L.orc_update_z(self.h, self.par.beta)
    def admm_update_l(self): self.L.orc_update_l(self.h, self.par.beta)
    def admm_update_lz(self): self.L.orc_update_lz(self.h, self.par.beta, self.par.MAX_MULTIPLIER)

    def admm_update_residual(self):
        self.L.orc_update_residual(self.h, _p(self.res))
        return self.res.copy()

    def admm_poststep(self): return self.L.orc_poststep(self.h)

    def inner_iteration(self):
        """inner_prestep .. update_residual of admm_two_level.jl:36-42."""
        self.admm_increment_inner()
        self.admm_inner_prestep()
        self.admm_update_x()
        self.admm_update_xbar()
        self.admm_update_z()
        self.admm_update_l()
        return self.admm_update_residual()

    def admm_two_level(self) -> EaInfo:
        info = EaInfo()
        p = params_struct(self.par)
        rc = self.L.orc_admm_two_level(self.h, C.byref(p), C.byref(info))
        if rc != 0:
            raise RuntimeError(f"orc_admm_two_level failed: {rc}")
        self.par.beta = info.beta
        return info

    def set_load(self, Pd, Qd):
        Pd = np.ascontiguousarray(Pd, dtype=np.float64); Qd = np.ascontiguousarray(Qd, dtype=np.float64)
        self.L.orc_set_load(self.h, _p(Pd), _p(Qd))

    def counters(self) -> dict:
        c = EaCounters()
        self.L.orc_get_counters(self.h, C.byref(c))
        return c.as_dict()

    def reset_counters(self): self.L.orc_reset_counters(self.h)


# -- unit-level helpers --------------------------------------------------------
def eval_f(x, param, Y, scale, reforder=False):
    x = np.ascontiguousarray(x, dtype=np.float64); param = np.ascontiguousarray(param, dtype=np.float64)
    Y = np.ascontiguousarray(Y, dtype=np.float64)
    fn = lib().orc_eval_f_reforder if reforder else lib().orc_eval_f
    return fn(_p(x), _p(param), _p(Y), scale)


def eval_gh(x, param, Y, scale):
    x = np.ascontiguousarray(x, dtype=np.float64); param = np.ascontiguousarray(param, dtype=np.float64)
    Y = np.ascontiguousarray(Y, dtype=np.float64)
    g = np.zeros(6); H = np.zeros(36)
    lib().orc_eval_gh(_p(x), _p(param), _p(Y), scale, _p(g), _p(H))
    return g, H.reshape(6, 6)


def tron_solve(x0, xl, xu, param, Y, scale, max_feval=500, max_minor=200, gtol=1e-6):
    x = np.array(x0, dtype=np.float64)
    xl = np.ascontiguousarray(xl, dtype=np.float64); xu = np.ascontiguousarray(xu, dtype=np.float64)
    param = np.ascontiguousarray(param, dtype=np.float64); Y = np.ascontiguousarray(Y, dtype=np.float64)
    minor = C.c_int(); nfev = C.c_int()
    st = lib().orc_tron_solve(_p(x), _p(xl), _p(xu), _p(param), _p(Y), scale, max_feval, max_minor, gtol,
                              C.byref(minor), C.byref(nfev))
    return x, st, minor.value, nfev.value


RAMP_FIELDS = {"u_curr": 0, "v_curr": 1, "l_curr": 2, "rho": 3, "rd": 4, "rp": 5, "z_outer": 6, "z_curr": 7,
               "z_prev": 8, "lz": 9, "Ax_plus_By": 10, "s_curr": 11}


class OracleMpModel:
    """CPU-oracle twin of ``ModelMpacopf`` (src/models/mpacopf/) + its operator functions.

    ``Pd``, ``Qd``: arrays (T, nbus) in MW / MVAr. ``models[t]`` are borrowed ``OracleModel`` views.
    """

    def __init__(self, grid, params, rho_pq: float, rho_va: float, Pd, Qd, ramp_ratio: float = 0.02):
        self.L = lib()
        self.grid, self.par = grid, params
        Pd = np.ascontiguousarray(Pd, dtype=np.float64)
        Qd = np.ascontiguousarray(Qd, dtype=np.float64)
        assert Pd.ndim == 2 and Pd.shape == Qd.shape and Pd.shape[1] == grid.nbus
        self.len_horizon = Pd.shape[0]
        gs, self._keep = make_grid_struct(grid)
        h = C.c_void_p()
        rc = self.L.orc_mp_create(C.byref(gs), self.len_horizon, _p(Pd), _p(Qd), float(ramp_ratio), C.byref(h))
        if rc != 0:
            raise RuntimeError(f"orc_mp_create failed: {rc}")
        self.h = h
        self.ngen = grid.ngen
        self.nvar = int(self.L.orc_mp_nvar(h))
        self.models = [OracleModel.borrowed(self.L.orc_mp_period(h, t), grid, params) for t in range(self.len_horizon)]
        self.L.orc_mp_init_solution(h, rho_pq, rho_va)
        self.inner = self.outer = self.cumul = 0
        self.res = np.zeros(4)
        self.err_ramp = 0.0

    def __del__(self):
        if getattr(self, "h", None):
            self.models = []
            self.L.orc_mp_destroy(self.h)
            self.h = None

    def set_threads(self, n): self.L.orc_mp_set_threads(self.h, int(n))

    def ramp(self, t: int, name: str) -> np.ndarray:
        ptr = self.L.orc_mp_ramp_vector(self.h, t, RAMP_FIELDS[name])
        return np.ctypeslib.as_array(ptr, shape=(self.ngen,))

    def gen_membuf(self, t: int) -> np.ndarray:
        ptr = self.L.orc_mp_gen_membuf(self.h, t)
        return np.ctypeslib.as_array(ptr, shape=(self.ngen, 8)).T

    def init_solution(self, rho_pq, rho_va): self.L.orc_mp_init_solution(self.h, rho_pq, rho_va)
    def admm_increment_outer(self): self.outer += 1
    def admm_increment_reset_inner(self): self.inner = 0
    def admm_increment_inner(self): self.inner += 1; self.cumul += 1
    def admm_outer_prestep(self): self.norm_z_prev = self.L.orc_mp_outer_prestep(self.h)
    def admm_inner_prestep(self): self.L.orc_mp_inner_prestep(self.h)

    def admm_update_x(self):
        self.L.orc_mp_update_x(self.h, self.inner, self.par.max_auglag, self.par.mu_max, self.par.scale)

    def admm_update_xbar(self): self.L.orc_mp_update_xbar(self.h)
    def admm_update_z(self): self.L.orc_mp_update_z(self.h, self.par.beta)
    def admm_update_l(self): self.L.orc_mp_update_l(self.h, self.par.beta)
    def admm_update_lz(self): self.L.orc_mp_update_lz(self.h, self.par.beta, self.par.MAX_MULTIPLIER)

    def admm_update_residual(self):
        self.L.orc_mp_update_residual(self.h, _p(self.res))
        return self.res.copy()

    def admm_poststep(self):
        e = np.zeros(1)
        obj = self.L.orc_mp_poststep(self.h, _p(e))
        self.err_ramp = float(e[0])
        return obj

    def admm_two_level(self) -> EaInfo:
        info = EaInfo()
        ps = params_struct(self.par)
        e = np.zeros(1)
        rc = self.L.orc_mp_admm_two_level(self.h, C.byref(ps), C.byref(info), _p(e))
        if rc != 0:
            raise RuntimeError(f"orc_mp_admm_two_level failed: {rc}")
        self.err_ramp = float(e[0])
        self.par.beta = info.beta
        return info


class OracleQpModel:
    """CPU-oracle twin of ``ModelQpsub`` (src/models/qpsub/) + its operator functions and ``admm_one_level``.

    ``data``: object with the fields of ``exaadmm_b200.qpsub.QpsubData`` (reference shapes)."""

    def __init__(self, grid, params, data, rho_pq: float, rho_va: float):
        self.L = lib()
        self.grid, self.par, self.data = grid, params, data
        gs, self._keep = make_grid_struct(grid)
        ds, self._keep2 = make_qpsub_struct(data, grid.nline, grid.ngen, grid.nbus)
        h = C.c_void_p()
        rc = self.L.orc_qp_create(C.byref(gs), C.byref(ds), C.byref(h))
        if rc != 0:
            raise RuntimeError(f"orc_qp_create failed: {rc}")
        self.h = h
        self.nvar = int(self.L.orc_qp_nvar(h))
        self.nline, self.ngen, self.nbus = grid.nline, grid.ngen, grid.nbus
        self.L.orc_qp_init_solution(h, rho_pq, rho_va)
        self.inner = self.outer = self.cumul = 0
        self.res = np.zeros(5)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_qp_destroy(self.h)
            self.h = None

    def set_threads(self, n): self.L.orc_qp_set_threads(self.h, int(n))

    def vec(self, name: str) -> np.ndarray:
        return np.ctypeslib.as_array(self.L.orc_qp_vector(self.h, FIELDS[name]), shape=(self.nvar,))

    def line_array(self, name: str) -> np.ndarray:
        """``sqp_line`` (6 x nline), ``qpsub_membuf`` (5 x nline) or ``lambda`` (4 x nline), as in the reference."""
        which, rows = QP_ARRAYS[name]
        return np.ctypeslib.as_array(self.L.orc_qp_line_array(self.h, which), shape=(self.nline, rows)).T

    def init_solution(self, rho_pq, rho_va): self.L.orc_qp_init_solution(self.h, rho_pq, rho_va)
    def admm_increment_outer(self): self.outer += 1
    def admm_increment_reset_inner(self): self.inner = 0
    def admm_increment_inner(self): self.inner += 1; self.cumul += 1

    def admm_update_x(self):
        self.L.orc_qp_update_x(self.h, self.inner, self.par.max_auglag, self.par.mu_max, self.par.scale)

    def admm_update_xbar(self): self.L.orc_qp_update_xbar(self.h)
    def admm_update_l_single(self): self.L.orc_qp_update_l_single(self.h)

    def admm_update_residual(self):
        self.L.orc_qp_update_residual(self.h, _p(self.res))
        return self.res.copy()

    def admm_poststep(self) -> dict:
        o, a = np.zeros(1), np.zeros(1)
        dw, dt, di = np.zeros(self.nbus), np.zeros(self.nbus), np.zeros(self.ngen + 6 * self.nline)
        self.L.orc_qp_poststep(self.h, _p(o), _p(a), _p(dw), _p(dt), _p(di))
        u = self.vec("u_curr")
        ls = 2 * self.ngen
        return {"objval": float(o[0]), "auglag": float(a[0]), "dw_sol": dw, "dtheta_sol": dt, "dual_infeas": di,
                "dpg_sol": u[0:ls:2].copy(), "dqg_sol": u[1:ls:2].copy(),
                "dline_var": self.line_array("sqp_line").copy(),
                "dline_fl": u[ls:].reshape(self.nline, 8)[:, :4].T.copy(), "lambda": self.line_array("lambda").copy()}

    def admm_one_level(self) -> EaInfo:
        info = EaInfo()
        ps = params_struct(self.par)
        rc = self.L.orc_qp_admm_one_level(self.h, C.byref(ps), C.byref(info))
        if rc != 0:
            raise RuntimeError(f"orc_qp_admm_one_level failed: {rc}")
        return info

    def counters(self) -> dict:
        c = (C.c_int64 * 4)()
        self.L.orc_qp_counters(self.h, c)
        return dict(zip(("line_calls", "auglag_iters", "tron_evals", "max_auglag_one_call"), map(int, c)))
