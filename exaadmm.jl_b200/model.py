"""``ModelAcopf`` and ``Solution`` — host mirrors of
``/root/reference/src/models/acopf/acopf_model.jl:6-94`` and
``/root/reference/src/utils/environment.jl:177-226`` whose arrays live in HBM
behind the C ABI.

``mod.solution.u_curr`` etc. return host copies (the reference's tests do
``copyto!(host, mod.solution.u_curr)``); assigning to an attribute uploads.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .capi import EaCounters, EaError, FIELDS, dptr
from .environment import AdmmEnv, IterationInformation
from .grid_data import GridData


class Solution:
    """Device-resident ``Solution{T,TD}``; attribute access = D2H copy in the
    reference layout ``[(pg,qg) x ngen | 8 x nline]``."""

    _scalars = ("overall_time", "max_viol_except_line", "max_line_viol_rateA", "cumul_iters", "status")

    def __init__(self, model: "ModelAcopf"):
        object.__setattr__(self, "_m", model)
        object.__setattr__(self, "overall_time", float("inf"))
        object.__setattr__(self, "max_viol_except_line", float("inf"))
        object.__setattr__(self, "max_line_viol_rateA", float("inf"))
        object.__setattr__(self, "cumul_iters", 0)
        object.__setattr__(self, "status", "NotSpecified")

    def __getattr__(self, name):
        if name in FIELDS:
            return self._m.get_vector(name)
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if name in FIELDS:
            self._m.set_vector(name, value)
        elif name in self._scalars:
            object.__setattr__(self, name, value)
        else:
            raise AttributeError(name)


class ModelAcopf:
    """``ModelAcopf{T,TD,TI,TM}(env)`` (acopf_model.jl:41-94) on one B200."""

    def __init__(self, env: AdmmEnv, ramp_ratio: float = 0.02, grid: GridData | None = None):
        if not env.use_gpu:
            raise NotImplementedError(
                "exaadmm_b200 implements the use_gpu=true path only; there is no CPU fallback "
                "(the CPU restatement under oracle/ is test infrastructure).")
        if not env.use_linelimit:
            raise NotImplementedError(
                "use_linelimit=false has no defined behaviour in the reference "
                "(polar_kernel_two_level_alternative is never defined; SURVEY.md F4)")
        self.lib = capi.load_library()
        self.env = env
        self.grid_data = grid if grid is not None else GridData.from_opfdata(env.data, tight_factor=env.tight_factor)
        g = self.grid_data
        self.n = 6                                   # acopf_model.jl:46
        self.nline_padded = g.nline
        self.nvar = 2 * g.ngen + 8 * g.nline         # :55
        self.nvar_padded = self.nvar
        self.gen_start = 1                           # :57 (1-based, as in the reference)
        self.line_start = 2 * g.ngen + 1             # :58
        self.pgmin_curr = g.pgmin.copy()             # :61-64
        self.pgmax_curr = g.pgmax.copy()
        g.ramp_rate = ramp_ratio * g.pgmax           # :66-67
        # env.params.obj_scale is still 1.0 here in solve_acopf -> no cost scaling (:69-73)
        if env.params.obj_scale != 1.0:
            g.c2 = g.c2 * env.params.obj_scale
            g.c1 = g.c1 * env.params.obj_scale
            g.c0 = g.c0 * env.params.obj_scale
        self.nvar_u = self.nvar
        self.nvar_u_padded = self.nvar
        self.nvar_v = 2 * g.ngen + 4 * g.nline + 2 * g.nbus
        self.bus_start = 2 * g.ngen + 4 * g.nline + 1

        gs, keep = capi.make_grid_struct(g)
        h = C.c_void_p()
        rc = self.lib.ea_create(C.byref(gs), int(env.gpu_no), C.byref(h))
        if rc != 0:
            raise EaError(rc, (self.lib.ea_last_error(None) or b"").decode())
        self.h = h
        del keep
        self.solution = Solution(self)
        self.gen_solution = None                     # EmptyGeneratorSolution
        self._check(self.lib.ea_init_solution(self.h, env.initial_rho_pq, env.initial_rho_va))   # :83
        self.info = IterationInformation()

    @classmethod
    def borrowed(cls, env: AdmmEnv, handle, grid: GridData, ramp_ratio: float = 0.02) -> "ModelAcopf":
        """A period of a multi-period model: the handle is owned by the ``ModelMpacopf`` (``ea_mp_period``) and has
        been initialised by it; nothing is created or destroyed here."""
        self = cls.__new__(cls)
        self.lib = capi.load_library()
        self.env = env
        self.grid_data = g = grid
        self.n = 6
        self.nline_padded = g.nline
        self.nvar = self.nvar_padded = self.nvar_u = self.nvar_u_padded = 2 * g.ngen + 8 * g.nline
        self.gen_start = 1
        self.line_start = 2 * g.ngen + 1
        self.pgmin_curr = g.pgmin.copy()
        self.pgmax_curr = g.pgmax.copy()
        g.ramp_rate = ramp_ratio * g.pgmax
        self.nvar_v = 2 * g.ngen + 4 * g.nline + 2 * g.nbus
        self.bus_start = 2 * g.ngen + 4 * g.nline + 1
        self.h = handle
        self._borrowed = True
        self.solution = Solution(self)
        self.gen_solution = None
        self.info = IterationInformation()
        return self

    # -- plumbing ---------------------------------------------------------------
    def _check(self, rc: int):
        if rc != 0:
            raise EaError(rc, (self.lib.ea_last_error(self.h) or b"").decode())

    def close(self):
        if getattr(self, "h", None):
            if not getattr(self, "_borrowed", False):
                self.lib.ea_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def get_vector(self, name: str) -> np.ndarray:
        out = np.empty(self.nvar, dtype=np.float64)
        self._check(self.lib.ea_get_vector(self.h, FIELDS[name], dptr(out), self.nvar))
        return out

    def set_vector(self, name: str, value) -> None:
        a = np.ascontiguousarray(value, dtype=np.float64)
        if a.shape != (self.nvar,):
            raise ValueError(f"{name}: expected shape ({self.nvar},), got {a.shape}")
        self._check(self.lib.ea_set_vector(self.h, FIELDS[name], dptr(a), self.nvar))

    @property
    def membuf(self) -> np.ndarray:
        """31 x nline host copy of ``mod.membuf`` (acopf_model.jl:87-89); row r of the
        reference (1-based) is ``membuf[r-1]``."""
        nl = self.grid_data.nline
        out = np.zeros((31, nl))
        row = np.empty(nl)
        for r in range(1, 32):
            self._check(self.lib.ea_get_membuf(self.h, r, dptr(row), nl))
            out[r - 1] = row
        return out

    def set_membuf_row(self, row_1based: int, values) -> None:
        a = np.ascontiguousarray(values, dtype=np.float64)
        self._check(self.lib.ea_set_membuf(self.h, int(row_1based), dptr(a), a.shape[0]))

    def set_load(self, Pd, Qd) -> None:
        Pd = np.ascontiguousarray(Pd, dtype=np.float64)
        Qd = np.ascontiguousarray(Qd, dtype=np.float64)
        self._check(self.lib.ea_set_load(self.h, dptr(Pd), dptr(Qd), Pd.shape[0]))
        self.grid_data.Pd, self.grid_data.Qd = Pd.copy(), Qd.copy()

    def set_pg_bounds(self, lo, hi) -> None:
        lo = np.ascontiguousarray(lo, dtype=np.float64)
        hi = np.ascontiguousarray(hi, dtype=np.float64)
        self._check(self.lib.ea_set_pg_bounds(self.h, dptr(lo), dptr(hi), lo.shape[0]))
        self.pgmin_curr, self.pgmax_curr = lo.copy(), hi.copy()

    def counters(self) -> dict:
        c = EaCounters()
        self._check(self.lib.ea_get_counters(self.h, C.byref(c)))
        return c.as_dict()

    def reset_counters(self) -> None:
        self._check(self.lib.ea_reset_counters(self.h))

    def set_option(self, name: str, value: float) -> None:
        self._check(self.lib.ea_set_option(self.h, name.encode(), float(value)))
