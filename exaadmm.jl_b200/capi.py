"""ctypes binding of the C ABI declared in ``include/exaadmm_b200.h``.

This is the Python twin of the Julia ``ccall`` glue in ``julia/`` (the
reference's host language is Julia, which is not available in the build image;
see INTEGRATION.md). The shared library is built in-tree by
``__graft_entry__.build()`` / ``csrc/Makefile``. There is no fallback: if the
library is missing, loading raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

from .grid_data import GridData

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "csrc" / "_build" / "libexaadmm_b200.so"

EA_OK = 0
EA_ERR_ARG, EA_ERR_CUDA, EA_ERR_NCCL, EA_ERR_STATE, EA_ERR_ALLOC = -1, -2, -3, -4, -5
STATUS_NAMES = {0: "NotSpecified", 1: "IterationLimit", 2: "Solved"}

FIELDS = {
    "u_curr": 0, "v_curr": 1, "l_curr": 2, "rho": 3, "z_curr": 4, "z_prev": 5, "lz": 6,
    "rp": 7, "rd": 8, "Ax_plus_By": 9,
    "u_prev": 10, "v_prev": 11, "l_prev": 12, "rp_prev": 13, "z_outer": 14,
}

_pd = C.POINTER(C.c_double)
_pi = C.POINTER(C.c_int64)

_GRID_DOUBLE = ("pgmin", "pgmax", "qgmin", "qgmax", "c2", "c1", "c0", "YshR", "YshI",
                "YffR", "YffI", "YftR", "YftI", "YttR", "YttI", "YtfR", "YtfI",
                "FrVmBound", "ToVmBound", "FrVaBound", "ToVaBound", "rateA")
_GRID_INT_A = ("FrStart", "ToStart", "GenStart", "FrIdx", "ToIdx", "GenIdx")
_GRID_DOUBLE_B = ("Pd", "Qd", "Vmin", "Vmax")


class EaGrid(C.Structure):
    _fields_ = ([("ngen", C.c_int64), ("nline", C.c_int64), ("nbus", C.c_int64), ("baseMVA", C.c_double)]
                + [(n, _pd) for n in _GRID_DOUBLE]
                + [(n, _pi) for n in _GRID_INT_A]
                + [(n, _pd) for n in _GRID_DOUBLE_B]
                + [("brBusIdx", _pi)])


class EaParams(C.Structure):
    _fields_ = [("mu_max", C.c_double), ("max_auglag", C.c_int32), ("verbose", C.c_int32),
                ("initial_beta", C.c_double), ("inc_c", C.c_double), ("theta", C.c_double),
                ("outer_eps", C.c_double), ("MAX_MULTIPLIER", C.c_double), ("scale", C.c_double),
                ("obj_scale", C.c_double), ("outer_iterlim", C.c_int64), ("inner_iterlim", C.c_int64)]


class EaInfo(C.Structure):
    _fields_ = [("status", C.c_int32), ("_pad", C.c_int32),
                ("inner", C.c_int64), ("outer", C.c_int64), ("cumul", C.c_int64),
                ("objval", C.c_double), ("primres", C.c_double), ("dualres", C.c_double),
                ("mismatch", C.c_double), ("auglag", C.c_double), ("eps_pri", C.c_double),
                ("norm_z_curr", C.c_double), ("norm_z_prev", C.c_double), ("beta", C.c_double),
                ("time_x_update", C.c_double), ("time_xbar_update", C.c_double),
                ("time_z_update", C.c_double), ("time_l_update", C.c_double),
                ("time_lz_update", C.c_double), ("time_projection", C.c_double),
                ("time_overall", C.c_double),
                ("time_generators", C.c_double), ("time_branches", C.c_double), ("time_buses", C.c_double)]


class EaCounters(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("line_calls", "auglag_iters", "tron_evals", "cg_iters",
                                           "chol_shifts", "rejected_steps", "max_auglag_hits",
                                           "max_evals_lane")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


_QP_FIELDS = ("Hs", "LH_1h", "RH_1h", "LH_1i", "RH_1i", "LH_1j", "RH_1j", "LH_1k", "RH_1k", "ls", "us", "line_res",
              "pgmax", "pgmin", "qgmax", "qgmin", "c1", "c2", "Pd", "Qd")
QP_ARRAYS = {"sqp_line": (0, 6), "qpsub_membuf": (1, 5), "lambda": (2, 4)}      # enum ea_qp_array: (id, rows)


class EaQpsubData(C.Structure):
    """``ea_qpsub_data_t``: the fields the SQP driver fills in on ``ModelQpsub`` (qpsub_model.jl:63-92)."""
    _fields_ = [(n, _pd) for n in _QP_FIELDS]


def make_qpsub_struct(d, nline: int, ngen: int, nbus: int):
    """``d``: any object with the ``_QP_FIELDS`` attributes in the reference's shapes (Hs (6 nline, 6), LH_1h (nline, 4),
    ls (nline, 6), line_res (4, nline) or None, ...). Returns (struct, keepalive)."""
    shapes = {"Hs": (6 * nline, 6), "LH_1h": (nline, 4), "LH_1i": (nline, 4), "LH_1j": (nline, 2), "LH_1k": (nline, 2),
              "RH_1h": (nline,), "RH_1i": (nline,), "RH_1j": (nline,), "RH_1k": (nline,), "ls": (nline, 6),
              "us": (nline, 6), "pgmax": (ngen,), "pgmin": (ngen,), "qgmax": (ngen,), "qgmin": (ngen,), "c1": (ngen,),
              "c2": (ngen,), "Pd": (nbus,), "Qd": (nbus,)}
    s, keep = EaQpsubData(), []
    for n in _QP_FIELDS:
        a = getattr(d, n, None)
        if n == "line_res":
            if a is None:
                continue                                   # NULL = zeros
            a = np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(4, nline).T)   # -> nline x 4
        else:
            a = np.ascontiguousarray(a, dtype=np.float64)
            if a.shape != shapes[n]:
                raise ValueError(f"qpsub field {n}: shape {a.shape}, expected {shapes[n]}")
        keep.append(a)
        setattr(s, n, a.ctypes.data_as(_pd))
    return s, keep


def make_grid_struct(g: GridData):
    """Build an ``ea_grid_t`` whose pointers borrow the numpy arrays of ``g``.
    Returns (struct, keepalive)."""
    s = EaGrid()
    keep = []
    s.ngen, s.nline, s.nbus, s.baseMVA = g.ngen, g.nline, g.nbus, float(g.baseMVA)
    for n in _GRID_DOUBLE + _GRID_DOUBLE_B:
        a = np.ascontiguousarray(getattr(g, n), dtype=np.float64)
        keep.append(a)
        setattr(s, n, a.ctypes.data_as(_pd))
    for n in _GRID_INT_A + ("brBusIdx",):
        a = np.ascontiguousarray(getattr(g, n), dtype=np.int64)
        keep.append(a)
        setattr(s, n, a.ctypes.data_as(_pi))
    return s, keep


def params_struct(par) -> EaParams:
    p = EaParams()
    for n, _ in EaParams._fields_:
        setattr(p, n, getattr(par, n))
    return p


class LibraryMissing(RuntimeError):
    pass


_lib = None

# name -> (restype, argtypes); must list every symbol include/exaadmm_b200.h declares
_H = C.c_void_p
RAMP_FIELDS = {"u_curr": 0, "v_curr": 1, "l_curr": 2, "rho": 3, "rd": 4, "rp": 5, "z_outer": 6, "z_curr": 7,
               "z_prev": 8, "lz": 9, "Ax_plus_By": 10, "s_curr": 11}      # enum ea_ramp_field

SIGNATURES = {
    "ea_abi_version": (C.c_int, []),
    "ea_last_error": (C.c_char_p, [_H]),
    "ea_device_count": (C.c_int, []),
    "ea_create": (C.c_int, [C.POINTER(EaGrid), C.c_int, C.POINTER(_H)]),
    "ea_destroy": (None, [_H]),
    "ea_init_solution": (C.c_int, [_H, C.c_double, C.c_double]),
    "ea_outer_prestep": (C.c_int, [_H, _pd]),
    "ea_inner_prestep": (C.c_int, [_H]),
    "ea_update_x_gen": (C.c_int, [_H]),
    "ea_update_x_line": (C.c_int, [_H, C.c_int64, C.c_int32, C.c_double, C.c_double]),
    "ea_update_x": (C.c_int, [_H, C.c_int64, C.c_int32, C.c_double, C.c_double]),
    "ea_update_xbar": (C.c_int, [_H]),
    "ea_update_z": (C.c_int, [_H, C.c_double]),
    "ea_update_l": (C.c_int, [_H, C.c_double]),
    "ea_update_residual": (C.c_int, [_H, _pd]),
    "ea_update_lz": (C.c_int, [_H, C.c_double, C.c_double]),
    "ea_poststep": (C.c_int, [_H, _pd]),
    "ea_inner_iteration": (C.c_int, [_H, C.c_int64, C.c_double, C.c_int32, C.c_double, C.c_double, _pd]),
    "ea_run_inner": (C.c_int, [_H, C.c_int64, C.c_double, C.c_int64, C.c_int32, C.c_double,
                               C.c_double, C.c_int32, _pi, _pd]),
    "ea_run_inner_from": (C.c_int, [_H, C.c_int64, C.c_double, C.c_int64, C.c_int64, C.c_int32, C.c_double,
                                    C.c_double, C.c_int32, _pi, _pd]),
    "ea_admm_two_level": (C.c_int, [_H, C.POINTER(EaParams), C.POINTER(EaInfo)]),
    "ea_nvar": (C.c_int64, [_H]),
    "ea_get_vector": (C.c_int, [_H, C.c_int, _pd, C.c_int64]),
    "ea_set_vector": (C.c_int, [_H, C.c_int, _pd, C.c_int64]),
    "ea_get_membuf": (C.c_int, [_H, C.c_int, _pd, C.c_int64]),
    "ea_set_membuf": (C.c_int, [_H, C.c_int, _pd, C.c_int64]),
    "ea_set_load": (C.c_int, [_H, _pd, _pd, C.c_int64]),
    "ea_set_pg_bounds": (C.c_int, [_H, _pd, _pd, C.c_int64]),
    "ea_set_partition": (C.c_int, [_H, C.c_int32, C.c_int32, C.c_int64, C.c_int64, _pi, _pi, C.c_int64, _pi, _pi, _pi, _pi,
                                   C.c_int64, C.c_int64]),
    "ea_nccl_unique_id": (C.c_int, [C.c_char_p, C.c_char_p]),
    "ea_comm_init": (C.c_int, [_H, C.c_char_p, C.c_char_p]),
    "ea_peer_export": (C.c_int, [_H, C.c_char_p]),
    "ea_peer_import": (C.c_int, [_H, C.c_char_p]),
    "ea_part_begin": (C.c_int, [_H, C.c_int64, C.c_double, C.c_int32, C.c_double, C.c_double]),
    "ea_part_get_message": (C.c_int, [_H, _pd, C.c_int64]),
    "ea_part_put_gathered": (C.c_int, [_H, _pd, C.c_int64]),
    "ea_part_end": (C.c_int, [_H, _pd]),
    "ea_get_counters": (C.c_int, [_H, C.POINTER(EaCounters)]),
    "ea_reset_counters": (C.c_int, [_H]),
    "ea_set_option": (C.c_int, [_H, C.c_char_p, C.c_double]),
    "ea_get_kernel_times": (C.c_int, [_H, _pd]),
    # multi-period model (ModelMpacopf)
    "ea_mp_last_error": (C.c_char_p, [_H]),
    "ea_mp_create": (C.c_int, [C.POINTER(EaGrid), C.c_int, C.c_int32, _pd, _pd, C.c_double, C.c_double, C.c_double,
                               C.POINTER(_H)]),
    "ea_mp_destroy": (None, [_H]),
    "ea_mp_len_horizon": (C.c_int32, [_H]),
    "ea_mp_nvar": (C.c_int64, [_H]),
    "ea_mp_period": (_H, [_H, C.c_int32]),
    "ea_mp_init_solution": (C.c_int, [_H, C.c_double, C.c_double]),
    "ea_mp_get_ramp_vector": (C.c_int, [_H, C.c_int32, C.c_int, _pd, C.c_int64]),
    "ea_mp_set_ramp_vector": (C.c_int, [_H, C.c_int32, C.c_int, _pd, C.c_int64]),
    "ea_mp_get_gen_membuf": (C.c_int, [_H, C.c_int32, C.c_int, _pd, C.c_int64]),
    "ea_mp_outer_prestep": (C.c_int, [_H, _pd]),
    "ea_mp_inner_prestep": (C.c_int, [_H]),
    "ea_mp_update_x": (C.c_int, [_H, C.c_int64, C.c_int32, C.c_double, C.c_double]),
    "ea_mp_update_xbar": (C.c_int, [_H]),
    "ea_mp_update_z": (C.c_int, [_H, C.c_double]),
    "ea_mp_update_l": (C.c_int, [_H, C.c_double]),
    "ea_mp_update_lz": (C.c_int, [_H, C.c_double, C.c_double]),
    "ea_mp_update_residual": (C.c_int, [_H, _pd]),
    "ea_mp_poststep": (C.c_int, [_H, _pd, _pd]),
    "ea_mp_run_inner": (C.c_int, [_H, C.c_int64, C.c_double, C.c_int64, C.c_int32, C.c_double, C.c_double, C.c_int32,
                                  C.POINTER(C.c_int64), _pd]),
    "ea_mp_admm_two_level": (C.c_int, [_H, C.POINTER(EaParams), C.POINTER(EaInfo), _pd]),
    "ea_mp_get_kernel_times": (C.c_int, [_H, _pd]),
    # one-level ADMM on the SQP sub-problem (ModelQpsub)
    "ea_qp_last_error": (C.c_char_p, [_H]),
    "ea_qp_create": (C.c_int, [C.POINTER(EaGrid), C.POINTER(EaQpsubData), C.c_int, C.POINTER(_H)]),
    "ea_qp_destroy": (None, [_H]),
    "ea_qp_nvar": (C.c_int64, [_H]),
    "ea_qp_init_solution": (C.c_int, [_H, C.c_double, C.c_double]),
    "ea_qp_get_vector": (C.c_int, [_H, C.c_int, _pd, C.c_int64]),
    "ea_qp_set_vector": (C.c_int, [_H, C.c_int, _pd, C.c_int64]),
    "ea_qp_get_line_array": (C.c_int, [_H, C.c_int, _pd, C.c_int64]),
    "ea_qp_set_line_array": (C.c_int, [_H, C.c_int, _pd, C.c_int64]),
    "ea_qp_update_x": (C.c_int, [_H, C.c_int64, C.c_int32, C.c_double, C.c_double]),
    "ea_qp_update_xbar": (C.c_int, [_H]),
    "ea_qp_update_l_single": (C.c_int, [_H]),
    "ea_qp_update_residual": (C.c_int, [_H, _pd]),
    "ea_qp_poststep": (C.c_int, [_H, _pd, _pd, _pd, _pd, _pd]),
    "ea_qp_admm_one_level": (C.c_int, [_H, C.POINTER(EaParams), C.POINTER(EaInfo)]),
    "ea_qp_set_option": (C.c_int, [_H, C.c_char_p, C.c_double]),
    "ea_qp_get_counters": (C.c_int, [_H, C.POINTER(C.c_int64)]),
    "ea_qp_get_kernel_times": (C.c_int, [_H, _pd]),
    # batch of independent load scenarios
    "ea_batch_last_error": (C.c_char_p, [_H]),
    "ea_batch_create": (C.c_int, [C.POINTER(EaGrid), C.c_int, C.c_int32, _pd, _pd, C.POINTER(_H)]),
    "ea_batch_destroy": (None, [_H]),
    "ea_batch_size": (C.c_int32, [_H]),
    "ea_batch_scenario": (_H, [_H, C.c_int32]),
    "ea_batch_init_solution": (C.c_int, [_H, C.c_double, C.c_double]),
    "ea_batch_admm_two_level": (C.c_int, [_H, C.POINTER(EaParams), C.POINTER(EaInfo)]),
    "ea_batch_set_option": (C.c_int, [_H, C.c_char_p, C.c_double]),
    "ea_batch_get_times": (C.c_int, [_H, _pd]),
    "ea_diag_fp64_peak": (C.c_int, [C.c_int, _pd]),
    "ea_diag_branch_eval": (C.c_int, [C.c_int, C.c_int64, _pd, _pd, _pd, C.c_double, _pd, _pd, _pd]),
    "ea_diag_branch_solve": (C.c_int, [C.c_int, C.c_int, C.c_int64, _pd, C.c_int32, C.c_double, C.c_double, _pd,
                                       C.POINTER(C.c_int32), C.POINTER(C.c_int64), _pd]),
}


def load_library(path: os.PathLike | None = None):
    """dlopen the in-tree CUDA library and declare every signature."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = Path(path) if path else Path(os.environ.get("EXAADMM_B200_LIB", LIB_PATH))
    if not p.exists():
        raise LibraryMissing(
            f"{p} not found: the CUDA library is not built. Run `python -c 'import __graft_entry__ as g; "
            f"g.build()'` (or `make -C {_PKG / 'csrc'}`). There is no CPU fallback on this path.")
    lib = C.CDLL(str(p))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library drift apart
        fn.restype = res
        fn.argtypes = args
    if lib.ea_abi_version() != 1:
        raise RuntimeError(f"ABI version mismatch: library {lib.ea_abi_version()}, binding 1")
    if path is None:
        _lib = lib
    return lib


class EaError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"exaadmm_b200 error {code}: {msg}")
        self.code = code


def dptr(a: np.ndarray):
    return a.ctypes.data_as(_pd)
