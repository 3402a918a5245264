"""``Parameters``, ``AdmmEnv``, ``IterationInformation``, ``ComponentInformation``:
the host-side structs behind the ``(env, mod)`` return of ``solve_acopf``.

Field names and defaults follow ``/root/reference/src/utils/environment.jl``
(``Parameters`` ``:6-76``, ``AdmmEnv`` ``:85-158``, ``ComponentInformation``
``:277-297``, ``IterationInformation`` ``:323-358``). Only what the single-period
two-level path reads is kept; the unused tuning fields of the reference
(``rho_max`` … ``Kf_mean``) are carried as plain attributes for completeness.
"""
from __future__ import annotations

from dataclasses import dataclass, field

from .matpower import OPFData, parse_matpower


@dataclass
class Parameters:
    mu_max: float = 1e8
    max_auglag: int = 50
    ABSTOL: float = 1e-6
    RELTOL: float = 1e-5
    rho_max: float = 1e6
    rho_min_pq: float = 5.0
    rho_min_w: float = 5.0
    eps_rp: float = 1e-4
    eps_rp_min: float = 1e-5
    rt_inc: float = 2.0
    rt_dec: float = 2.0
    eta: float = 0.99
    verbose: int = 1
    shift_lines: int = 0
    initial_beta: float = 1e3
    beta: float = 1e3
    inc_c: float = 6.0
    theta: float = 0.8
    outer_eps: float = 2e-4
    shmem_size: int = 0
    Kf: int = 100
    Kf_mean: int = 10
    MAX_MULTIPLIER: float = 1e12
    DUAL_TOL: float = 1e-8
    outer_iterlim: int = 20
    inner_iterlim: int = 1000
    scale: float = 1e-4
    obj_scale: float = 1.0


@dataclass
class ComponentInformation:
    err_pg: float = 0.0
    err_qg: float = 0.0
    err_vm: float = 0.0
    err_real: float = 0.0
    err_reactive: float = 0.0
    err_rateA: float = 0.0
    err_ramp: float = 0.0
    num_rateA_viols: int = 0
    time_generators: float = 0.0
    time_branches: float = 0.0
    time_buses: float = 0.0


@dataclass
class IterationInformation:
    status: str = "NotSpecified"
    inner: int = 0
    outer: int = 0
    cumul: int = 0
    objval: float = 0.0
    primres: float = 0.0
    dualres: float = 0.0
    mismatch: float = 0.0
    auglag: float = 0.0
    eps_pri: float = 0.0
    norm_z_curr: float = 0.0
    norm_z_prev: float = 0.0
    time_x_update: float = 0.0
    time_xbar_update: float = 0.0
    time_z_update: float = 0.0
    time_l_update: float = 0.0
    time_lz_update: float = 0.0
    time_projection: float = 0.0
    time_overall: float = 0.0
    user: ComponentInformation = field(default_factory=ComponentInformation)

    def fill(self, val=0):
        """``Base.fill!(info, val)`` (environment.jl:360-381); status untouched."""
        for k in ("inner", "outer", "cumul"):
            setattr(self, k, int(val))
        for k in ("objval", "primres", "dualres", "mismatch", "auglag", "eps_pri", "norm_z_curr",
                  "norm_z_prev", "time_x_update", "time_xbar_update", "time_z_update", "time_l_update",
                  "time_lz_update", "time_projection", "time_overall"):
            setattr(self, k, float(val))
        self.user = ComponentInformation()


class AdmmEnv:
    """``AdmmEnv{T,TD,TI,TM}(case, rho_pq, rho_va; ...)`` (environment.jl:104-158).

    ``rho_pq`` / ``rho_va`` must be floats, as in the reference where they are
    typed ``::Float64`` and an integer literal raises a ``MethodError``.
    """

    def __init__(self, case, rho_pq: float, rho_va: float, *, case_format: str = "matpower",
                 use_gpu: bool = False, ka_device=None, use_linelimit: bool = True,
                 use_mpi: bool = False, use_projection: bool = False, gpu_no: int = 0,
                 verbose: int = 1, tight_factor: float = 1.0, droop: float = 0.04,
                 storage_ratio: float = 0.0, storage_charge_max: float = 1.0,
                 horizon_length: int = 1, load_prefix: str = ""):
        if not isinstance(rho_pq, float) or not isinstance(rho_va, float):
            raise TypeError("rho_pq and rho_va must be Float64 (environment.jl:109,154)")
        if case_format.lower() not in ("matpower", "pglib"):
            raise ValueError(f"unsupported case_format {case_format!r}")
        if isinstance(case, OPFData):
            self.data, self.case = case, case.case
        else:
            self.case = str(case)
            self.data = parse_matpower(case, verbose=verbose)
        self.storage_ratio = storage_ratio
        self.droop = droop
        self.initial_rho_pq = rho_pq
        self.initial_rho_va = rho_va
        self.tight_factor = tight_factor
        self.horizon_length = horizon_length
        self.use_gpu = use_gpu
        self.ka_device = ka_device
        self.use_linelimit = use_linelimit
        self.use_mpi = use_mpi
        self.use_projection = use_projection
        self.load_specified = False
        self.gpu_no = gpu_no
        self.params = Parameters()
        self.params.verbose = verbose
        self.load = None
        if load_prefix:
            from .rolling import get_load
            self.load = get_load(load_prefix)
            self.load_specified = True
