"""``solve_acopf(case; kwargs...) -> (env, mod)`` — the user entry point.

Same keyword set, defaults and return as
``/root/reference/src/interface/solve_acopf.jl:1-46``. Only the ``use_gpu=true``,
``ka_device=nothing`` combination (reference case 3, ``:11``) is implemented; the
others raise, because this package has no CPU and no KernelAbstractions path.
"""
from __future__ import annotations

from .admm_two_level import admm_two_level
from .environment import AdmmEnv
from .model import ModelAcopf


def solve_acopf(case, *, case_format="matpower", outer_iterlim=20, inner_iterlim=1000,
                rho_pq=400.0, rho_va=40000.0, obj_scale=1.0, scale=1e-4, storage_ratio=0.0,
                storage_charge_max=1.0, use_gpu=False, ka_device=None, use_linelimit=True,
                use_projection=False, tight_factor=1.0, outer_eps=2e-4, gpu_no=0, verbose=1,
                mode="fused"):
    if ka_device is not None:
        raise NotImplementedError("KernelAbstractions devices are out of scope: no KA/ROCm dispatch on this path")
    if not use_gpu:
        raise NotImplementedError(
            "use_gpu=false: the reference's CPU path is not part of this package (no CPU fallback); "
            "pass use_gpu=True")

    env = AdmmEnv(case, rho_pq, rho_va, case_format=case_format, use_gpu=use_gpu, ka_device=ka_device,
                  use_linelimit=use_linelimit, use_projection=use_projection, tight_factor=tight_factor,
                  gpu_no=gpu_no, storage_ratio=storage_ratio, storage_charge_max=storage_charge_max,
                  verbose=verbose)
    mod = ModelAcopf(env)

    env.params.scale = scale
    env.params.obj_scale = obj_scale          # stored, never read again on this path
    env.params.outer_eps = outer_eps
    env.params.outer_iterlim = outer_iterlim
    env.params.inner_iterlim = inner_iterlim
    env.params.shmem_size = 8 * (14 * mod.n + 3 * mod.n ** 2) + 8 * (4 * mod.n)   # solve_acopf.jl:41 (informational)

    admm_two_level(env, mod, None, mode=mode)
    return env, mod
