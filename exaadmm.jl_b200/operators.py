"""The reference's operator API for the ACOPF model on the B200 path.

One Python function per generic function of ``docs/src/dev.md:65-155`` — the
boundary ``admm_two_level`` and the reference's tests drive — each body being a
single call through the C ABI (the Julia twin is one ``ccall`` per method, see
``julia/``). Signatures keep ``(env, mod, device=None)``; ``device`` is accepted
and ignored, as the dispatch tag it is in the reference.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .capi import dptr


def _mp(mod):
    """Dispatch on the model type, as the reference's multiple dispatch does: ``ModelMpacopf`` has its own methods
    (src/models/mpacopf/), implemented in mpacopf.py."""
    if getattr(mod, "is_multiperiod", False):
        from . import mpacopf
        return mpacopf
    if getattr(mod, "is_qpsub", False):         # ModelQpsub methods (src/models/qpsub/), implemented in qpsub.py
        from . import qpsub
        return qpsub
    return None


# -- counters: acopf_admm_increment.jl:1-36 ------------------------------------
def admm_increment_outer(env, mod, device=None):
    if _mp(mod):
        return _mp(mod).admm_increment_outer(env, mod)
    mod.info.outer += 1


def admm_increment_reset_inner(env, mod, device=None):
    if _mp(mod):
        return _mp(mod).admm_increment_reset_inner(env, mod)
    mod.info.inner = 0


def admm_increment_inner(env, mod, device=None):
    if _mp(mod):
        return _mp(mod).admm_increment_inner(env, mod)
    mod.info.inner += 1
    mod.info.cumul += 1


# -- pre/post steps: acopf_admm_prepoststep_gpu.jl -------------------------------
def admm_outer_prestep(env, mod, device=None):
    if _mp(mod):
        return _mp(mod).admm_outer_prestep(env, mod)
    out = C.c_double()
    mod._check(mod.lib.ea_outer_prestep(mod.h, C.byref(out)))
    mod.info.norm_z_prev = out.value


def admm_inner_prestep(env, mod, device=None):
    if _mp(mod):
        return _mp(mod).admm_inner_prestep(env, mod)
    mod._check(mod.lib.ea_inner_prestep(mod.h))


def admm_poststep(env, mod, device=None):
    if env.use_projection:
        # the reference calls pf_projection, which is commented out of the module
        # (ExaAdmm.jl:127-134) -> UndefVarError there; refuse explicitly here.
        raise NotImplementedError("use_projection=true is not available on this path (SURVEY.md §2)")
    if _mp(mod):
        return _mp(mod).admm_poststep(env, mod)
    out = C.c_double()
    mod._check(mod.lib.ea_poststep(mod.h, C.byref(out)))
    mod.info.objval = out.value


# -- x update: acopf_admm_update_x_gpu.jl ----------------------------------------
def acopf_admm_update_x_gen(env, mod, gen_solution=None):
    mod._check(mod.lib.ea_update_x_gen(mod.h))


def acopf_admm_update_x_line(env, mod):
    par = env.params
    mod._check(mod.lib.ea_update_x_line(mod.h, mod.info.inner, par.max_auglag, par.mu_max, par.scale))


def admm_update_x(env, mod, device=None):
    if _mp(mod):
        return _mp(mod).admm_update_x(env, mod)
    acopf_admm_update_x_gen(env, mod, mod.gen_solution)
    acopf_admm_update_x_line(env, mod)


def admm_update_xbar(env, mod, device=None):
    if _mp(mod):
        return _mp(mod).admm_update_xbar(env, mod)
    mod._check(mod.lib.ea_update_xbar(mod.h))


def admm_update_l_single(env, mod, device=None):
    """One-level ADMM only (qpsub_admm_update_l_single_gpu.jl); defined for ``ModelQpsub``."""
    if not getattr(mod, "is_qpsub", False):
        raise TypeError("admm_update_l_single has a method for ModelQpsub only (as in the reference)")
    return _mp(mod).admm_update_l_single(env, mod)


def admm_update_z(env, mod, device=None):
    if _mp(mod):
        return _mp(mod).admm_update_z(env, mod)
    mod._check(mod.lib.ea_update_z(mod.h, env.params.beta))


def admm_update_l(env, mod, device=None):
    if _mp(mod):
        return _mp(mod).admm_update_l(env, mod)
    mod._check(mod.lib.ea_update_l(mod.h, env.params.beta))


def admm_update_lz(env, mod, device=None):
    if _mp(mod):
        return _mp(mod).admm_update_lz(env, mod)
    mod._check(mod.lib.ea_update_lz(mod.h, env.params.beta, env.params.MAX_MULTIPLIER))


def admm_update_residual(env, mod, device=None):
    if _mp(mod):
        return _mp(mod).admm_update_residual(env, mod)
    out = np.zeros(4)
    mod._check(mod.lib.ea_update_residual(mod.h, dptr(out)))
    info = mod.info
    info.primres, info.dualres, info.norm_z_curr, info.mismatch = (float(x) for x in out)


# -- fused fast path (no reference twin: replaces the body of the inner loop) -------
def admm_inner_iteration(env, mod, device=None):
    """inner_prestep + update_x + update_xbar + update_z + update_l + update_residual
    (admm_two_level.jl:36-42) as two fused launches."""
    par = env.params
    out = np.zeros(4)
    mod._check(mod.lib.ea_inner_iteration(mod.h, mod.info.inner, par.beta, par.max_auglag, par.mu_max,
                                          par.scale, dptr(out)))
    info = mod.info
    info.primres, info.dualres, info.norm_z_curr, info.mismatch = (float(x) for x in out)


def admm_run_inner(env, mod, chunk: int = 0):
    """The whole inner ``while`` of one outer iteration on the device."""
    if _mp(mod):
        return _mp(mod).admm_run_inner(env, mod, chunk)
    par = env.params
    out = np.zeros(4)
    done = C.c_int64()
    mod._check(mod.lib.ea_run_inner(mod.h, mod.info.outer, par.beta, par.inner_iterlim, par.max_auglag,
                                    par.mu_max, par.scale, int(chunk), C.byref(done), dptr(out)))
    info = mod.info
    info.inner = int(done.value)
    info.cumul += int(done.value)
    if done.value > 0:
        info.primres, info.dualres, info.norm_z_curr, info.mismatch = (float(x) for x in out)
