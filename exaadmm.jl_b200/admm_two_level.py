"""``admm_two_level(env, mod, device)`` — host control flow of the two-level ADMM.

``mode="stepwise"`` restates ``/root/reference/src/algorithms/admm_two_level.jl:1-88``
line by line on top of the operator API (12 ABI calls per inner iteration, each
synchronous, like the reference). ``mode="fused"`` (default) keeps the outer loop
here and runs each inner ``while`` on the device (``ea_run_inner``);
``mode="native"`` hands the whole algorithm to ``ea_admm_two_level``. All three
produce the same iterates.
"""
from __future__ import annotations

import ctypes as C
import math
import time

from . import operators as ops
from .capi import EaInfo, STATUS_NAMES, params_struct

_HDR = ("%8s  %8s  %10s  %10s  %10s  %10s  %10s  %10s  %10s  %10s  %10s" %
        ("Outer", "Inner", "Objval", "AugLag", "PrimRes", "EpsPrimRes", "DualRes", "||z||", "||Ax+By||",
         "OuterTol", "Beta"))


def _row(info, outer_tol, beta):
    return ("%8d  %8d  %10.3e  %10.3e  %10.3e  %10.3e  %10.3e  %10.3e  %10.3e  %10.3e  %10.3e" %
            (info.outer, info.inner, info.objval, info.auglag, info.primres, info.eps_pri, info.dualres,
             info.norm_z_curr, info.mismatch, outer_tol, beta))


def print_statistics(env, mod):
    """``/root/reference/src/utils/print_statistics.jl:1-20``."""
    info = mod.info
    print(" ** Statistics")
    print("Objective value  . . . . . . . . . %12.6e" % info.objval)
    print("Residual . . . . . . . . . . . . . %12.6e" % info.mismatch)
    print("Outer iterations . . . . . . . . . %12d" % info.outer)
    print("Cumulative iterations  . . . . . . %12d" % info.cumul)
    print("Time per iteration . . . . . . . . %12.3f (secs/iter)" % (info.time_overall / max(info.cumul, 1)))
    print("Overall time . . . . . . . . . . . %12.3f (secs)" % (info.time_overall + info.time_projection))
    print("Projection time  . . . . . . . . . %12.3f (secs)" % info.time_projection)
    print("Generator time . . . . . . . . . . %12.3f (secs)" % info.user.time_generators)
    print("Branch time. . . . . . . . . . . . %12.3f (secs)" % info.user.time_branches)
    print("Bus time . . . . . . . . . . . . . %12.3f (secs)" % info.user.time_buses)
    print("G+Br+B time. . . . . . . . . . . . %12.3f (secs)" %
          (info.user.time_generators + info.user.time_branches + info.user.time_buses))


def _native(env, mod):
    par = env.params
    p = params_struct(par)
    out = EaInfo()
    mod._check(mod.lib.ea_admm_two_level(mod.h, C.byref(p), C.byref(out)))
    info = mod.info
    info.status = STATUS_NAMES[out.status]
    for k in ("inner", "outer", "cumul", "objval", "primres", "dualres", "mismatch", "auglag", "eps_pri",
              "norm_z_curr", "norm_z_prev", "time_x_update", "time_xbar_update", "time_z_update",
              "time_l_update", "time_lz_update", "time_overall"):
        setattr(info, k, getattr(out, k))
    info.user.time_generators = out.time_generators
    info.user.time_branches = out.time_branches
    info.user.time_buses = out.time_buses
    par.beta = out.beta


def admm_two_level(env, mod, device=None, mode: str = "fused", talk: bool = True):
    par, info = env.params, mod.info
    if mode == "native":
        if getattr(mod, "is_multiperiod", False):
            from .mpacopf import admm_two_level_native
            admm_two_level_native(env, mod)
        else:
            _native(env, mod)
        if par.verbose > 0 and talk:          # talk: rank 0 of a partitioned solve (every rank passes the same verbose)
            print_statistics(env, mod)
        return

    sqrt_d = math.sqrt(mod.nvar)
    outer_tol = sqrt_d * par.outer_eps

    info.fill(0)
    info.mismatch = math.inf
    info.norm_z_prev = info.norm_z_curr = math.inf
    par.beta = par.initial_beta

    if par.verbose > 0:
        ops.admm_update_residual(env, mod, device)
        print(_HDR)
        print(_row(info, outer_tol, par.beta))

    info.status = "IterationLimit"
    stepwise = (mode == "stepwise") or par.verbose > 0
    t0 = time.perf_counter()
    while info.outer < par.outer_iterlim:
        ops.admm_increment_outer(env, mod, device)
        ops.admm_outer_prestep(env, mod, device)
        ops.admm_increment_reset_inner(env, mod, device)
        if stepwise:
            while info.inner < par.inner_iterlim:
                ops.admm_increment_inner(env, mod, device)
                if mode == "stepwise" or getattr(mod, "is_multiperiod", False):
                    ops.admm_inner_prestep(env, mod, device)
                    ops.admm_update_x(env, mod, device)
                    ops.admm_update_xbar(env, mod, device)
                    ops.admm_update_z(env, mod, device)
                    ops.admm_update_l(env, mod, device)
                    ops.admm_update_residual(env, mod, device)
                else:
                    ops.admm_inner_iteration(env, mod, device)
                info.eps_pri = sqrt_d / (2500 * info.outer)
                if par.verbose > 0:
                    if info.cumul % 50 == 0:
                        print(_HDR)
                    print(_row(info, outer_tol, par.beta))
                if info.primres <= info.eps_pri:
                    break
        else:
            ops.admm_run_inner(env, mod)
            info.eps_pri = sqrt_d / (2500 * info.outer)

        if info.mismatch <= outer_tol:
            info.status = "Solved"
            break

        ops.admm_update_lz(env, mod, device)
        if info.norm_z_curr > par.theta * info.norm_z_prev:
            par.beta = min(par.inc_c * par.beta, 1e24)

    info.time_overall = time.perf_counter() - t0
    ops.admm_poststep(env, mod, device)
    if par.verbose > 0:
        print_statistics(env, mod)
