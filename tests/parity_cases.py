"""CUDA path vs the CPU oracle on the BASELINE.json configurations (synthetic stand-ins of the named sizes): whole solves
and per-iterate lock-step runs. Used in-process by tests/test_gpu_baseline_configs.py for the fast build and as a script
(`python tests/parity_cases.py <case> ...`, JSON on stdout) in a child process whose EXAADMM_B200_LIB points at the
PARITY build of the library (csrc/Makefile: no fused multiply-add, IEEE division / square root, the objective in the
oracle's operation order).

The analogue in the reference is test/algorithms/acopf_update_gpu.jl:26-194: the GPU path must reproduce the CPU path's
iterates (atol 1e-6 there) and its status / outer / cumulative counts / objective."""
import json
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

# name: (named case, rho_pq, rho_va, scale)   BASELINE.json configs 2-5 (+ a rho at which the 1354-like grid converges)
CONFIGS = {
    "case1354pegase": ("case1354pegase", 1e1, 1e3, 1e-4),
    "case1354pegase_conv": ("case1354pegase", 4e2, 4e4, 1e-4),
    "case2869pegase": ("case2869pegase", 1e1, 1e3, 1e-4),
    "case13659pegase": ("case13659pegase", 5e1, 5e3, 1e-4),
    "ACTIVSg70k": ("ACTIVSg70k", 3e4, 3e5, 1e-5),
}


def _models(config, **params):
    from exaadmm_b200.environment import AdmmEnv, Parameters
    from exaadmm_b200.model import ModelAcopf
    from exaadmm_b200.synthetic import named_case
    from oracle.oracle import OracleModel
    name, rho_pq, rho_va, scale = CONFIGS[config]
    env = AdmmEnv(named_case(name), rho_pq, rho_va, use_gpu=True, verbose=0, tight_factor=0.99)
    mod = ModelAcopf(env)
    opar = Parameters(); opar.verbose = 0
    for par in (env.params, opar):
        par.scale = scale
        for k, v in params.items():
            setattr(par, k, v)
    if "parity" in os.environ.get("EXAADMM_B200_LIB", ""):
        # the PARITY build evaluates sin / cos with the portable formulas of oracle/portable_sincos.h: put the oracle on
        # the same ones, so that the two sides round every operation of an iteration identically
        from oracle.oracle import set_portable_sincos
        set_portable_sincos(True)
    om = OracleModel(mod.grid_data, opar, rho_pq, rho_va)
    om.set_threads(min(16, os.cpu_count() or 1))
    return env, mod, om


def full_solve(config, outer_iterlim=20, inner_iterlim=1000):
    """solve_acopf's loop on both sides; status / counts / objective / residuals and the final u."""
    from exaadmm_b200.admm_two_level import admm_two_level
    env, mod, om = _models(config, outer_iterlim=outer_iterlim, inner_iterlim=inner_iterlim)
    admm_two_level(env, mod, None, mode="native")
    oi = om.admm_two_level()
    g, status = mod.info, {0: "NotSpecified", 1: "IterationLimit", 2: "Solved"}
    out = {
        "gpu": {"status": g.status, "outer": int(g.outer), "cumul": int(g.cumul), "objval": g.objval, "primres": g.primres,
                "dualres": g.dualres, "mismatch": g.mismatch},
        "oracle": {"status": status[int(oi.status)], "outer": int(oi.outer), "cumul": int(oi.cumul), "objval": oi.objval,
                   "primres": oi.primres, "dualres": oi.dualres, "mismatch": oi.mismatch},
        "max_abs_du": float(np.abs(mod.solution.u_curr - om.vec("u_curr")).max()),
        "max_abs_dv": float(np.abs(mod.solution.v_curr - om.vec("v_curr")).max()),
    }
    mod.close()
    return out


def lockstep(config, n_iter, outer_updates=()):
    """The two sides advance one inner iteration at a time from the same start; worst difference of every field over
    all iterations, and how many branches ever differ in the penalty ladder (membuf row 27)."""
    from exaadmm_b200 import operators as ops
    env, mod, om = _models(config)
    par, opar = env.params, om.par
    ops.admm_increment_outer(env, mod); ops.admm_outer_prestep(env, mod); ops.admm_increment_reset_inner(env, mod)
    om.admm_increment_outer(); om.admm_outer_prestep(); om.admm_increment_reset_inner()
    worst, norms_rel, mu_diff = {}, 0.0, 0
    for it in range(1, n_iter + 1):
        ores = np.asarray(om.inner_iteration())
        ops.admm_increment_inner(env, mod); ops.admm_inner_iteration(env, mod)
        for name in ("u_curr", "v_curr", "z_curr"):
            d = float(np.abs(getattr(mod.solution, name) - om.vec(name)).max())
            worst[name] = max(worst.get(name, 0.0), d)
        dl = float(np.abs(mod.solution.l_curr - om.vec("l_curr")).max()) / par.beta      # lambda = -(lz + beta z)
        worst["l_curr/beta"] = max(worst.get("l_curr/beta", 0.0), dl)
        got = np.array([mod.info.primres, mod.info.dualres, mod.info.norm_z_curr, mod.info.mismatch])
        norms_rel = max(norms_rel, float(np.max(np.abs(got - ores) / np.maximum(np.abs(ores), 1e-300))))
        mu_diff = max(mu_diff, int(np.count_nonzero(mod.membuf[26] != om.membuf()[26])))
        if it in outer_updates:
            ops.admm_update_lz(env, mod); om.admm_update_lz()
            par.beta *= 6.0; opar.beta *= 6.0
            ops.admm_increment_outer(env, mod); ops.admm_increment_reset_inner(env, mod)
            om.admm_increment_outer(); om.admm_increment_reset_inner()
    nline = int(mod.grid_data.nline)
    mod.close()
    return {"worst": worst, "norms_rel": norms_rel, "mu_diff": mu_diff, "nline": nline, "iterations": n_iter}


if __name__ == "__main__":
    what, config = sys.argv[1], sys.argv[2]
    args = [int(a) for a in sys.argv[3:]]
    res = full_solve(config, *args) if what == "solve" else lockstep(config, *args)
    from exaadmm_b200 import capi
    res["library"] = os.environ.get("EXAADMM_B200_LIB", str(capi.LIB_PATH))
    print("RESULT " + json.dumps(res))
