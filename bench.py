#!/usr/bin/env python
"""bench.py — ADMM inner iterations/sec of the two-level ACOPF solver on B200.

Workload (BASELINE.json metric): ACTIVSg70k-like synthetic grid (70000 buses /
10390 generators / 88207 branches, seed 70000; the real file is not available
offline), rho_pq=3e4 rho_va=3e5 scale=1e-5 obj_scale=2.0 tight_factor=0.99,
outer_iterlim=20 inner_iterlim=1000 (reference README.md:55-66).

A "step" is ONE ADMM inner iteration (x-update of all generators and branches,
bus consensus, z / lambda updates, residual norms) of the real solve trajectory:
W warm-up iterations are the first W iterations of the solve, the K timed ones
follow (the two-level driver keeps running underneath: when an inner loop ends
the outer update happens inside the timed region; if the solve converges the
state is re-initialised and the trajectory restarts).

  value  device-resident throughput: K iterations, each bracketed by CUDA events (both kernels back to back inside
         the bracket), the L2 flushed (256 MB write, outside the brackets) before every iteration; a second pass
         with one bracket per kernel feeds the rooflines
  value_l2_resident  the same K iterations back to back from the CUDA graph, no flush, no per-kernel events
  e2e    the call a user makes: host arrays -> ea_create (H2D) -> init -> full
         two-level solve -> solution back on the host (D2H); cumul / wall time
  roofline       dominant kernel (branch x-update) against the FP64 pipe,
                 plus the HBM-bound bus kernel
  cpu_baseline   the CPU oracle (restatement of the reference's use_gpu=false path)
                 on 1 host core, bounded sample

Multi-GPU (N>1, torchrun, one process per GPU): the SAME solve, bus-partitioned over the N GPUs
(BASELINE configs 3-4; exaadmm.jl_b200/partition.py: graph cut, cut branches solved redundantly, one exchange
per inner iteration - xbar halves of the cut-branch ends + residual partial sums - by peer-memory stores over
NVLink fused into the bus kernel, or ncclAllGather with --exchange nccl). "scaling": "strong": value = inner
iterations of the one solve per second. Rank 0 also solves the case on one GPU and the run FAILS unless the
partitioned solve ends with the same status / outer / cumulative count and objective (`partition_parity`).
--mode replicas keeps the round-1 weak-scaling measurement (independent load scenarios, no collective);
its number is also reported as the secondary key `replicas`.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (named case, rho_pq, rho_va, scale)
    "ACTIVSg70k": ("ACTIVSg70k", 3e4, 3e5, 1e-5),
    "case13659pegase": ("case13659pegase", 5e1, 5e3, 1e-4),
    "case1354pegase": ("case1354pegase", 1e1, 1e3, 1e-4),
    "case2869pegase": ("case2869pegase", 1e1, 1e3, 1e-4),
    "tiny": ("tiny", 4e2, 4e4, 1e-4),          # 200 buses / 280 branches: latency probe (one branch per warp)
}


def make_grid(workload):
    import exaadmm_b200 as ea
    from exaadmm_b200.synthetic import named_case
    name = WORKLOADS[workload][0]
    if name == "tiny":
        from exaadmm_b200.synthetic import synthetic_case
        data = synthetic_case(200, 30, 280, seed=200)
    else:
        data = named_case(name)
    return ea.GridData.from_opfdata(data, tight_factor=0.99), data


def base_config(workload, grid, par, rho_pq, rho_va):
    """The part of `config` both arms share (same workload, same algorithm parameters)."""
    return {"workload": f"{workload}-like synthetic grid (seed=nbus; real MATPOWER file unavailable offline)",
            "nbus": grid.nbus, "ngen": grid.ngen, "nline": grid.nline, "nvar": 2 * grid.ngen + 8 * grid.nline,
            "rho_pq": rho_pq, "rho_va": rho_va, "scale": par.scale, "obj_scale": par.obj_scale, "tight_factor": 0.99,
            "outer_iterlim": 20, "inner_iterlim": 1000}


def default_params(workload):
    from exaadmm_b200.environment import Parameters
    _, rho_pq, rho_va, scale = WORKLOADS[workload]
    par = Parameters()
    par.verbose = 0
    par.scale = scale
    par.obj_scale = 2.0 if workload == "ACTIVSg70k" else 1.0     # inert
    par.outer_iterlim = 20
    par.inner_iterlim = 1000
    return par, rho_pq, rho_va


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region: NVML polled every 5 ms from a thread (a query costs well
    under a millisecond; `nvidia-smi -lms` needs ~100 ms to produce its first line, longer than a short timed region),
    plus one sample at the moment the region ends. Falls back to one nvidia-smi query if NVML is unavailable."""

    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.nvml = None
        self.hd = None
        self.sm, self.reasons, self.mx = [], set(), None
        self.running = False
        self.thread = None

    def _sample(self):
        n = self.nvml
        self.sm.append(float(n.nvmlDeviceGetClockInfo(self.hd, n.NVML_CLOCK_SM)))
        try:
            mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.hd))
        except Exception:
            mask = 0
        for name, bit in self.REASONS:
            if mask & bit:
                self.reasons.add(name)

    def _loop(self):
        while self.running:
            try:
                self._sample()
            except Exception:
                break
            time.sleep(0.005)

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.hd = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.hd, pynvml.NVML_CLOCK_SM))
            self.running = True
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()
        except Exception:
            self.nvml = None

    def stop(self):
        if self.nvml is not None:
            try:
                self._sample()                      # the region has just ended: still at its clocks
            except Exception:
                pass
            self.running = False
            if self.thread:
                self.thread.join(timeout=1.0)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.mx,
                    "reasons": sorted(self.reasons), "samples": len(self.sm), "source": "nvml, 5 ms period"}
        try:                                        # no NVML binding: one nvidia-smi query right after the region
            q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            out = subprocess.run(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                 capture_output=True, text=True, timeout=10).stdout.strip().split(",")
            p = [t.strip() for t in out]
            reasons = [n for (n, _), v in zip(self.REASONS, p[2:6]) if v.lower().startswith("active")]
            return {"sm_mhz": float(p[0]), "sm_max_mhz": float(p[1]), "reasons": sorted(reasons), "samples": 1,
                    "source": "nvidia-smi, one query at the end of the region"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi / NVML unavailable"], "samples": 0}


# ---------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """--impl reference: the CPU oracle (restatement of the reference's CPU path; the Julia
    reference cannot run here) with all host threads, same workload / metric."""
    if rank != 0:
        return
    from oracle.oracle import OracleModel
    grid, _ = make_grid(args.workload)
    par, rho_pq, rho_va = default_params(args.workload)
    threads = os.cpu_count() or 1
    m = OracleModel(grid, par, rho_pq, rho_va)
    m.set_threads(threads)
    par.beta = par.initial_beta
    m.admm_increment_outer(); m.admm_outer_prestep(); m.admm_increment_reset_inner()
    for _ in range(args.warmup):
        m.inner_iteration()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        m.inner_iteration()
    dt = time.perf_counter() - t0
    val = args.steps / dt
    # the reference's CPU path is serial (acopf_auglag_linelimit_kernel_cpu.jl:21 `for I=...`, acopf_bus_kernel_cpu.jl:11):
    # the faithful figure is ONE thread; measured on a short sample of the same iterations
    m1 = OracleModel(grid, par, rho_pq, rho_va)
    m1.set_threads(1)
    m1.admm_increment_outer(); m1.admm_outer_prestep(); m1.admm_increment_reset_inner()
    n1 = max(3, min(args.steps, 10))
    t1 = time.perf_counter()
    for _ in range(n1):
        m1.inner_iteration()
    val1 = n1 / (time.perf_counter() - t1)
    line = {
        "impl": "reference", "metric": "admm_inner_iterations_per_sec", "value": val, "unit": "iterations/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": dict(base_config(args.workload, grid, par, rho_pq, rho_va),
                       parallelism=f"{threads} host threads (OpenMP over branches and buses), rank 0 only"),
        "cpu_baseline": {"value": val, "unit": "iterations/s", "cores": threads, "kind": "port",
                         "sample": f"first {args.warmup}+{args.steps} inner iterations of the solve, full grid; C restatement "
                                   f"of the reference's CPU path with OpenMP over branches and buses on {threads} threads - "
                                   f"NOT reference behaviour (the reference's CPU loops are serial, and it is Julia)",
                         "value_1_thread": val1,
                         "sample_1_thread": f"first {n1} inner iterations, serial loops as in the reference"},
        "e2e": {"value": val, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------
def scenario_block(workload, ids, local_rank, budget, oracle_check=True):
    """BASELINE config 5 on ONE GPU: the load scenarios `ids` of `workload` (every Pd, Qd x iid U[0.95, 1.05], seed
    nbus + s) solved (a) concurrently as stand-alone solves, one CUDA stream and one host thread each - other
    scenarios' kernels fill the SMs a solve leaves idle while its slowest branch finishes -, (b) as ONE batch
    (ea_batch_*: one branch kernel over all (scenario, branch) pairs, per-scenario termination), (c) one after the other.
    Whole solves with an (outer, inner) budget; returns iterations/s of each and the per-scenario outcomes."""
    from exaadmm_b200.scenarios import ScenarioBatch, solve_scenarios, scenario_loads
    grid_base, data = make_grid(workload)
    par, rho_pq, rho_va = default_params(workload)
    kw = dict(scale=par.scale, outer_iterlim=budget[0], inner_iterlim=budget[1])
    out = {"workload": f"{workload}-like synthetic grid, {len(ids)} load scenarios on this GPU (loads x U[0.95,1.05], seed nbus + s)",
           "rho_pq": rho_pq, "rho_va": rho_va, "scale": par.scale, "outer_iterlim": budget[0], "inner_iterlim": budget[1]}
    best = {}
    for name, conc in (("concurrent_streams", len(ids)), ("sequential", 1)):
        for rep in range(2 if name == "concurrent_streams" else 1):
            res, wall = solve_scenarios(data, ids, rho_pq=rho_pq, rho_va=rho_va, tight_factor=0.99, max_concurrent=conc,
                                        gpu_no=local_rank, **kw)
            cum = [int(m.info.cumul) for _, m in res]
            st = [m.info.status for _, m in res]
            obj0 = res[0][1].info.objval
            for _, m in res:
                m.close()
            if name not in best or wall < best[name][0]:
                best[name] = (wall, cum, st, obj0)
    for rep in range(2):
        b = ScenarioBatch(data, ids, rho_pq=rho_pq, rho_va=rho_va, tight_factor=0.99, gpu_no=local_rank)
        wall = b.solve(**kw)
        cum = [int(m.info.cumul) for m in b.models]
        st = [m.info.status for m in b.models]
        obj0 = b.models[0].info.objval
        b.close()
        if "batched" not in best or wall < best["batched"][0]:
            best["batched"] = (wall, cum, st, obj0)
    for name, (wall, cum, st, obj0) in best.items():
        out[name] = {"iterations_per_s": sum(cum) / wall, "seconds": wall, "iterations": sum(cum),
                     "solved": st.count("Solved"), "cumul_min": min(cum), "cumul_max": max(cum)}
    out["same_iteration_counts_in_all_three"] = best["batched"][1] == best["concurrent_streams"][1] == best["sequential"][1]
    if oracle_check:
        from oracle.oracle import OracleModel
        par.outer_iterlim, par.inner_iterlim = budget
        om = OracleModel(grid_base, par, rho_pq, rho_va)
        om.set_threads(os.cpu_count() or 1)
        Pd, Qd = scenario_loads(grid_base, ids[0])
        om.set_load(Pd, Qd)
        oi = om.admm_two_level()
        out["oracle_check_scenario0"] = {"cumul_equal": int(oi.cumul) == best["batched"][1][0],
                                         "objective_rel_diff": abs(best["batched"][3] - oi.objval) / abs(oi.objval)}
    return out


def run_scenarios(args, rank, world, local_rank):
    """--scenarios S: config 5 as the headline of the line - S scenarios sharded over the ranks (rank r takes r, r+N, ...)."""
    import torch
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ids = list(range(rank, args.scenarios, world))
    blk = scenario_block(args.workload, ids, local_rank, (args.scenario_outer, args.scenario_inner), oracle_check=(rank == 0))
    its = torch.tensor([blk["concurrent_streams"]["iterations"], blk["batched"]["iterations"]], dtype=torch.float64, device="cuda")
    tm = torch.tensor([blk["concurrent_streams"]["seconds"], blk["batched"]["seconds"]], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(its, op=dist.ReduceOp.SUM)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    if rank == 0:
        grid, _ = make_grid(args.workload)
        par, rho_pq, rho_va = default_params(args.workload)
        v_streams, v_batch = float(its[0] / tm[0]), float(its[1] / tm[1])
        line = {"metric": "admm_inner_iterations_per_sec", "value": max(v_streams, v_batch), "unit": "iterations/s", "n_gpus": world,
                "value_method": "batched (ea_batch_*)" if v_batch >= v_streams else "concurrent stand-alone solves",
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / max(v_streams, v_batch), "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {**base_config(args.workload, grid, par, rho_pq, rho_va),
                           "outer_iterlim": args.scenario_outer, "inner_iterlim": args.scenario_inner,
                           "parallelism": f"{args.scenarios} load scenarios sharded over {world} GPU(s) ({len(ids)} on rank 0), whole "
                                          "solves; per GPU either one batch (one branch kernel over all (scenario, branch) "
                                          "pairs) or concurrent stand-alone solves (one stream + host thread each); no "
                                          "data-path collective",
                           "l2_policy": "whole solves back to back (the scenarios' working sets together exceed nothing: "
                                        "64 x 4.5 MB at 2869 buses - L2 resident, as in production)"},
                "value_batched": v_batch, "value_concurrent_streams": v_streams,
                "e2e": {"value": max(v_streams, v_batch), "unit": "iterations/s", "what": "wall clock of the solves incl. the D2H "
                        "of control blocks / norms; model creation (H2D) excluded",
                        "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 32},
                "gpu_launches": int(2 * float(its[0])),
                "scenario_block_rank0": blk}
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch
    from exaadmm_b200 import capi
    from exaadmm_b200.capi import EaInfo, EaCounters, dptr, make_grid_struct, params_struct

    lib = capi.load_library()
    if lib.ea_device_count() < 1:
        raise RuntimeError("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU baseline)")
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    grid, data = make_grid(args.workload)
    par, rho_pq, rho_va = default_params(args.workload)
    partitioned = world > 1 and args.mode == "partitioned"
    nvar = 2 * grid.ngen + 8 * grid.nline                    # of the whole problem (tolerances scale with sqrt(nvar))

    def check(rc, h=None):
        if rc != 0:
            raise RuntimeError(f"exaadmm_b200 error {rc}: {(lib.ea_last_error(h) or b'').decode()}")

    def scenario_grid(r):
        # scenario sharding (--mode replicas): rank r solves load scenario r: every (Pd, Qd) scaled by iid U[0.99, 1.01],
        # seed nbus + r. (+-5 % per bus, the spread SURVEY 8d suggests, makes this synthetic grid stall - 20 x 1000
        # iterations without converging, oracle and GPU alike - while +-1 % converges like the base case.)
        import copy
        g = copy.copy(grid)
        f = np.random.default_rng(grid.nbus + r).uniform(0.99, 1.01, grid.nbus)
        g.Pd = grid.Pd * f
        g.Qd = grid.Qd * f
        return g

    # ---- what one rank holds: the whole grid (1 GPU / replicas) or its part of the bus partition ----------------------
    part_info = None
    if partitioned:
        from exaadmm_b200.environment import AdmmEnv
        from exaadmm_b200.partition import partition_buses, cut_statistics
        from exaadmm_b200.partitioned import make_partitioned_model, init_comm, init_peer_exchange
        part = partition_buses(grid, world)
        env = AdmmEnv(data, rho_pq, rho_va, use_gpu=True, tight_factor=0.99, gpu_no=local_rank, verbose=0)

        def new_handle():
            t0 = time.perf_counter()
            mod, lg = make_partitioned_model(env, grid, part, rank)        # ea_create(rank-local grid) + ea_set_partition
            t1 = time.perf_counter()
            init_comm(mod, rank)                                          # NCCL communicator of the handle (scalar collectives)
            if args.exchange == "peer":
                init_peer_exchange(mod)                                   # CUDA IPC: peer stores replace the all-gather
            part_info["message_bytes_per_rank_per_iteration"] = 8 * (4 + 4 * lg.max_send)
            return mod, mod.h, lg.grid, t1 - t0
        stats = cut_statistics(grid, part)
        part_info = {"parts": world, "cut": stats, "exchange": args.exchange}
    else:
        my_grid = scenario_grid(rank) if world > 1 else grid
        gs, keep = make_grid_struct(my_grid)

        def new_handle():
            t0 = time.perf_counter()
            h = C.c_void_p()
            check(lib.ea_create(C.byref(gs), local_rank, C.byref(h)))
            return None, h, my_grid, time.perf_counter() - t0

    def drop_handle(mod, h):
        if mod is not None:
            mod.close()
        else:
            lib.ea_destroy(h)

    # ---------------- e2e: the call a user makes (host arrays in, solution on host out) -------------
    def solve_e2e():
        p = params_struct(par)
        info = EaInfo()
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        mod, h, lgrid, t_create = new_handle()
        t1 = time.perf_counter()
        check(lib.ea_init_solution(h, rho_pq, rho_va), h)
        check(lib.ea_set_option(h, b"count_work", 0.0), h)
        check(lib.ea_admm_two_level(h, C.byref(p), C.byref(info)), h)
        t2 = time.perf_counter()
        n_local = 2 * lgrid.ngen + 8 * lgrid.nline
        u = np.empty(n_local)
        check(lib.ea_get_vector(h, 0, dptr(u), n_local), h)
        t3 = time.perf_counter()
        drop_handle(mod, h)
        # partitioned: the communicator / IPC set-up of the handle (one-off per process group in a deployment) is timed
        # separately and not part of the end-to-end figure; ea_create + H2D of the local grid is
        dt = t_create + (t3 - t1)
        return info, dt, {"create_h2d_s": t_create, "comm_setup_s": (t1 - t0) - t_create, "init_and_solve_s": t2 - t1,
                          "d2h_s": t3 - t2}, lgrid

    solve_e2e()                                   # warm-up (context, module load)
    if not partitioned:
        solve_e2e()
    info_e2e, t_e2e, e2e_split, lgrid = solve_e2e()
    grid_bytes = sum(getattr(lgrid, n).nbytes for n in capi._GRID_DOUBLE + capi._GRID_DOUBLE_B + capi._GRID_INT_A) \
        + lgrid.brBusIdx.nbytes
    d2h_bytes = 8.0 * (2 * lgrid.ngen + 8 * lgrid.nline)
    if dist is not None:
        tt = torch.tensor([t_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_e2e = float(tt[0])

    # ---- partitioned: the same solve on ONE GPU (rank 0) must end identically --------------------------------------
    parity = None
    if partitioned and rank == 0:
        gs1, keep1 = make_grid_struct(grid)
        h1 = C.c_void_p()
        check(lib.ea_create(C.byref(gs1), local_rank, C.byref(h1)))
        check(lib.ea_init_solution(h1, rho_pq, rho_va), h1)
        p1 = params_struct(par); i1 = EaInfo()
        check(lib.ea_admm_two_level(h1, C.byref(p1), C.byref(i1)), h1)
        lib.ea_destroy(h1)
        parity = {"single_gpu": {"status": capi.STATUS_NAMES[i1.status], "outer": i1.outer, "cumul": i1.cumul, "objval": i1.objval,
                                 "mismatch": i1.mismatch, "solver_time_s": i1.time_overall},
                  "partitioned": {"status": capi.STATUS_NAMES[info_e2e.status], "outer": info_e2e.outer, "cumul": info_e2e.cumul,
                                  "objval": info_e2e.objval, "mismatch": info_e2e.mismatch,
                                  "solver_time_s": info_e2e.time_overall}}
        same = (i1.status, i1.outer, i1.cumul) == (info_e2e.status, info_e2e.outer, info_e2e.cumul) and \
            abs(i1.objval - info_e2e.objval) <= 1e-9 * abs(i1.objval)
        parity["identical_counts_and_objective"] = bool(same)
        if not same:
            raise RuntimeError(f"bench.py: the {world}-GPU partitioned solve differs from the single-GPU solve: {parity}")

    # ---------------- device-resident throughput over K steps ------------------------------------
    class Driver:
        """admm_two_level control flow (admm_two_level.jl:29-77), resumable in slices of n iterations."""

        def __init__(self, h, nline):
            self.h = h
            self.nline = nline
            self.restarts = 0
            self.reset()

        def reset(self):
            self.outer = 0; self.inner = 0; self.beta = par.initial_beta
            self.need_outer = True; self.solved = False
            self.res = np.zeros(4); self.norm_z_prev = 0.0

        def run(self, n):
            h = self.h
            done = 0
            sqrt_d = math.sqrt(nvar)
            while done < n:
                if self.need_outer:
                    if self.solved or self.outer >= par.outer_iterlim:
                        check(lib.ea_init_solution(h, rho_pq, rho_va), h)
                        z = np.zeros(self.nline)
                        for row in (25, 26, 27):
                            check(lib.ea_set_membuf(h, row, dptr(z), self.nline), h)
                        self.restarts += 1
                        self.reset()
                    self.outer += 1
                    nz = C.c_double()
                    check(lib.ea_outer_prestep(h, C.byref(nz)), h)
                    self.norm_z_prev = nz.value
                    self.inner = 0
                    self.need_outer = False
                budget = min(n - done, par.inner_iterlim - self.inner)
                got = C.c_int64()
                check(lib.ea_run_inner_from(h, self.outer, self.beta, self.inner, self.inner + budget, par.max_auglag,
                                            par.mu_max, par.scale, 0, C.byref(got), dptr(self.res)), h)
                done += got.value - self.inner
                self.inner = got.value
                ended = self.res[0] <= sqrt_d / (2500 * self.outer) or self.inner >= par.inner_iterlim
                if ended:
                    if self.res[3] <= sqrt_d * par.outer_eps:
                        self.solved = True
                    else:
                        check(lib.ea_update_lz(h, self.beta, par.MAX_MULTIPLIER), h)
                        if self.res[2] > par.theta * self.norm_z_prev:
                            self.beta = min(par.inc_c * self.beta, 1e24)
                    self.need_outer = True
            return done

    def trajectory(kernel_timing, sample_clocks, flush_mb=0):
        """init -> W warm-up iterations -> K timed iterations of the same solve trajectory."""
        mod, h, lg, _ = new_handle()
        check(lib.ea_init_solution(h, rho_pq, rho_va), h)
        check(lib.ea_set_option(h, b"count_work", 1.0 if kernel_timing == 1 else 0.0), h)    # the counters feed the rooflines (pass B)
        drv = Driver(h, lg.nline)
        drv.run(args.warmup)
        check(lib.ea_reset_counters(h), h)
        check(lib.ea_set_option(h, b"kernel_timing", float(kernel_timing)), h)
        if flush_mb:
            check(lib.ea_set_option(h, b"l2_flush_mb", float(flush_mb)), h)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        sampler = ClockSampler(local_rank) if sample_clocks else None
        if sampler:
            sampler.start()
        t0 = time.perf_counter()
        ran = drv.run(args.steps)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        if dist is not None:
            dist.barrier()
        clocks = sampler.stop() if sampler else None
        kt = (C.c_double * 8)()
        check(lib.ea_get_kernel_times(h, kt), h)      # CUDA events on the library's own stream
        cnt = EaCounters()
        check(lib.ea_get_counters(h, C.byref(cnt)), h)
        drop_handle(mod, h)
        return dict(ran=ran, wall=wall, kt=list(kt), cnt=cnt.as_dict(), clocks=clocks, restarts=drv.restarts)

    # Two passes over the same K iterations of the same trajectory:
    #  B  the measurement the timing rules ask for: every iteration preceded by a 256 MB write (the 126 MB L2 is evicted;
    #     the 100 MB working set would otherwise stay resident from one iteration to the next), both kernels of every
    #     iteration bracketed by CUDA events on the library's stream, the flush outside the brackets. `value`, the
    #     rooflines and the work counters come from this pass.
    #  A  the loop as a user runs it: iterations back to back from the CUDA graph, no events, L2 as it comes
    #     (`value_l2_resident`, `gpu_launches`, the clocks sample).
    #  C  pass B with ONE bracket per iteration (both kernels back to back inside it) instead of one per kernel: an event
    #     between two kernels costs ~4 us of launch latency that the loop does not have. `value` comes from this pass.
    FLUSH_MB = 256
    A = trajectory(0, True)
    B = trajectory(1, False, FLUSH_MB)
    Cp = trajectory(2, False, FLUSH_MB)
    ran, dt, clocks, kt, cnt = A["ran"], A["wall"], A["clocks"], B["kt"], B["cnt"]
    t_res = A["kt"][0]               # device time inside ea_run_inner* (events around the enqueued chunks), pass A
    t_dev = Cp["kt"][2]              # summed event-bracketed time of the K iterations, pass C
    t_dev_kernels = B["kt"][2] + B["kt"][4]
    launches = int(A["kt"][1] + A["kt"][3] + A["kt"][5])
    t_x_all = t_b_all = None
    if dist is not None:
        tt = torch.tensor([t_dev, dt, t_res, B["kt"][2], B["kt"][4]], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_dev, dt, t_res, t_x_all, t_b_all = (float(v) for v in tt)
        tl = torch.tensor([launches], device="cuda", dtype=torch.int64)
        dist.all_reduce(tl, op=dist.ReduceOp.SUM)
        launches = int(tl[0])

    # secondary (N > 1, partitioned mode): the round-1 weak-scaling figure - N independent load scenarios, no collective
    replicas = None
    if partitioned and not args.no_replicas:
        gs_r, keep_r = make_grid_struct(scenario_grid(rank))
        hr = C.c_void_p()
        check(lib.ea_create(C.byref(gs_r), local_rank, C.byref(hr)))
        check(lib.ea_init_solution(hr, rho_pq, rho_va), hr)
        check(lib.ea_set_option(hr, b"count_work", 0.0), hr)
        drv = Driver(hr, grid.nline)
        drv.run(args.warmup)
        check(lib.ea_reset_counters(hr), hr)
        dist.barrier(); torch.cuda.synchronize()
        ran_r = drv.run(args.steps)
        torch.cuda.synchronize()
        ktr = (C.c_double * 8)()
        check(lib.ea_get_kernel_times(hr, ktr), hr)
        lib.ea_destroy(hr)
        tr = torch.tensor([ktr[0]], device="cuda", dtype=torch.float64)
        dist.all_reduce(tr, op=dist.ReduceOp.MAX)
        replicas = {"value": world * ran_r / float(tr[0]), "unit": "iterations/s", "scaling": "weak",
                    "what": f"{world} independent load scenarios (loads x U[0.99,1.01]), one per GPU, no collective; "
                            "iterations back to back from the CUDA graph (as value_l2_resident)"}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---------------- roofline -------------------------------------------------------------------
    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    fp64_peak = C.c_double(0.0)
    check(lib.ea_diag_fp64_peak(local_rank, C.byref(fp64_peak)))     # DFMA probe on this device (not in MEASURED_PEAKS.json)
    n_x, t_x, n_b, t_b = kt[1], kt[2], kt[3], kt[4]          # launches and summed CUDA-event durations (s), pass B
    # Algorithmic work per launch:
    #   branch kernel: 568 B per branch (60 reads + 11 writes) + 128 B per generator;
    #                  FP64 flops = 1645 per objective evaluation incl. its share of the TRON algebra
    #                  (ncu op counts / evaluations, profiles/r1_fp64_ops_per_launch.csv)
    #   bus kernel:    64 B per entry (u,z,lambda,rho,lz read; v,z,lambda written) + 48 B per bus + CSR pointers
    ln, gn, bn = lgrid.nline, lgrid.ngen, lgrid.nbus          # what ONE launch of this rank processes
    x_bytes = 568.0 * ln + 128.0 * gn
    bus_bytes = 64.0 * (2 * gn + 8 * ln) + 48.0 * bn + 8.0 * (bn + 1)
    evals = cnt["tron_evals"]
    # ALGORITHMIC flops: per objective evaluation f = 119; gradient + Hessian = 350 + 1006 at every accepted
    # point and at the start of every TRON solve (the reference evaluates them there only); dtron per step (Cauchy search:
    # 2 matrix-vector products + projections ~200, Cholesky 72, projected search ~150, model / norms ~100 = 522) and per
    # CG iteration (1 matrix-vector product + 2 triangular solves = 144). Steps = evaluations - TRON solves.
    n_gh = evals - cnt["rejected_steps"]
    n_steps = max(evals - cnt["auglag_iters"], 0)
    flops_alg = 119.0 * evals + 1356.0 * n_gh + 522.0 * n_steps + 144.0 * cnt["cg_iters"]
    flops_exec = 828.0 * evals        # executed DFMA x 2 + DMUL + DADD per evaluation (ncu, profiles/r2_xupdate.txt: 232.5 MFLOP in a
                                      # launch of 281 k evaluations; 1645 in round 1)
    avg_x = t_x / n_x if n_x else float("nan")
    avg_b = t_b / n_b if n_b else float("nan")
    roof = {   # dominant kernel of the step. An FP64-pipe kernel by its arithmetic (nothing here is a contraction; 568 B per
               # branch make it 2 % of the HBM roofline) - in fact bound by the serial chain of its slowest branch.
        "kernel": "k_xupdate (generators + branch augmented-Lagrangian / TRON solves)", "bound": "fp64",
        "achieved": flops_alg / t_x / 1e12 if t_x else None, "peak": fp64_peak.value, "unit": "TFLOP/s",
        "frac": (flops_alg / t_x / 1e12 / fp64_peak.value) if (t_x and fp64_peak.value) else None,
        "peak_source": "ea_diag_fp64_peak: 8 independent DFMA chains per thread, measured in this run (MEASURED_PEAKS.json has "
                       "no FP64 figure)",
        "algorithmic_flops_per_launch": flops_alg / n_x if n_x else None,
        "algorithmic_flops": "119 nfev + 1356 ngev + 522 steps + 144 cg (SURVEY 8d; counters of this pass)",
        "executed_flops_per_launch": flops_exec / n_x if n_x else None,
        "frac_executed": (flops_exec / t_x / 1e12 / fp64_peak.value) if (t_x and fp64_peak.value) else None,
        "traffic": 44.7e6, "traffic_source": "ncu dram__bytes_read+write per launch at the 1-GPU size (cold L2, as in the timed "
                                            "pass), profiles/r2_xupdate.txt",
        "avg_launch_us": 1e6 * avg_x, "share_of_step": t_x / (t_x + t_b) if (t_x + t_b) else None,
        "hbm": {"achieved": x_bytes / avg_x / 1e9, "peak": hbm_peak, "unit": "GB/s", "frac": x_bytes / avg_x / 1e9 / hbm_peak,
                "algorithmic_bytes_per_launch": x_bytes, "peak_source": peak_src},
        "evaluations_per_launch": evals / n_x if n_x else None,
        "critical_path": {"max_evaluations_of_one_branch": cnt["max_evals_lane"],
                          "mean_evaluations_per_branch": evals / max(cnt["line_calls"], 1),
                          "note": "launch time ~ work queue drained (35-45 us) + the rest of the slowest branch at ~2 us per evaluation "
                                  "(one lane, serial: 4.4 us per augmented-Lagrangian iteration; DESIGN.md section 6b)"},
    }
    roof_bus = {"kernel": "k_bus<fused> (bus consensus + z + lambda + residual norms)", "bound": "hbm",
                "achieved": bus_bytes / avg_b / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": bus_bytes / avg_b / 1e9 / hbm_peak, "traffic": 34.7e6,
                "traffic_source": "ncu dram bytes per launch, profiles/r2_bus.txt (cold cache: reads; the 17 MB it writes stay in "
                                  "the L2; in the loop the 50 MB working set is L2-resident)",
                "algorithmic_bytes_per_launch": bus_bytes, "avg_launch_us": 1e6 * avg_b,
                "peak_source": peak_src,
                "share_of_step": t_b / (t_x + t_b) if (t_x + t_b) else None}
    if partitioned:
        roof_bus["kernel"] = "k_bus<fused> + exchange (peer stores / all-gather) + k_finish (ghost install, norms, termination)"

    # ---------------- CPU baseline (1 core, bounded sample) ----------------------------------------
    cpu = None
    if not args.no_cpu_baseline:
        from oracle.oracle import OracleModel
        om = OracleModel(grid, par, rho_pq, rho_va)
        om.set_threads(1)
        par.beta = par.initial_beta
        om.admm_increment_outer(); om.admm_outer_prestep(); om.admm_increment_reset_inner()
        n_cpu = args.cpu_steps
        tc = time.perf_counter()
        for _ in range(n_cpu):
            om.inner_iteration()
        tc = time.perf_counter() - tc
        cpu = {"value": n_cpu / tc, "unit": "iterations/s", "cores": 1, "kind": "port",
               "sample": f"first {n_cpu} inner iterations of the same solve on the full grid, serial loops "
                         f"(oracle restatement of the reference's use_gpu=false path, not Julia)"}

    # ---------------- BASELINE configs 2 and 3 in the same line (1 GPU): whole solves ---------------------------
    other = None
    if world == 1 and not args.no_other_configs and args.workload == "ACTIVSg70k":
        other = {}
        # (outer, inner) budget: the README's rho of the pegase files stalls on the synthetic stand-ins and the
        # stalled regime is slow (thousands of branches walk the whole penalty ladder every iteration): 3 x 300 there
        for name, wl, rpq, rva, budget in (("config2_case1354pegase", "case1354pegase", 1e1, 1e3, (3, 300)),
                                           ("config2_case1354pegase_converging_rho", "case1354pegase", 4e2, 4e4, (20, 1000)),
                                           ("config3_case13659pegase", "case13659pegase", 5e1, 5e3, (3, 300))):
            g2, _ = make_grid(wl)
            p2, _, _ = default_params(wl)
            p2.outer_iterlim, p2.inner_iterlim = budget
            gs2, keep2 = make_grid_struct(g2)
            best = None
            for rep in range(2):
                h2 = C.c_void_p()
                t0 = time.perf_counter()
                check(lib.ea_create(C.byref(gs2), local_rank, C.byref(h2)))
                check(lib.ea_init_solution(h2, rpq, rva), h2)
                ps = params_struct(p2); i2 = EaInfo()
                check(lib.ea_admm_two_level(h2, C.byref(ps), C.byref(i2)), h2)
                u2 = np.empty(2 * g2.ngen + 8 * g2.nline)
                check(lib.ea_get_vector(h2, 0, dptr(u2), u2.size), h2)
                t = time.perf_counter() - t0
                lib.ea_destroy(h2)
                if best is None or t < best[0]:
                    best = (t, i2)
            t, i2 = best
            st = capi.STATUS_NAMES[i2.status]
            other[name] = {"workload": f"{wl}-like synthetic grid", "nbus": g2.nbus, "nline": g2.nline, "rho_pq": rpq, "rho_va": rva,
                           "scale": p2.scale, "outer_iterlim": budget[0], "inner_iterlim": budget[1], "status": st, "outer": i2.outer, "cumul": i2.cumul, "objval": i2.objval,
                           "iterations_per_s_e2e": i2.cumul / t, "seconds_e2e": t,
                           "time_to_converge_s": t if st == "Solved" else None,
                           "note": None if st == "Solved" else "the README's rho for the real pegase file stalls on the synthetic "
                                   "stand-in (oracle alike, SURVEY 8d): throughput over a 3 x 300 budget"}

    config5 = None
    if world == 1 and not args.no_other_configs and args.workload == "ACTIVSg70k":
        # what ONE GPU does in BASELINE config 5 (64 scenarios of case2869pegase over 8 GPUs): 8 scenarios
        config5 = scenario_block("case2869pegase", list(range(8)), local_rank, (3, 300))

    window = (f"inner iterations {args.warmup + 1}-{args.warmup + args.steps} of the solve trajectory from the flat start "
              f"(12 outer / 447 inner in all; the first ~10 inner iterations of every outer iteration restart the penalty "
              f"ladder and are the expensive ones: a short window early in the solve, like the driver's --steps 20 --warmup 5, "
              f"measures that regime; steady state is ~1.5x faster)")
    value = (ran if partitioned else world * ran) / t_dev if t_dev > 0 else None
    if world == 1:
        parallelism = "1 GPU"
    elif partitioned:
        parallelism = (f"ONE solve bus-partitioned over {world} GPUs (graph cut: {part_info['cut']['cut_lines']} of {grid.nline} "
                       f"branches cut and solved redundantly), one exchange per inner iteration by "
                       + ("peer-memory stores over NVLink fused into the bus kernel" if args.exchange == "peer"
                          else "ncclAllGather on the library's stream"))
    else:
        parallelism = f"{world} independent load scenarios (loads x U[0.99,1.01]), one per GPU, no collective"
    e2e_mult = 1 if (partitioned or world == 1) else world
    line = {
        "metric": "admm_inner_iterations_per_sec", "value": value, "unit": "iterations/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / ran,
        "higher_is_better": True, "scaling": "strong" if partitioned else "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {**base_config(args.workload, grid, par, rho_pq, rho_va),
                   "parallelism": parallelism,
                   "timed_window": window,
                   "l2_policy": f"L2 flushed between timed iterations: a {FLUSH_MB} MB write precedes every iteration (the "
                                "100 MB working set - 15 vectors x 5.8 MB + grid - would otherwise stay in the 126 MB L2); "
                                "value = K / (sum over the K iterations of the CUDA-event time of the iteration - both "
                                "kernels back to back inside one bracket, flush outside the brackets); the rooflines come "
                                "from a second pass with one bracket per kernel (ms_per_step_kernel_brackets: an event "
                                "between two kernels adds ~4 us); value_l2_resident = the same K iterations back to back "
                                "from the CUDA graph, no flush, no events inside (what a solve sees)",
                   "restarts_in_timed_region": A["restarts"]},
        "value_l2_resident": (ran if partitioned else world * ran) / t_res if t_res > 0 else None,
        "ms_per_step_l2_resident": 1e3 * t_res / ran,
        "wall_ms_per_step": 1e3 * dt / ran,
        "ms_per_step_kernel_brackets": 1e3 * t_dev_kernels / ran,
        "gpu_launches": launches,
        "clocks": clocks,
        "e2e": {"value": e2e_mult * info_e2e.cumul / t_e2e, "unit": "iterations/s",
                "h2d_bytes_per_step": grid_bytes / max(info_e2e.cumul, 1), "d2h_bytes_per_step": d2h_bytes / max(info_e2e.cumul, 1),
                "what": ("ea_create(host grid arrays of the rank's part) + ea_set_partition + ea_init_solution + ea_admm_two_level + "
                         "ea_get_vector(u); max over ranks; the handle's NCCL communicator / IPC mapping (comm_setup_s, one-off per "
                         "process group) is timed separately and not included; h2d / d2h bytes are per rank") if partitioned else
                        "ea_create(host grid arrays) + ea_init_solution + ea_admm_two_level + ea_get_vector(u) ; wall clock",
                "time_to_converge_s": t_e2e, "solver_time_s": info_e2e.time_overall, "split": e2e_split,
                "status": capi.STATUS_NAMES[info_e2e.status], "outer": info_e2e.outer, "cumul": info_e2e.cumul,
                "objval": info_e2e.objval, "mismatch": info_e2e.mismatch},
        "roofline": roof,
        "roofline_bus_kernel": roof_bus,
        "work_counters": cnt,
        "cpu_baseline": cpu,
    }
    if partitioned:
        line["partition"] = {**part_info, "local_lines_rank0": lgrid.nline, "local_buses_rank0": lgrid.nbus,
                             "x_update_us_max_over_ranks": 1e6 * t_x_all / n_x if n_x else None,
                             "bus_exchange_finish_us_max_over_ranks": 1e6 * t_b_all / n_b if n_b else None}
        line["partition_parity"] = parity
        line["replicas"] = replicas
    if other is not None:
        line["other_configs"] = other
    if config5 is not None:
        line["config5_scenarios_8_of_64_per_gpu"] = config5
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="ACTIVSg70k", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-steps", type=int, default=40)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mode", default="partitioned", choices=["partitioned", "replicas"],
                    help="N > 1: one bus-partitioned solve over the N GPUs (strong scaling) or N independent load scenarios")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="partitioned mode: per-iteration exchange by peer-memory stores (CUDA IPC over NVLink) or ncclAllGather")
    ap.add_argument("--no-replicas", action="store_true", help="partitioned mode: skip the secondary weak-scaling measurement")
    ap.add_argument("--no-other-configs", action="store_true", help="1 GPU: skip the whole solves of BASELINE configs 2, 3 and 5")
    ap.add_argument("--scenarios", type=int, default=0,
                    help="BASELINE config 5 as the headline: this many load scenarios of --workload, sharded over the GPUs")
    ap.add_argument("--scenario-outer", type=int, default=3)
    ap.add_argument("--scenario-inner", type=int, default=300)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args, rank, world)
    elif args.scenarios > 0:
        run_scenarios(args, rank, world, local_rank)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
