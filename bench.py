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

  value  device-resident throughput: K iterations, the two kernels of each bracketed by CUDA events, the L2
         flushed (256 MB write, outside the brackets) before every iteration
  value_l2_resident  the same K iterations back to back from the CUDA graph, no flush, no per-kernel events
  e2e    the call a user makes: host arrays -> ea_create (H2D) -> init -> full
         two-level solve -> solution back on the host (D2H); cumul / wall time
  roofline       dominant kernel (branch x-update) against the FP64 pipe,
                 plus the HBM-bound bus kernel
  cpu_baseline   the CPU oracle (restatement of the reference's use_gpu=false path)
                 on 1 host core, bounded sample

Multi-GPU (N>1, torchrun): independent load scenarios sharded over ranks, one
process per GPU, no data-path collective ("scaling": "weak").
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
    line = {
        "impl": "reference", "metric": "admm_inner_iterations_per_sec", "value": val, "unit": "iterations/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": dict(base_config(args.workload, grid, par, rho_pq, rho_va),
                       parallelism=f"{threads} host threads (OpenMP over branches and buses), rank 0 only"),
        "cpu_baseline": {"value": val, "unit": "iterations/s", "cores": threads, "kind": "port",
                         "sample": f"first {args.warmup}+{args.steps} inner iterations of the solve, full grid, "
                                   f"OpenMP over branches and buses (oracle restatement, not Julia)"},
        "e2e": {"value": val, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


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

    grid, _ = make_grid(args.workload)
    par, rho_pq, rho_va = default_params(args.workload)
    if world > 1:
        # scenario sharding: rank r solves load scenario r: every (Pd, Qd) scaled by iid U[0.99, 1.01], seed nbus + r.
        # (+-5 % per bus, the spread SURVEY 8d suggests, makes this synthetic grid stall - 20 x 1000 iterations without
        #  converging, oracle and GPU alike - while +-1 % converges like the base case: 12 outer / 447 inner iterations.)
        rng = np.random.default_rng(grid.nbus + rank)
        f = rng.uniform(0.99, 1.01, grid.nbus)
        grid.Pd = grid.Pd * f
        grid.Qd = grid.Qd * f
    nvar = 2 * grid.ngen + 8 * grid.nline
    gs, keep = make_grid_struct(grid)

    def check(rc, h=None):
        if rc != 0:
            raise RuntimeError(f"exaadmm_b200 error {rc}: {(lib.ea_last_error(h) or b'').decode()}")

    # ---------------- e2e: the call a user makes (host arrays in, solution on host out) -------------
    def solve_e2e():
        u = np.empty(nvar)
        p = params_struct(par)
        info = EaInfo()
        t0 = time.perf_counter()
        h = C.c_void_p()
        check(lib.ea_create(C.byref(gs), local_rank, C.byref(h)))
        t1 = time.perf_counter()
        check(lib.ea_init_solution(h, rho_pq, rho_va), h)
        check(lib.ea_set_option(h, b"count_work", 0.0), h)
        check(lib.ea_admm_two_level(h, C.byref(p), C.byref(info)), h)
        t2 = time.perf_counter()
        check(lib.ea_get_vector(h, 0, dptr(u), nvar), h)
        dt = time.perf_counter() - t0
        lib.ea_destroy(h)
        return info, dt, {"create_h2d_s": t1 - t0, "init_and_solve_s": t2 - t1, "d2h_s": dt - (t2 - t0)}

    solve_e2e()                                   # warm-up (context, module load)
    solve_e2e()
    info_e2e, t_e2e, e2e_split = solve_e2e()
    grid_bytes = sum(getattr(grid, n).nbytes for n in capi._GRID_DOUBLE + capi._GRID_DOUBLE_B + capi._GRID_INT_A) \
        + grid.brBusIdx.nbytes

    # ---------------- device-resident throughput over K steps ------------------------------------
    class Driver:
        """admm_two_level control flow (admm_two_level.jl:29-77), resumable in slices of n iterations."""

        def __init__(self, h):
            self.h = h
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
                        z = np.zeros(grid.nline)
                        for row in (25, 26, 27):
                            check(lib.ea_set_membuf(h, row, dptr(z), grid.nline), h)
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
        h = C.c_void_p()
        check(lib.ea_create(C.byref(gs), local_rank, C.byref(h)))
        check(lib.ea_init_solution(h, rho_pq, rho_va), h)
        check(lib.ea_set_option(h, b"count_work", 1.0 if kernel_timing else 0.0), h)
        drv = Driver(h)
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
        lib.ea_destroy(h)
        return dict(ran=ran, wall=wall, kt=list(kt), cnt=cnt.as_dict(), clocks=clocks, restarts=drv.restarts)

    # Two passes over the same K iterations of the same trajectory:
    #  B  the measurement the timing rules ask for: every iteration preceded by a 256 MB write (the 126 MB L2 is evicted;
    #     the 100 MB working set would otherwise stay resident from one iteration to the next), both kernels of every
    #     iteration bracketed by CUDA events on the library's stream, the flush outside the brackets. `value`, the
    #     rooflines and the work counters come from this pass.
    #  A  the loop as a user runs it: iterations back to back from the CUDA graph, no events, L2 as it comes
    #     (`value_l2_resident`, `gpu_launches`, the clocks sample).
    FLUSH_MB = 256
    A = trajectory(0, True)
    B = trajectory(1, False, FLUSH_MB)
    ran, dt, clocks, kt, cnt = A["ran"], A["wall"], A["clocks"], B["kt"], B["cnt"]
    t_res = A["kt"][0]               # device time inside ea_run_inner* (events around the enqueued chunks), pass A
    t_dev = B["kt"][2] + B["kt"][4]  # summed event-bracketed kernel time of the K iterations, pass B
    launches = int(A["kt"][1] + A["kt"][3] + A["kt"][5])
    if dist is not None:
        tt = torch.tensor([t_dev, dt, t_res], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_dev, dt, t_res = float(tt[0]), float(tt[1]), float(tt[2])

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
    x_bytes = 568.0 * grid.nline + 128.0 * grid.ngen
    bus_bytes = 64.0 * nvar + 48.0 * grid.nbus + 8.0 * (grid.nbus + 1)
    evals = cnt["tron_evals"]
    flops = 1645.0 * evals
    avg_x = t_x / n_x if n_x else float("nan")
    avg_b = t_b / n_b if n_b else float("nan")
    roof = {   # dominant kernel of the step (share of the step below). It is neither HBM- nor tensor-bound: it is
               # bound by the serial chain of the slowest branch; both rooflines are reported.
        "kernel": "k_xupdate (generators + branch augmented-Lagrangian / TRON solves)", "bound": "hbm",
        "achieved": x_bytes / avg_x / 1e9, "peak": hbm_peak, "unit": "GB/s", "frac": x_bytes / avg_x / 1e9 / hbm_peak,
        "traffic": 45.5e6, "traffic_source": "ncu dram__bytes_read+write per launch (cold L2, as in the timed pass), profiles/r1_fp64_ops_per_launch.csv",
        "peak_source": peak_src, "algorithmic_bytes_per_launch": x_bytes, "avg_launch_us": 1e6 * avg_x,
        "share_of_step": t_x / (t_x + t_b) if (t_x + t_b) else None,
        "fp64": {"achieved_tflops": flops / t_x / 1e12 if t_x else None, "peak_tflops": fp64_peak.value,
                 "frac": (flops / t_x / 1e12 / fp64_peak.value) if (t_x and fp64_peak.value) else None,
                 "peak_source": "ea_diag_fp64_peak (8 independent DFMA chains/thread, measured in this run)",
                 "flops_per_evaluation": 1645.0, "evaluations_per_launch": evals / n_x if n_x else None},
        "critical_path": {"max_evaluations_of_one_branch": cnt["max_evals_lane"],
                          "mean_evaluations_per_branch": evals / max(cnt["line_calls"], 1),
                          "note": "launch time ~ (evaluations of the slowest branch / 2) x ~7.5 us per serial TRON round"},
    }
    roof_bus = {"kernel": "k_bus<fused> (bus consensus + z + lambda + residual norms)", "bound": "hbm",
                "achieved": bus_bytes / avg_b / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": bus_bytes / avg_b / 1e9 / hbm_peak, "traffic": 31.9e6,
                "traffic_source": "ncu dram bytes per launch (cold cache; the 50 MB working set is L2-resident in the loop)",
                "algorithmic_bytes_per_launch": bus_bytes, "avg_launch_us": 1e6 * avg_b,
                "share_of_step": t_b / (t_x + t_b) if (t_x + t_b) else None}

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

    value = world * ran / t_dev if t_dev > 0 else None
    line = {
        "metric": "admm_inner_iterations_per_sec", "value": value, "unit": "iterations/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / ran,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {**base_config(args.workload, grid, par, rho_pq, rho_va),
                   "parallelism": "1 GPU" if world == 1 else f"{world} independent load scenarios (loads x U[0.99,1.01]), one per GPU, no collective",
                   "l2_policy": f"L2 flushed between timed iterations: a {FLUSH_MB} MB write precedes every iteration (the "
                                "100 MB working set - 15 vectors x 5.8 MB + grid - would otherwise stay in the 126 MB L2); "
                                "value = K / (sum of the CUDA-event times of the two kernels of each iteration, flush "
                                "outside the brackets); value_l2_resident = the same K iterations back to back from the "
                                "CUDA graph, no flush, no per-kernel events (what a solve sees)",
                   "restarts_in_timed_region": A["restarts"]},
        "value_l2_resident": world * ran / t_res if t_res > 0 else None,
        "ms_per_step_l2_resident": 1e3 * t_res / ran,
        "wall_ms_per_step": 1e3 * dt / ran,
        "gpu_launches": launches,
        "clocks": clocks,
        "e2e": {"value": world * info_e2e.cumul / t_e2e, "unit": "iterations/s",
                "h2d_bytes_per_step": grid_bytes / max(info_e2e.cumul, 1), "d2h_bytes_per_step": 8.0 * nvar / max(info_e2e.cumul, 1),
                "what": "ea_create(host grid arrays) + ea_init_solution + ea_admm_two_level + ea_get_vector(u) ; wall clock",
                "time_to_converge_s": t_e2e, "solver_time_s": info_e2e.time_overall, "split": e2e_split,
                "status": capi.STATUS_NAMES[info_e2e.status], "outer": info_e2e.outer, "cumul": info_e2e.cumul,
                "objval": info_e2e.objval, "mismatch": info_e2e.mismatch},
        "roofline": roof,
        "roofline_bus_kernel": roof_bus,
        "work_counters": cnt,
        "cpu_baseline": cpu,
    }
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
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
