"""Bus-partitioned solve of ONE case over several GPUs, one process per GPU.

Set-up (partition, rank-local grids, exchange maps) is `partition.py`; the per-iteration
exchange is one NCCL all-gather issued by the library on its own stream (see
`ea_set_partition` in include/exaadmm_b200.h). `torch.distributed` is used only to hand the
NCCL unique id to every rank.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

from . import capi
from .admm_two_level import admm_two_level
from .capi import dptr
from .environment import AdmmEnv
from .grid_data import GridData
from .model import ModelAcopf
from .partition import LocalGrid, build_local_grids, partition_buses

_pi = C.POINTER(C.c_int64)


def _ip(a):
    return np.ascontiguousarray(a, dtype=np.int64).ctypes.data_as(_pi)


def torch_nccl_path() -> str | None:
    """libnccl bundled with torch (so that the process uses one NCCL)."""
    try:
        import nvidia.nccl
        for p in Path(nvidia.nccl.__path__[0]).rglob("libnccl.so*"):
            return str(p)
    except Exception:
        pass
    return None


def make_partitioned_model(env: AdmmEnv, grid: GridData, part: np.ndarray, rank: int) -> tuple[ModelAcopf, LocalGrid]:
    """ModelAcopf on the rank-local grid + ea_set_partition."""
    lg = build_local_grids(grid, part, only=rank)[rank]
    mod = ModelAcopf(env, grid=lg.grid)
    keep = [np.ascontiguousarray(a, dtype=np.int64) for a in
            (lg.send_line, lg.send_end, lg.ghost_line, lg.ghost_end, lg.ghost_src_rank, lg.ghost_src_pos)]
    mod._check(mod.lib.ea_set_partition(mod.h, rank, lg.nparts, lg.n_owned_bus, len(lg.send_line), _ip(keep[0]), _ip(keep[1]),
                                        len(lg.ghost_line), _ip(keep[2]), _ip(keep[3]), _ip(keep[4]), _ip(keep[5]),
                                        lg.max_send, 2 * grid.ngen + 8 * grid.nline))
    mod.local_grid = lg
    return mod, lg


def init_comm(mod: ModelAcopf, rank: int, nccl_lib: str | None = None) -> None:
    """Create the NCCL communicator of the handle; the id travels over torch.distributed."""
    import torch.distributed as dist
    lib = mod.lib
    path = (nccl_lib or torch_nccl_path() or "").encode() or None
    buf = C.create_string_buffer(128)
    if rank == 0:
        rc = lib.ea_nccl_unique_id(path, buf)
        if rc != 0:
            raise capi.EaError(rc, (lib.ea_last_error(None) or b"").decode())
    box = [bytes(buf.raw)]
    dist.broadcast_object_list(box, src=0)
    mod._check(lib.ea_comm_init(mod.h, path, box[0]))


def init_peer_exchange(mod: ModelAcopf) -> None:
    """Switch the per-iteration exchange from ncclAllGather to direct peer-memory stores (CUDA IPC over
    NVLink): all-gather the 64-byte IPC handles of the ranks' exchange buffers and import them."""
    import torch.distributed as dist
    buf = C.create_string_buffer(64)
    mod._check(mod.lib.ea_peer_export(mod.h, buf))
    handles = [None] * dist.get_world_size()
    dist.all_gather_object(handles, bytes(buf.raw))
    mod._check(mod.lib.ea_peer_import(mod.h, b"".join(handles)))
    dist.barrier()


def solve_acopf_partitioned(case, rank: int, world: int, *, outer_iterlim=20, inner_iterlim=1000, rho_pq=400.0,
                            rho_va=40000.0, obj_scale=1.0, scale=1e-4, tight_factor=1.0, outer_eps=2e-4, gpu_no=None,
                            verbose=0, part: np.ndarray | None = None, exchange: str = "peer"):
    """`solve_acopf` for one case split over `world` GPUs (call from every rank of an
    initialised torch.distributed job). Returns (env, mod, local_grid); `mod.info` holds the
    global iteration counts / residuals / objective on every rank, `mod.solution.*` the
    rank-local vectors (assemble with partition.assemble_global)."""
    env = AdmmEnv(case, rho_pq, rho_va, use_gpu=True, tight_factor=tight_factor,
                  gpu_no=rank if gpu_no is None else gpu_no, verbose=verbose if rank == 0 else 0)
    grid = GridData.from_opfdata(env.data, tight_factor=tight_factor)
    if part is None:
        part = partition_buses(grid, world)
    mod, lg = make_partitioned_model(env, grid, part, rank)
    init_comm(mod, rank)
    if exchange == "peer":
        init_peer_exchange(mod)
    p = env.params
    p.scale, p.obj_scale, p.outer_eps, p.outer_iterlim, p.inner_iterlim = scale, obj_scale, outer_eps, outer_iterlim, inner_iterlim
    # the SAME verbose on every rank: ea_admm_two_level picks its control path (one exchange per host round trip vs
    # chunks of iterations) from it, and ranks on different paths would issue different numbers of collectives. The
    # library prints on rank 0 only; the statistics block printed from here likewise.
    p.verbose = verbose
    admm_two_level(env, mod, None, mode="native", talk=(rank == 0))
    return env, mod, lg


# ---- single-process loopback (tests): all ranks' handles on one GPU, exchange through the host ----
def loopback_iteration(mods: list[ModelAcopf], inner: int, beta: float):
    par = mods[0].env.params
    stride = 4 + 4 * mods[0].local_grid.max_send
    msgs = np.zeros((len(mods), stride))
    for r, m in enumerate(mods):
        m._check(m.lib.ea_part_begin(m.h, inner, beta, par.max_auglag, par.mu_max, par.scale))
        m._check(m.lib.ea_part_get_message(m.h, dptr(msgs[r]), stride))
    flat = np.ascontiguousarray(msgs.ravel())
    res = np.zeros(4)
    for m in mods:
        m._check(m.lib.ea_part_put_gathered(m.h, dptr(flat), flat.size))
        m._check(m.lib.ea_part_end(m.h, dptr(res)))
    return res.copy()
