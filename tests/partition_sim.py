"""CPU simulation of the partitioned inner iteration (test infrastructure).

Each rank drives the CPU oracle on its rank-local grid (exaadmm_b200.partition) and the
ranks exchange exactly what the GPU path exchanges - the xbar half of every cut branch end
plus four residual partial sums, in one all-gather - over torch.distributed (gloo). Used by
tests/test_partition.py to validate the partition / ownership / exchange maps against the
single-domain oracle without a GPU.
"""
import numpy as np
import torch
import torch.distributed as dist

from exaadmm_b200.environment import Parameters
from oracle.oracle import OracleModel


class PartitionedOracle:
    def __init__(self, lg, par: Parameters, rho_pq: float, rho_va: float):
        self.lg = lg
        self.par = par
        self.m = OracleModel(lg.grid, par, rho_pq, rho_va)
        self.send_idx = lg.half_positions(lg.send_line, lg.send_end)          # (n_send, 4)
        self.ghost_idx = lg.half_positions(lg.ghost_line, lg.ghost_end)       # (n_ghost, 4)
        self.stride = 4 + 4 * lg.max_send
        self.inner = 0

    def inner_iteration(self):
        m, lg = self.m, self.lg
        m.admm_increment_inner(); m.admm_inner_prestep(); m.admm_update_x()
        m.admm_update_xbar()                       # ghost buses produce garbage for ghost halves: overwritten below
        v = m.vec("v_curr")
        # phase 1 of the exchange: xbar halves of the owned cut-branch ends
        seg = np.zeros(self.stride)
        seg[4:4 + 4 * len(lg.send_line)] = v[self.send_idx].ravel()
        gathered = self._allgather(seg)
        src = gathered[lg.ghost_src_rank, :][:, 4:].reshape(len(lg.ghost_line), -1, 4) if len(lg.ghost_line) else None
        if src is not None:
            v[self.ghost_idx] = src[np.arange(len(lg.ghost_line)), lg.ghost_src_pos]
        m.admm_update_z(); m.admm_update_l()       # redundant for ghost entries, from the received xbar
        # phase 2: residual partial sums over owned entries (the GPU path packs these into the same message)
        u, z, zp = m.vec("u_curr"), m.vec("z_curr"), m.vec("z_prev")
        o = lg.owned_entry
        rp = u[o] - v[o] + z[o]
        part = np.array([np.sum(rp * rp), np.sum((z[o] - zp[o]) ** 2), np.sum(z[o] ** 2), np.sum((rp - z[o]) ** 2)])
        seg2 = np.zeros(self.stride); seg2[:4] = part
        tot = self._allgather(seg2)[:, :4].sum(axis=0)            # fixed rank order on every rank
        return np.sqrt(tot)

    def _allgather(self, seg):
        t = torch.from_numpy(seg.copy())
        outs = [torch.empty_like(t) for _ in range(dist.get_world_size())]
        dist.all_gather(outs, t)
        return np.stack([o.numpy() for o in outs])


def worker(rank, world, port, case_args, n_iter, ret):
    import os
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import exaadmm_b200 as ea
    from exaadmm_b200.synthetic import synthetic_case
    from exaadmm_b200.partition import partition_buses, build_local_grids
    grid = ea.GridData.from_opfdata(synthetic_case(*case_args), tight_factor=0.99)
    part = partition_buses(grid, world)
    lg = build_local_grids(grid, part)[rank]
    par = Parameters(); par.verbose = 0
    sim = PartitionedOracle(lg, par, 4e2, 4e4)
    sim.m.admm_increment_outer(); sim.m.admm_increment_reset_inner()
    res = [sim.inner_iteration() for _ in range(n_iter)]
    out = {k: sim.m.vec(k).copy() for k in ("u_curr", "v_curr", "z_curr", "l_curr")}
    out["res"] = np.array(res)
    ret[rank] = out
    dist.destroy_process_group()
