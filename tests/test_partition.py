"""Multi-GPU host logic on CPU: bus-graph partition, rank-local grids, ownership and
exchange maps; and a world_size-2 gloo run of the partitioned iteration (CPU oracle as
the per-rank compute engine) against the single-domain oracle."""
import numpy as np
import pytest

import exaadmm_b200 as ea
from exaadmm_b200.environment import Parameters
from exaadmm_b200.partition import (assemble_global, build_local_grids, cut_statistics, partition_buses)
from exaadmm_b200.synthetic import synthetic_case
from oracle.oracle import OracleModel

CASE = (400, 60, 560, 400)


@pytest.fixture(scope="module")
def grid():
    return ea.GridData.from_opfdata(synthetic_case(*CASE), tight_factor=0.99)


@pytest.mark.parametrize("nparts", [1, 2, 3, 4, 8])
def test_partition_is_balanced_and_cut_is_small(grid, nparts):
    part = partition_buses(grid, nparts)
    assert part.shape == (grid.nbus,) and part.min() == 0 and part.max() == nparts - 1
    st = cut_statistics(grid, part)
    sizes = np.array(st["bus_counts"])
    assert sizes.min() > 0.6 * grid.nbus / nparts and sizes.max() < 1.4 * grid.nbus / nparts
    if nparts > 1:
        assert 0 < st["cut_fraction"] < (0.15 if nparts == 2 else 0.6)      # tiny grid: 50-bus parts
    np.testing.assert_array_equal(part, partition_buses(grid, nparts))          # deterministic


@pytest.mark.parametrize("nparts", [2, 4])
def test_local_grids_cover_everything_exactly_once(grid, nparts):
    part = partition_buses(grid, nparts)
    lgs = build_local_grids(grid, part)
    nvar = 2 * grid.ngen + 8 * grid.nline
    owner_count = np.zeros(nvar, dtype=int)
    for lg in lgs:
        assert lg.grid.nbus == lg.n_owned_bus + lg.stats["ghost_buses"]
        np.testing.assert_array_equal(part[lg.bus_global[: lg.n_owned_bus]], lg.rank)
        assert np.all(part[lg.bus_global[lg.n_owned_bus:]] != lg.rank)
        owner_count[lg.entry_global[lg.owned_entry]] += 1
        # every ghost end finds its value at (src_rank, src_pos) of that rank's send list
        for k in range(len(lg.ghost_line)):
            src = lgs[lg.ghost_src_rank[k]]
            pos = lg.ghost_src_pos[k]
            assert src.line_global[src.send_line[pos]] == lg.line_global[lg.ghost_line[k]]
            assert src.send_end[pos] == lg.ghost_end[k]
        # branch data travelled with the branch
        np.testing.assert_array_equal(lg.grid.YffR, grid.YffR[lg.line_global])
        np.testing.assert_array_equal(lg.grid.FrVaBound[0::2], grid.FrVaBound[0::2][lg.line_global])
        np.testing.assert_array_equal(lg.grid.Pd[: lg.n_owned_bus], grid.Pd[lg.bus_global[: lg.n_owned_bus]])
    np.testing.assert_array_equal(owner_count, 1)                                 # a partition of the entries
    assert sum(lg.stats["gens"] for lg in lgs) == grid.ngen
    cut = cut_statistics(grid, part)["cut_lines"]
    assert sum(lg.stats["lines"] for lg in lgs) == grid.nline + cut               # cut branches are duplicated once


def test_two_rank_gloo_run_matches_single_domain_oracle(grid):
    import torch.multiprocessing as mp
    from partition_sim import worker
    n_iter = 12
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(worker, args=(world, port, CASE, n_iter, ret), nprocs=world, join=True)

    par = Parameters(); par.verbose = 0
    ref = OracleModel(grid, par, 4e2, 4e4)
    ref.admm_increment_outer(); ref.admm_increment_reset_inner()
    ref_res = np.array([ref.inner_iteration() for _ in range(n_iter)])
    part = partition_buses(grid, world)
    lgs = build_local_grids(grid, part)
    nvar = 2 * grid.ngen + 8 * grid.nline
    for name in ("u_curr", "v_curr", "z_curr", "l_curr"):
        glob = assemble_global(lgs, [ret[r][name] for r in range(world)], nvar)
        np.testing.assert_array_equal(glob, ref.vec(name))                        # same arithmetic, same order: bitwise
    for r in range(world):
        np.testing.assert_allclose(ret[r]["res"], ref_res, rtol=1e-13)            # norms: different summation order
        # ghost copies agree with the owner's values
        lg = lgs[r]
        glob_v = assemble_global(lgs, [ret[q]["v_curr"] for q in range(world)], nvar)
        np.testing.assert_array_equal(ret[r]["v_curr"], glob_v[lg.entry_global])
