"""Bus-graph partition of one ACOPF case over P GPUs (one process per GPU).

The reference is single-GPU (its MPI fields are vestigial), so this
is new design, following SURVEY.md §8(e):

* buses are partitioned by a graph cut; every entry of the ADMM vectors belongs to
  exactly one bus (generator -> its bus; branch entries pij,qij,wi,ti -> from-bus,
  pji,qji,wj,tj -> to-bus: the scatter pattern of ``acopf_bus_kernel_gpu.jl:96-114``),
  so entry ownership = bus ownership;
* a cut branch is solved redundantly on both ranks it touches (same inputs, same
  code, bit-identical result), each rank runs the bus update only for its own buses,
  and the only exchange per inner iteration is the xbar (``v``) half of every cut
  branch end, gathered together with the 4 residual partial sums of each rank in ONE
  all-gather; z and lambda of the remote ("ghost") halves are then updated
  redundantly from the received xbar.

Everything here is host-side set-up (numpy). ``LocalGrid`` carries the rank-local
``GridData`` (owned buses first, ghost buses after) and the exchange maps; every rank
derives all of it from the global grid and the partition vector, so set-up needs no
communication.
"""
from __future__ import annotations

from dataclasses import dataclass, field, replace

import numpy as np

from .grid_data import GridData, _csr_1based


# ---------------------------------------------------------------------------------------
# partitioner
# ---------------------------------------------------------------------------------------
def _adjacency(grid: GridData):
    import scipy.sparse as sp
    f = grid.brBusIdx[0::2] - 1
    t = grid.brBusIdx[1::2] - 1
    n = grid.nbus
    A = sp.coo_matrix((np.ones(2 * len(f)), (np.concatenate([f, t]), np.concatenate([t, f]))), shape=(n, n)).tocsr()
    A.data[:] = 1.0
    return A


def _bfs_order(A, nodes: np.ndarray) -> np.ndarray:
    """Breadth-first order of the sub-graph induced by `nodes`, started from a
    pseudo-peripheral node; disconnected pieces are appended."""
    import scipy.sparse.csgraph as cg
    sub = A[nodes][:, nodes]
    n = len(nodes)
    seen = np.zeros(n, dtype=bool)
    order = []
    while len(order) < n:
        start = int(np.flatnonzero(~seen)[0])
        for _ in range(2):                                # two sweeps -> pseudo-peripheral start
            o = cg.breadth_first_order(sub, start, directed=False, return_predecessors=False)
            start = int(o[-1])
        o = cg.breadth_first_order(sub, start, directed=False, return_predecessors=False)
        o = o[~seen[o]]
        seen[o] = True
        order.extend(o.tolist())
    return nodes[np.array(order, dtype=np.int64)]


def _refine(A, part: np.ndarray, weight: np.ndarray, a: int, b: int, tol: float, sweeps: int = 4):
    """Greedy boundary refinement between parts a and b: move a bus to the other side if
    that removes more cut edges than it adds and keeps the weight balance within tol."""
    indptr, indices = A.indptr, A.indices
    wa, wb = weight[part == a].sum(), weight[part == b].sum()
    target = (wa + wb) / 2.0
    for _ in range(sweeps):
        moved = 0
        cand = np.flatnonzero((part == a) | (part == b))
        for v in cand:
            nb = indices[indptr[v]:indptr[v + 1]]
            pv = part[v]
            other = b if pv == a else a
            same = int(np.sum(part[nb] == pv))
            oth = int(np.sum(part[nb] == other))
            if oth > same:
                w_from = wa if pv == a else wb
                w_to = wb if pv == a else wa
                if w_to + weight[v] <= target * (1 + tol) and w_from - weight[v] >= target * (1 - tol):
                    part[v] = other
                    if pv == a:
                        wa -= weight[v]; wb += weight[v]
                    else:
                        wb -= weight[v]; wa += weight[v]
                    moved += 1
        if moved == 0:
            break


def partition_buses(grid: GridData, nparts: int, balance_tol: float = 0.03) -> np.ndarray:
    """Recursive BFS bisection + greedy refinement. Bus weight = 1 + number of incident
    branch ends (the work the bus and its branches generate). Deterministic."""
    if nparts < 1:
        raise ValueError("nparts must be >= 1")
    n = grid.nbus
    part = np.zeros(n, dtype=np.int32)
    if nparts == 1:
        return part
    A = _adjacency(grid)
    deg = np.bincount(grid.brBusIdx - 1, minlength=n).astype(np.float64)
    weight = 1.0 + deg

    def split(nodes: np.ndarray, first: int, k: int):
        if k == 1:
            part[nodes] = first
            return
        k1 = k // 2
        order = _bfs_order(A, nodes)
        cw = np.cumsum(weight[order])
        cutpos = int(np.searchsorted(cw, cw[-1] * k1 / k)) + 1
        cutpos = min(max(cutpos, 1), len(order) - 1)
        left, right = order[:cutpos], order[cutpos:]
        part[left] = first
        part[right] = first + k1
        if k == 2:
            _refine(A, part, weight, first, first + k1, balance_tol)
            left = np.flatnonzero(part == first)
            right = np.flatnonzero(part == first + k1)
            left = left[np.isin(left, nodes)]
            right = right[np.isin(right, nodes)]
        split(np.sort(left), first, k1)
        split(np.sort(right), first + k1, k - k1)

    split(np.arange(n, dtype=np.int64), 0, nparts)
    return part


def cut_statistics(grid: GridData, part: np.ndarray) -> dict:
    f = part[grid.brBusIdx[0::2] - 1]
    t = part[grid.brBusIdx[1::2] - 1]
    cut = int(np.sum(f != t))
    sizes = np.bincount(part, minlength=int(part.max()) + 1)
    return {"cut_lines": cut, "cut_fraction": cut / max(grid.nline, 1), "bus_counts": sizes.tolist()}


# ---------------------------------------------------------------------------------------
# rank-local grids and exchange maps
# ---------------------------------------------------------------------------------------
FROM_POS = np.array([0, 1, 4, 6])      # positions of (p, q, w, theta) of the from-end in the 8-entry branch record
TO_POS = np.array([2, 3, 5, 7])


@dataclass
class LocalGrid:
    rank: int
    nparts: int
    grid: GridData                     # local: owned buses first, ghost buses after
    n_owned_bus: int
    bus_global: np.ndarray             # local bus   -> global bus (0-based)
    line_global: np.ndarray            # local line  -> global line
    gen_global: np.ndarray             # local gen   -> global gen
    owned_entry: np.ndarray            # bool[nvar_local]: entry owned by this rank (reference layout)
    entry_global: np.ndarray           # int64[nvar_local]: global reference-layout index of every local entry
    send_line: np.ndarray              # local line / end (0 from, 1 to) of owned ends that are ghosts elsewhere,
    send_end: np.ndarray               #   ordered by (global line, end): position = index in this list
    ghost_line: np.ndarray             # local line / end of the ends owned by another rank
    ghost_end: np.ndarray
    ghost_src_rank: np.ndarray         # owner rank of each ghost end
    ghost_src_pos: np.ndarray          # its position in the owner's send list
    send_counts: np.ndarray            # send-list length of every rank
    stats: dict = field(default_factory=dict)

    @property
    def max_send(self) -> int:
        return int(self.send_counts.max()) if len(self.send_counts) else 0

    def half_positions(self, line: np.ndarray, end: np.ndarray) -> np.ndarray:
        """(n, 4) reference-layout indices of the (p, q, w, theta) entries of branch ends."""
        base = 2 * self.grid.ngen + 8 * np.asarray(line, dtype=np.int64)
        pos = np.where(np.asarray(end)[:, None] == 0, FROM_POS[None, :], TO_POS[None, :])
        return base[:, None] + pos


def _subgrid(g: GridData, bus_ids: np.ndarray, line_ids: np.ndarray, gen_ids: np.ndarray) -> GridData:
    """Restrict a GridData to the given buses / branches / generators (given in the local
    order wanted), renumbering buses; relative order inside every bus list is preserved."""
    nb = len(bus_ids)
    gmap = -np.ones(g.nbus, dtype=np.int64)
    gmap[bus_ids] = np.arange(nb)
    fr = gmap[g.brBusIdx[0::2][line_ids] - 1]
    to = gmap[g.brBusIdx[1::2][line_ids] - 1]
    assert np.all(fr >= 0) and np.all(to >= 0)
    two = np.empty(2 * len(line_ids), dtype=np.int64)
    two[0::2] = 2 * line_ids
    two[1::2] = 2 * line_ids + 1
    # generator -> bus from the global CSR
    gen_bus_global = _gen_bus(g)
    gb = gmap[gen_bus_global[gen_ids]] if len(gen_ids) else np.zeros(0, dtype=np.int64)
    fr_start, fr_idx = _csr_1based(fr, nb)
    to_start, to_idx = _csr_1based(to, nb)
    gen_start, gen_idx = _csr_1based(gb, nb)
    br = np.empty(2 * len(line_ids), dtype=np.int64)
    br[0::2] = fr + 1
    br[1::2] = to + 1
    pick_l = lambda a: np.ascontiguousarray(a[line_ids])
    pick_2l = lambda a: np.ascontiguousarray(a[two])
    pick_g = lambda a: np.ascontiguousarray(a[gen_ids])
    pick_b = lambda a: np.ascontiguousarray(a[bus_ids])
    return GridData(
        baseMVA=g.baseMVA, ngen=len(gen_ids), nline=len(line_ids), nbus=nb,
        pgmin=pick_g(g.pgmin), pgmax=pick_g(g.pgmax), qgmin=pick_g(g.qgmin), qgmax=pick_g(g.qgmax),
        c2=pick_g(g.c2), c1=pick_g(g.c1), c0=pick_g(g.c0),
        YshR=pick_b(g.YshR), YshI=pick_b(g.YshI),
        YffR=pick_l(g.YffR), YffI=pick_l(g.YffI), YftR=pick_l(g.YftR), YftI=pick_l(g.YftI),
        YttR=pick_l(g.YttR), YttI=pick_l(g.YttI), YtfR=pick_l(g.YtfR), YtfI=pick_l(g.YtfI),
        FrVmBound=pick_2l(g.FrVmBound), ToVmBound=pick_2l(g.ToVmBound),
        FrVaBound=pick_2l(g.FrVaBound), ToVaBound=pick_2l(g.ToVaBound), rateA=pick_l(g.rateA),
        FrStart=fr_start, FrIdx=fr_idx, ToStart=to_start, ToIdx=to_idx, GenStart=gen_start, GenIdx=gen_idx,
        Pd=pick_b(g.Pd), Qd=pick_b(g.Qd), Vmin=pick_b(g.Vmin), Vmax=pick_b(g.Vmax), brBusIdx=br,
    )


def _gen_bus(g: GridData) -> np.ndarray:
    """bus (0-based) of every generator, from the bus -> generator CSR."""
    gen_bus = np.empty(g.ngen, dtype=np.int64)
    gen_bus[np.asarray(g.GenIdx, dtype=np.int64) - 1] = np.repeat(np.arange(g.nbus, dtype=np.int64), np.diff(g.GenStart))
    return gen_bus


def build_local_grids(grid: GridData, part: np.ndarray, only: int | None = None) -> list[LocalGrid]:
    """The ranks' local grids. `only`: build just that rank's (what a rank of a multi-GPU job needs; the list then
    holds None for the others); building all lets the tests check global consistency."""
    nparts = int(part.max()) + 1
    fbus = grid.brBusIdx[0::2] - 1
    tbus = grid.brBusIdx[1::2] - 1
    fpart, tpart = part[fbus], part[tbus]
    gen_bus = _gen_bus(grid)
    cut = np.flatnonzero(fpart != tpart)
    # send list of rank r: owned ends of cut branches, ordered by (global line, end); key = 2 * line + end
    cut_keys = np.concatenate([2 * cut, 2 * cut + 1])
    cut_owner = np.concatenate([fpart[cut], tpart[cut]])
    order = np.argsort(cut_keys, kind="stable")
    cut_keys, cut_owner = cut_keys[order], cut_owner[order]
    send_keys = [cut_keys[cut_owner == r] for r in range(nparts)]
    send_counts = np.array([len(k) for k in send_keys], dtype=np.int64)
    # position of every cut-branch end in its owner's send list
    pos_of_key = np.empty(len(cut_keys), dtype=np.int64)
    for r in range(nparts):
        pos_of_key[cut_owner == r] = np.arange(send_counts[r])

    out = []
    for r in range(nparts):
        if only is not None and r != only:
            out.append(None)
            continue
        owned = np.flatnonzero(part == r)
        lines = np.flatnonzero((fpart == r) | (tpart == r))
        ghosts = np.unique(np.concatenate([tbus[lines][tpart[lines] != r], fbus[lines][fpart[lines] != r]])) \
            if len(lines) else np.zeros(0, dtype=np.int64)
        buses = np.concatenate([owned, ghosts]).astype(np.int64)
        gens = np.flatnonzero(part[gen_bus] == r)
        lg = _subgrid(grid, buses, lines, gens)
        nvar = 2 * len(gens) + 8 * len(lines)
        owned_entry = np.zeros(nvar, dtype=bool)
        owned_entry[: 2 * len(gens)] = True
        entry_global = np.empty(nvar, dtype=np.int64)
        entry_global[0:2 * len(gens):2] = 2 * gens
        entry_global[1:2 * len(gens):2] = 2 * gens + 1
        base_l = 2 * len(gens) + 8 * np.arange(len(lines))
        base_g = 2 * grid.ngen + 8 * lines
        for k in range(8):
            entry_global[base_l + k] = base_g + k
        f_own = fpart[lines] == r
        t_own = tpart[lines] == r
        for k in FROM_POS:
            owned_entry[base_l[f_own] + k] = True
        for k in TO_POS:
            owned_entry[base_l[t_own] + k] = True
        # local index of a global line (lines is sorted)
        sk = send_keys[r]
        s_line = np.searchsorted(lines, sk // 2).astype(np.int64)
        s_end = (sk % 2).astype(np.int64)
        # ghost ends, ordered by (local line, end): ends of this rank's lines owned by another rank
        loc = np.arange(len(lines), dtype=np.int64)
        gk_local = np.concatenate([2 * loc[~f_own], 2 * loc[~t_own] + 1])
        gk_global = np.concatenate([2 * lines[~f_own], 2 * lines[~t_own] + 1])
        gk_rank = np.concatenate([fpart[lines][~f_own], tpart[lines][~t_own]])
        o = np.argsort(gk_local, kind="stable")
        gk_local, gk_global, gk_rank = gk_local[o], gk_global[o], gk_rank[o]
        g_pos = pos_of_key[np.searchsorted(cut_keys, gk_global)] if len(gk_global) else np.zeros(0, dtype=np.int64)
        out.append(LocalGrid(
            rank=r, nparts=nparts, grid=lg, n_owned_bus=len(owned), bus_global=buses, line_global=lines,
            gen_global=gens, owned_entry=owned_entry, entry_global=entry_global,
            send_line=s_line, send_end=s_end,
            ghost_line=(gk_local // 2).astype(np.int64), ghost_end=(gk_local % 2).astype(np.int64),
            ghost_src_rank=gk_rank.astype(np.int64), ghost_src_pos=g_pos.astype(np.int64),
            send_counts=send_counts,
            stats={"owned_buses": len(owned), "ghost_buses": len(ghosts), "lines": len(lines),
                   "cut_lines": int(np.sum(fpart[lines] != tpart[lines])), "gens": len(gens)}))
    return out


def assemble_global(locals_: list[LocalGrid], vectors: list[np.ndarray], nvar_global: int) -> np.ndarray:
    """Global reference-layout vector from every rank's local vector (owner's copy wins)."""
    out = np.zeros(nvar_global)
    for lg, v in zip(locals_, vectors):
        out[lg.entry_global[lg.owned_entry]] = v[lg.owned_entry]
    return out
