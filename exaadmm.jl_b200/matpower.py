"""MATPOWER ``.m`` case reader (host side, set-up only).

Behaviour follows the reference's loader so that the same file produces the
same flat arrays:

* ``/root/reference/src/utils/parse_matpower.jl:1-241``  (keys ``mpc.baseMVA``,
  ``mpc.bus``, ``mpc.gen``, ``mpc.branch``, ``mpc.gencost``; rows with
  ``status != 1`` are dropped for generators (``:129``) and branches (``:156``);
  generator P/Q limits are divided by ``baseMVA`` (``:139-141``); the k-th
  ``gencost`` row belongs to the k-th ``gen`` row (``:145-147``)).
* ``/root/reference/src/utils/opfdata.jl:132-254`` (exactly one reference bus
  (``:158-164``); ``±Inf`` limits become ``±999.99`` *after* the p.u. scaling
  (``:199-206``); only polynomial cost with three coefficients (``:220``)).

The tokenizer here is table-driven (numpy matrices per block) rather than the
reference's dictionary-of-rows; it also accepts several ``;``-separated rows on
one text line, which the reference does not need.
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

BUS_COLS = ("bus_i", "type", "Pd", "Qd", "Gs", "Bs", "area", "Vm", "Va",
            "baseKV", "zone", "Vmax", "Vmin")
GEN_COLS = ("bus", "Pg", "Qg", "Qmax", "Qmin", "Vg", "mBase", "status",
            "Pmax", "Pmin")
BRANCH_COLS = ("fbus", "tbus", "r", "x", "b", "rateA", "rateB", "rateC",
               "ratio", "angle", "status", "angmin", "angmax")
GENCOST_COLS = ("cost_type", "startup", "shutdown", "n", "c2", "c1", "c0")

_KEYS = ("mpc.bus", "mpc.gen", "mpc.branch", "mpc.gencost")


class MatpowerFormatError(ValueError):
    pass


def _to_float(tok: str) -> float:
    t = tok.strip()
    low = t.lower()
    if low in ("inf", "+inf"):
        return float("inf")
    if low == "-inf":
        return float("-inf")
    return float(t)


def _read_blocks(text: str):
    """Return (baseMVA, {key: 2-D float array})."""
    base = None
    blocks: dict[str, np.ndarray] = {}
    lines = text.splitlines()
    i = 0
    while i < len(lines):
        raw = lines[i]
        code = raw.split("%", 1)[0].strip()
        i += 1
        if not code:
            continue
        m = re.match(r"^(mpc\.\w+)\s*=\s*(.*)$", code)
        if not m:
            continue
        key, rest = m.group(1), m.group(2)
        if key == "mpc.baseMVA":
            base = _to_float(rest.replace(";", ""))
            continue
        if key not in _KEYS:
            # unknown block (mpc.version, mpc.areas, bus_name ...): skip to its end
            if "[" in rest and "]" not in rest:
                while i < len(lines) and "]" not in lines[i].split("%", 1)[0]:
                    i += 1
                i += 1
            continue
        body = rest.split("[", 1)[1] if "[" in rest else ""
        closed = "]" in body
        body = body.split("]", 1)[0]
        while not closed:
            if i >= len(lines):
                raise MatpowerFormatError(f"unterminated block {key}")
            seg = lines[i].split("%", 1)[0]
            i += 1
            if "]" in seg:
                closed = True
                seg = seg.split("]", 1)[0]
            body += "\n" + seg
        rows = []
        for row in re.split(r"[;\n]", body):
            toks = row.replace(",", " ").split()
            if toks:
                rows.append([_to_float(t) for t in toks])
        if rows:
            width = min(len(r) for r in rows)
            blocks[key] = np.array([r[:width] for r in rows], dtype=np.float64)
        else:
            blocks[key] = np.zeros((0, 0))
    if base is None:
        raise MatpowerFormatError("mpc.baseMVA not found")
    for k in _KEYS:
        if k not in blocks:
            raise MatpowerFormatError(f"{k} not found")
    return base, blocks


@dataclass
class OPFData:
    """Flat (struct-of-arrays) form of the reference's ``OPFData``
    (``/root/reference/src/utils/opfdata.jl:71-84``). All indices 0-based."""
    case: str
    baseMVA: float
    # buses (file order)
    bus_i: np.ndarray
    bus_type: np.ndarray
    Pd: np.ndarray            # MW   (NOT p.u.; divided by baseMVA inside the bus update)
    Qd: np.ndarray            # MVAr
    Gs: np.ndarray
    Bs: np.ndarray
    Vmax: np.ndarray
    Vmin: np.ndarray
    bus_ref: int              # 0-based index of the single reference bus
    # active branches (file order)
    br_from: np.ndarray       # 0-based bus index
    br_to: np.ndarray
    br_r: np.ndarray
    br_x: np.ndarray
    br_b: np.ndarray
    br_rateA: np.ndarray      # MVA, raw
    br_ratio: np.ndarray
    br_angle: np.ndarray      # degrees
    # active generators (file order)
    gen_bus: np.ndarray       # 0-based bus index
    Pmax: np.ndarray          # p.u.
    Pmin: np.ndarray
    Qmax: np.ndarray
    Qmin: np.ndarray
    coeff: np.ndarray         # (ngen, 3): c2, c1, c0 (unscaled, $/MW^2h ...)
    n_gen_total: int = 0
    n_branch_total: int = 0
    extras: dict = field(default_factory=dict)

    @property
    def nbus(self) -> int:
        return int(self.bus_i.shape[0])

    @property
    def nline(self) -> int:
        return int(self.br_from.shape[0])

    @property
    def ngen(self) -> int:
        return int(self.gen_bus.shape[0])


def parse_matpower_text(text: str, case: str = "case") -> OPFData:
    base, blk = _read_blocks(text)
    bus = blk["mpc.bus"]
    gen = blk["mpc.gen"]
    br = blk["mpc.branch"]
    gc = blk["mpc.gencost"]
    if bus.shape[1] < len(BUS_COLS):
        raise MatpowerFormatError("mpc.bus needs 13 columns")
    if gen.shape[1] < len(GEN_COLS):
        raise MatpowerFormatError("mpc.gen needs at least 10 columns")
    if br.shape[1] < 11:
        raise MatpowerFormatError("mpc.branch needs at least 11 columns")
    if gc.shape[0] < gen.shape[0] or gc.shape[1] < len(GENCOST_COLS):
        raise MatpowerFormatError("mpc.gencost: need one 7-column row per generator")

    bus_i = bus[:, 0].astype(np.int64)
    if np.any(bus_i <= 0):
        raise MatpowerFormatError("bus ids must be positive")          # opfdata.jl:143
    if len(np.unique(bus_i)) != len(bus_i):
        raise MatpowerFormatError("duplicate bus id")                   # opfdata.jl:514
    bus_type = bus[:, 1].astype(np.int64)
    refs = np.flatnonzero(bus_type == 3)
    if len(refs) == 0:
        raise MatpowerFormatError("reference bus was not found")       # parse_matpower.jl:99-102
    if len(refs) > 1:
        raise MatpowerFormatError("More than one reference bus present in the data")  # opfdata.jl:158-160
    id2idx = {int(b): k for k, b in enumerate(bus_i)}

    g_on = gen[:, 7].astype(np.int64) == 1
    l_on = br[:, 10].astype(np.int64) == 1
    gen_a = gen[g_on]
    gc_a = gc[: gen.shape[0]][g_on]
    br_a = br[l_on]
    if np.any(gc_a[:, 0].astype(np.int64) != 2) or np.any(gc_a[:, 3].astype(np.int64) != 3):
        raise MatpowerFormatError("only polynomial (type 2) cost with n = 3 is supported")  # opfdata.jl:220

    def lim(col, sign):
        v = gen_a[:, col] / base
        return np.where(np.isinf(v), sign * 999.99, v)

    try:
        gen_bus = np.array([id2idx[int(b)] for b in gen_a[:, 0]], dtype=np.int64)
        br_from = np.array([id2idx[int(b)] for b in br_a[:, 0]], dtype=np.int64)
        br_to = np.array([id2idx[int(b)] for b in br_a[:, 1]], dtype=np.int64)
    except KeyError as e:
        raise MatpowerFormatError(f"unknown bus id {e.args[0]}") from None

    ncol = br_a.shape[1]
    return OPFData(
        case=case, baseMVA=float(base),
        bus_i=bus_i, bus_type=bus_type,
        Pd=bus[:, 2].copy(), Qd=bus[:, 3].copy(), Gs=bus[:, 4].copy(), Bs=bus[:, 5].copy(),
        Vmax=bus[:, 11].copy(), Vmin=bus[:, 12].copy(), bus_ref=int(refs[0]),
        br_from=br_from, br_to=br_to,
        br_r=br_a[:, 2].copy(), br_x=br_a[:, 3].copy(), br_b=br_a[:, 4].copy(),
        br_rateA=br_a[:, 5].copy(),
        br_ratio=br_a[:, 8].copy() if ncol > 8 else np.zeros(len(br_a)),
        br_angle=br_a[:, 9].copy() if ncol > 9 else np.zeros(len(br_a)),
        gen_bus=gen_bus,
        Pmax=lim(8, +1.0), Pmin=lim(9, -1.0), Qmax=lim(3, +1.0), Qmin=lim(4, -1.0),
        coeff=gc_a[:, 4:7].copy(),
        n_gen_total=int(gen.shape[0]), n_branch_total=int(br.shape[0]),
    )


def parse_matpower(filename, verbose: int = 0) -> OPFData:
    """``opf_loaddata(case; case_format="matpower")`` of the reference
    (``opfdata.jl:408-415``)."""
    p = Path(filename)
    data = parse_matpower_text(p.read_text(), case=p.name.split(".")[0])
    if verbose >= 1:  # parse_matpower.jl:231-237
        print(f" ** Statistics of {data.case}")
        print(f"  # buses     : {data.nbus:5d}")
        print(f"  # generators: {data.n_gen_total:5d} ({data.ngen:5d} active)")
        print(f"  # branches  : {data.n_branch_total:5d} ({data.nline:5d} active)")
        print(f"  # gencost   : {data.n_gen_total:5d} ({data.ngen:5d} active)")
    return data


def write_matpower(path, *, baseMVA, bus, gen, branch, gencost, name="case") -> None:
    """Write a case in the one-row-per-line form the reference parser needs
    (``parse_matpower.jl:1-16,60-63``)."""
    def block(key, mat):
        out = [f"{key} = ["]
        for row in mat:
            out.append("\t" + "\t".join(repr(float(x)) if not float(x).is_integer()
                                        else str(int(x)) for x in row) + ";")
        out.append("];")
        return "\n".join(out)

    txt = [f"function mpc = {name}", "mpc.version = '2';", f"mpc.baseMVA = {baseMVA:g};",
           block("mpc.bus", bus), block("mpc.gen", gen), block("mpc.branch", branch),
           block("mpc.gencost", gencost), ""]
    Path(path).write_text("\n".join(txt))
