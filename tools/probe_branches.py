"""Latency probe of the two branch drivers on the hard-branch fixture (tests/golden/hard_branches.npz), on the GPU:
one problem per lane vs one per warp (a lone lane). Prints SM cycles per TRON solve (AL iteration) of
the long chains and checks the drivers against each other and against the host build of the same code."""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from exaadmm_b200 import capi  # noqa: E402


def solve(lib, per_warp, prob, meta):
    n = prob.shape[0]
    sol = np.zeros((n, 13)); work = np.zeros((n, 6), dtype=np.int32); cyc = np.zeros(n, dtype=np.int64)
    ms = C.c_double(0)
    p = np.ascontiguousarray(prob)
    rc = lib.ea_diag_branch_solve(0, per_warp, n, capi.dptr(p), int(meta[2]), float(meta[1]), float(meta[0]),
                                  capi.dptr(sol), work.ctypes.data_as(C.POINTER(C.c_int32)),
                                  cyc.ctypes.data_as(C.POINTER(C.c_int64)), C.byref(ms))
    if rc:
        raise RuntimeError(lib.ea_last_error(None).decode())
    return sol, work, cyc, ms.value


def main():
    lib = capi.load_library()
    d = np.load(ROOT / "tests" / "golden" / "hard_branches.npz")
    prob, sol0, work0, meta = d["prob"], d["sol"], d["work"], d["meta"]
    long_ = work0[:, 0] >= 15                      # the penalty-ladder chains
    res = {}
    for pw in (0, 1):
        sol, work, cyc, ms = solve(lib, pw, prob, meta)
        res[pw] = (sol, work)
        al = work[:, 0].astype(float)
        per_al = cyc[long_] / al[long_]
        per_ev = cyc / np.maximum(work[:, 1], 1)
        print(f"{'1 problem / warp  ' if pw else '32 problems / warp'}: kernel {ms * 1e3:8.1f} us | long chains "
              f"(n={long_.sum()}): cycles per AL iteration median {np.median(per_al):8.0f} min {per_al.min():8.0f} | "
              f"cycles per evaluation (all) median {np.median(per_ev):7.0f} | max cycles {cyc.max()}")
    a, wa = res[1]
    b, wb = res[0]
    print(f"per lane vs per warp: bitwise equal {np.array_equal(a, b) and np.array_equal(wa, wb)}")
    print(f"vs host build of the device code: max |x, F diff| {np.abs(a[:, :10] - sol0[:, :10]).max():.3e}; "
          f"evals equal {np.array_equal(wa[:, 1], work0[:, 1])} AL equal {np.array_equal(wa[:, 0], work0[:, 0])} "
          f"rejected equal {np.array_equal(wa[:, 4], work0[:, 4])}")


if __name__ == "__main__":
    main()
