import json
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import exaadmm_b200 as ea  # noqa: E402
from exaadmm_b200.environment import Parameters  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu():
    try:
        from exaadmm_b200 import capi
        return capi.load_library().ea_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    return json.loads((ROOT / "tests" / "golden" / "case9_reference_golden.json").read_text())


@pytest.fixture(scope="session")
def case9_grid():
    return ea.GridData.from_opfdata(ea.parse_matpower(ea.CASE9))


@pytest.fixture()
def golden_params():
    p = Parameters()
    p.scale = 1e-4
    p.initial_beta = 1e3
    p.beta = 1e3
    p.verbose = 0
    return p


def _load_harness(name):
    import ctypes as C
    hdir = ROOT / "tests" / "harness"
    subprocess.run(["make", "-C", str(hdir)], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    H = C.CDLL(str(hdir / name))
    pd = C.POINTER(C.c_double)
    H.hh_solve_branch.argtypes = [pd, pd, pd, pd, pd, C.c_longlong, C.c_int, C.c_double, C.c_double, pd,
                                  C.POINTER(C.c_int)]
    H.hh_solve_branch_oracle_eval.argtypes = [C.c_void_p, C.c_void_p, pd, pd, pd, pd, pd, C.c_longlong, C.c_int,
                                              C.c_double, C.c_double, C.POINTER(C.c_int)]
    H.hh_eval.argtypes = [pd, pd, pd, C.c_double, pd, pd, pd]
    H.hh_solve_gen.argtypes = [pd, pd, pd, pd, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int,
                               C.c_double, C.POINTER(C.c_int)]
    return H


@pytest.fixture(scope="session")
def host_harness():
    """The product's device routines compiled for the host (tests/harness), FMA as on the GPU."""
    return _load_harness("_build_host_harness.so")


@pytest.fixture(scope="session")
def host_harness_nofma():
    """Same, built with -DEA_NO_FMA: arithmetic bit-identical to the oracle's."""
    return _load_harness("_build_host_harness_nofma.so")


@pytest.fixture(scope="session")
def host_harness_parity():
    """Same, with the arithmetic of the library's PARITY build (-DEA_NO_FMA -DEA_PARITY): the product's own
    objective in the oracle's operation order."""
    return _load_harness("_build_host_harness_parity.so")


def branch_inputs(grid, u, v, z, l, rho, membuf, I):
    """Inputs of one branch sub-problem as the reference stages them
    (acopf_auglag_linelimit_kernel_cpu.jl:23-80)."""
    import math
    p = 2 * grid.ngen + 8 * I
    xl = np.array([grid.FrVmBound[2 * I], grid.ToVmBound[2 * I], grid.FrVaBound[2 * I], grid.ToVaBound[2 * I],
                   -grid.rateA[I], -grid.rateA[I]])
    xu = np.array([grid.FrVmBound[2 * I + 1], grid.ToVmBound[2 * I + 1], grid.FrVaBound[2 * I + 1],
                   grid.ToVaBound[2 * I + 1], 0.0, 0.0])
    x = np.array([math.sqrt(u[p + 4]), math.sqrt(u[p + 5]), u[p + 6], u[p + 7],
                  -(u[p] * u[p] + u[p + 1] * u[p + 1]), -(u[p + 2] * u[p + 2] + u[p + 3] * u[p + 3])])   # not ** 2: pow() may round differently
    x = np.minimum(xu, np.maximum(xl, x))
    param = np.zeros(31)
    param[0:8] = l[p:p + 8]
    param[8:16] = rho[p:p + 8]
    param[16:24] = v[p:p + 8] - z[p:p + 8]
    param[24:27] = membuf[24:27, I]
    param[28] = grid.rateA[I]
    Y = np.array([grid.YffR[I], grid.YffI[I], grid.YftR[I], grid.YftI[I], grid.YttR[I], grid.YttI[I],
                  grid.YtfR[I], grid.YtfI[I]])
    return x, xl, xu, param, Y


@pytest.fixture(scope="session")
def host_harness_literal():
    """FMA build with the direct step (tron::newton_step) switched off: the literal TRON algorithm on every step."""
    return _load_harness("_build_host_harness_literal.so")
