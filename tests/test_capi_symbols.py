"""The C-ABI library loads on a CPU-only box, exports every symbol
include/exaadmm_b200.h declares, and refuses to compute without a GPU."""
import ctypes as C
import re
from pathlib import Path

import pytest

import exaadmm_b200 as ea
from exaadmm_b200 import capi

ROOT = Path(__file__).resolve().parent.parent


def _declared():
    txt = (ROOT / "include" / "exaadmm_b200.h").read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ea_[a-z_0-9]+)\s*\(", txt)))


def test_header_and_binding_agree():
    assert _declared() == sorted(capi.SIGNATURES)


def test_library_exports_every_declared_symbol():
    lib = capi.load_library()
    for name in _declared():
        assert hasattr(lib, name), name
    assert lib.ea_abi_version() == 1


def test_create_fails_loudly_without_gpu(case9_grid):
    lib = capi.load_library()
    if lib.ea_device_count() > 0:
        pytest.skip("GPU present")
    gs, keep = capi.make_grid_struct(case9_grid)
    h = C.c_void_p()
    rc = lib.ea_create(C.byref(gs), 0, C.byref(h))
    assert rc == capi.EA_ERR_CUDA and not h.value
    assert b"no CUDA device" in lib.ea_last_error(None)


def test_solve_acopf_has_no_cpu_path():
    from exaadmm_b200.solve_acopf import solve_acopf
    with pytest.raises(NotImplementedError, match="no CPU fallback"):
        solve_acopf(ea.CASE9, use_gpu=False, verbose=0)
    with pytest.raises(NotImplementedError):
        solve_acopf(ea.CASE9, use_gpu=True, ka_device="CUDABackend", verbose=0)
    with pytest.raises(TypeError):                       # rho typed ::Float64 in the reference
        solve_acopf(ea.CASE9, use_gpu=True, rho_pq=400, verbose=0)
