"""Import shim: the package directory is ``exaadmm.jl_b200/`` (the name the
project layout prescribes), which is not a valid Python identifier. Importing
``exaadmm_b200`` loads that directory as a regular package under this name."""
import importlib.util as _u
import sys as _s
from pathlib import Path as _P

_d = _P(__file__).resolve().parent / "exaadmm.jl_b200"
_spec = _u.spec_from_file_location("exaadmm_b200", _d / "__init__.py", submodule_search_locations=[str(_d)])
_m = _u.module_from_spec(_spec)
_s.modules["exaadmm_b200"] = _m
_spec.loader.exec_module(_m)
