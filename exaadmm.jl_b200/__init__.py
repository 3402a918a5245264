"""B200-native two-level ADMM for ACOPF — drop-in for the hot path of
ExaAdmm.jl's ``solve_acopf(...; use_gpu=true)``.

Host side (this package) mirrors the reference's operator API; the compute path
is the CUDA library behind ``include/exaadmm_b200.h`` (``csrc/``).
"""
from .matpower import OPFData, parse_matpower, parse_matpower_text, write_matpower, MatpowerFormatError
from .grid_data import GridData

from .environment import AdmmEnv, Parameters, IterationInformation, ComponentInformation  # noqa: E402
from .model import ModelAcopf, Solution  # noqa: E402
from .solve_acopf import solve_acopf  # noqa: E402
from .admm_two_level import admm_two_level, print_statistics  # noqa: E402
from .mpacopf import ModelMpacopf, SolutionRamping, solve_mpacopf  # noqa: E402
from .qpsub import ModelQpsub, solve_qpsub, admm_one_level  # noqa: E402

__all__ = ["AdmmEnv", "Parameters", "IterationInformation", "ComponentInformation", "ModelAcopf", "Solution",
           "solve_acopf", "solve_mpacopf", "solve_qpsub", "ModelQpsub", "admm_one_level", "ModelMpacopf", "SolutionRamping", "admm_two_level", "print_statistics", "OPFData", "parse_matpower", "parse_matpower_text", "write_matpower",
           "MatpowerFormatError", "GridData"]

CASE9 = str(__import__("pathlib").Path(__file__).resolve().parent / "data" / "case9.m")
