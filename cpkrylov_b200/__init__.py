"""cpkrylov_b200 -- B200-native (sm_100a) drop-in for cpkrylov's per-iteration hot
path: the constraint-preconditioned Krylov loops cpcg / cpcglanczos / cpminres /
cpsymmlq / cpgmres / cpdqgmres on [H B'; B -C] and the opLDL2 preconditioner
apply, behind the reference's own interface (reg_cpkrylov driver signature,
Spot-style ``M*z`` operator contract).

Python is only the host language of this mirror (the reference's host is
MATLAB; see matlab/ and INTEGRATION.md for the MEX side).  All arithmetic of the
iteration happens in libcpk_b200.so on the GPU; there is no CPU fallback.
"""
from .operators import opLDL2, KktSystem                      # noqa: F401
from .solvers import (cpcg, cpcglanczos, cpminres, cpsymmlq, cpgmres, cpdqgmres,   # noqa: F401
                      reg_cpkrylov, reg_solve_on, SolverError, SOLVERS)
from ._lib import CpkError, CpkLibraryMissing, LIB_PATH     # noqa: F401
from . import ldl                                              # noqa: F401
from .matio import load_mat_system, system_from_K, solve_sequence, solve_ipm_sequence   # noqa: F401

__all__ = ["opLDL2", "KktSystem", "cpcg", "cpcglanczos", "cpminres", "cpsymmlq", "cpgmres",
           "cpdqgmres", "reg_cpkrylov", "reg_solve_on", "SolverError", "SOLVERS", "CpkError", "CpkLibraryMissing",
           "load_mat_system", "system_from_K", "solve_sequence", "solve_ipm_sequence"]
