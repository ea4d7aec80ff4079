"""A few stand-alone opLDL2 applies on cvxqp1_m (one-CTA team); used under ncu."""
import os, sys, warnings, ctypes as ct
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); warnings.filterwarnings("ignore")
import numpy as np
import cpkrylov_b200 as cp
from cpkrylov_b200 import _lib
from helpers import load_system, load_factors
s = load_system("cvxqp1_m"); fac = load_factors("cvxqp1_m", "superlu")
M = cp.opLDL2(s["G"], s["A"], -s["C"], factors=fac)
M.nitref = 0
z = np.random.default_rng(0).standard_normal(s["N"]); y = np.empty(s["N"])
L = _lib.lib()
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    stt = _lib.StatsStruct(); _lib.check(L.cpk_ldl2_apply(M.handle, z.ctypes.data, y.ctypes.data, 0, ct.byref(stt)))
print("apply us", stt.t_solve_ms * 1e3)
