import os, sys, warnings, ctypes as ct
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); warnings.filterwarnings("ignore")
import numpy as np
import cpkrylov_b200 as cp
from cpkrylov_b200 import _lib
from helpers import load_system, load_factors, EX_OPTS
os.environ["CPK_TEAM"] = "cta"; os.environ["CPK_VERBOSE"] = "1"
s = load_system("cvxqp1_m"); fac = load_factors("cvxqp1_m", "superlu")
x, st, fl, S = cp.reg_cpkrylov("cpminres", s["rhs"], s["Q"], s["A"], s["C"], s["G"], dict(EX_OPTS, profile=True), factors=fac, return_system=True)
g = st["gpu"]; tot = sum(g["phase_cycles"].values())
print("cpminres cta: iters", st["niters"], "ms", g["t_solve_ms"], "us/iter", 1e3 * g["t_solve_ms"] / st["niters"], {k: round(v / tot, 3) for k, v in g["phase_cycles"].items()}, "cycles/iter", tot / st["niters"])
L = _lib.lib(); M = S.M
z = np.random.default_rng(0).standard_normal(s["N"]); y = np.empty(s["N"])
M.nitref = 0
for mode in (0, 2):
    _lib.check(L.cpk_ldl2_set_track_rnorm(M.handle, mode))
    ts = []
    for _ in range(5):
        stt = _lib.StatsStruct(); _lib.check(L.cpk_ldl2_apply(M.handle, z.ctypes.data, y.ctypes.data, 0, ct.byref(stt))); ts.append(stt.t_solve_ms * 1e3)
    print("apply nitref=0 mode", mode, "us", np.median(ts), "cycles", [stt.phase_cycles[i] for i in range(8)], M.info())
S.close()
