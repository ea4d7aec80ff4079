"""Phase shares of the one-CTA team on the fixture systems (profile=1: cycles of thread 0 per phase)."""
import os, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); warnings.filterwarnings("ignore")
import cpkrylov_b200 as cp
from helpers import load_system, load_factors, EX_OPTS
for name, method, extra in (("cvxqp1_m", "cpminres", {}), ("cvxqp1_m", "cpcg", {}), ("cvxqp1_m", "cpminres", {"nitref": 0, "force_itref": False, "residual_update": False}),
                            ("cvxqp2_s", "cpgmres", {"restart": 100}), ("cvxqp2_s", "cpdqgmres", {"mem": 100})):
    s = load_system(name); fac = load_factors(name, "superlu")
    for rep in range(2):
        x, st, fl, S = cp.reg_cpkrylov(method, s["rhs"], s["Q"], s["A"], s["C"], s["G"], dict(EX_OPTS, profile=True, **extra), factors=fac, return_system=True)
        S.close()
    g = st["gpu"]; tot = sum(g["phase_cycles"].values()); it = max(st["niters"], 1)
    print(name, method, extra, "iters", st["niters"], "us/iter %.1f" % (1e3 * g["t_solve_ms"] / it),
          {k: "%.0f us (%.0f%%)" % (v / 1.965e3 / it, 100.0 * v / tot) for k, v in g["phase_cycles"].items() if v},
          "napply %d nldl %d nresid %d" % (g["napply"], g["nldlsolve"], g["nresid"]))
