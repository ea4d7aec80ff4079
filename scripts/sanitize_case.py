"""Tiny end-to-end case for compute-sanitizer (memcheck): the cvxqp2 example through
both team shapes and both LDL' walks."""
import os, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); warnings.filterwarnings("ignore")
import numpy as np
import cpkrylov_b200 as cp
from helpers import load_system, load_factors, EX_OPTS, relerr
s = load_system("cvxqp2_s")
for team in ("cta", "grid"):
    for sf in ("0", "1"):
        os.environ["CPK_TEAM"] = team; os.environ["CPK_LDL_SYNCFREE"] = sf
        for kind, meth, extra in (("densebk", "cpgmres", {"restart": 30}), ("superlu", "cpminres", {}), ("superlu", "cpcg", {})):
            x, st, fl = cp.reg_cpkrylov(meth, s["rhs"], s["Q"], s["A"], s["C"], s["G"], dict(EX_OPTS, itmax=40, **extra), factors=load_factors("cvxqp2_s", kind))
            print(team, sf, kind, meth, st["niters"], fl["solved"])
print("done")
