"""Quick GPU-vs-oracle comparison on the two reference systems (development aid)."""
import os, sys, time, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
warnings.filterwarnings("ignore")
import numpy as np
import cpkrylov_b200 as cp
from oracle import cpk_oracle as orc
from helpers import load_system, load_factors, EX_OPTS, kp_of, relerr

def run(name, kind, meth, extra, team):
    os.environ["CPK_TEAM"] = team
    s = load_system(name)
    fac = load_factors(name, kind)
    o = dict(EX_OPTS); o.update(extra)
    xo, so, fo = orc.reg_cpkrylov(meth, s["rhs"], s["Q"], s["A"], s["C"], s["G"], o, factor=lambda K: fac)
    t = time.time()
    try:
        xg, sg, fg = cp.reg_cpkrylov(meth, s["rhs"], s["Q"], s["A"], s["C"], s["G"], o, factors=fac)
    except Exception as e:
        print(name, kind, meth, extra, team, "GPU ERROR", type(e).__name__, e); return
    ho = so.get("residHistory", so.get("cgresidHistory")); hg = sg.get("residHistory", sg.get("cgresidHistory"))
    L = min(len(ho), len(hg))
    hd = np.max(np.abs(ho[:L] - hg[:L]) / np.maximum(np.abs(ho[:L]), 1e-300))
    print("%-9s %-7s %-11s %-16s %-4s it %3d/%3d solved %d/%d relx %.2e hist %.2e gpu %.2f ms (wall %.2fs)" % (
        name, kind, meth, extra, team, sg["niters"], so["niters"], fg["solved"], fo["solved"], relerr(xg, xo), hd,
        sg["gpu"]["t_solve_ms"], time.time() - t))

for team in ("cta", "grid"):
    for meth, extra in [("cpminres", {}), ("cpcg", {}), ("cpcglanczos", {}), ("cpsymmlq", {}), ("cpdqgmres", {"mem": 2}), ("cpgmres", {"restart": 50})]:
        run("cvxqp1_m", "superlu", meth, extra, team)
    for kind in ("superlu", "densebk"):
        for meth, extra in [("cpgmres", {"restart": 100}), ("cpdqgmres", {"mem": 100}), ("cpgmres", {"restart": 20}), ("cpdqgmres", {"mem": 10})]:
            run("cvxqp2_s", kind, meth, extra, team)
