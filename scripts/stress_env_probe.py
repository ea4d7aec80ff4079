"""Stress system (k=6 windowed B, filled factor): cpcg solve time under the set-up knobs given in the environment.
usage: [CPK_LDL_...=..] python scripts/stress_env_probe.py [g] [label]"""
import os, sys, time, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); warnings.filterwarnings("ignore")
import numpy as np
import cpkrylov_b200 as cp
from cpkrylov_b200 import synth
from cpkrylov_b200.ldl import ldl_superlu
g = int(sys.argv[1]) if len(sys.argv) > 1 else 40
label = sys.argv[2] if len(sys.argv) > 2 else ""
w = synth.kkt_lap3d(g=g, k=6, window=64, seed_B=2)
fac = ldl_superlu(synth.kp_matrix(w))
o = dict(print=False)
ts = []
t0 = time.time()
for _ in range(3):
    x, st, fl = cp.reg_cpkrylov("cpcg", w["rhs"], w["H"], w["B"], w["C"], w["G"], o, factors=fac)
    ts.append(st["gpu"]["t_solve_ms"])
print(label, "g", g, "iters", st["niters"], fl["solved"], "ms min %.2f" % min(ts), "s per call %.1f" % ((time.time() - t0) / 3),
      "err vs x* %.2e" % (np.linalg.norm(x - w["xstar"]) / np.linalg.norm(w["xstar"])), flush=True)
