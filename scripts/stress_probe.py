"""SURVEY 8d 'stress' variant at reduced scale: k=6 windowed B (deep, filled L) -- parity vs the
oracle and timing of both LDL' walks."""
import os, sys, time, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); warnings.filterwarnings("ignore")
import numpy as np
import cpkrylov_b200 as cp
from cpkrylov_b200 import synth
from cpkrylov_b200.ldl import ldl_superlu
from oracle import cpk_oracle as orc
g = int(sys.argv[1]) if len(sys.argv) > 1 else 40
w = synth.kkt_lap3d(g=g, k=6, window=64, seed_B=2)
t = time.time(); fac = ldl_superlu(synth.kp_matrix(w)); tf = time.time() - t
print("n", w["n"], "m", w["m"], "nnz(L)", fac[0].nnz, "factor %.1fs" % tf, flush=True)
o = dict(print=False)
xo, so, fo = orc.reg_cpkrylov("cpcg", w["rhs"], w["H"], w["B"], w["C"], w["G"], o, factor=lambda K: fac)
print("oracle iters", so["niters"], fo["solved"], "stime %.2fs" % so["stime"], flush=True)
for env in ({}, {"CPK_LDL_SYNCFREE": "1"}, {"CPK_LDL_SYNCFREE": "0"}, {}):
    os.environ.pop("CPK_LDL_SYNCFREE", None); os.environ.update(env)
    x, st, fl, S = cp.reg_cpkrylov("cpcg", w["rhs"], w["H"], w["B"], w["C"], w["G"], o, factors=fac, return_system=True)
    print(env, "gpu iters", st["niters"], fl["solved"], "ms %.2f" % st["gpu"]["t_solve_ms"], "rel diff vs oracle %.2e" % (np.linalg.norm(x - xo) / np.linalg.norm(xo)),
          "err vs x* %.2e" % (np.linalg.norm(x - w["xstar"]) / np.linalg.norm(w["xstar"])), S.M.info(), flush=True)
    S.close()
