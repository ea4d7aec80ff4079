"""Where the set-up time of one small system goes (cfg 5 pattern): factorization on the host, opLDL2
creation (host symbolic work + uploads), KKT system creation."""
import os, sys, time, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); warnings.filterwarnings("ignore")
import numpy as np
import cpkrylov_b200 as cp
from cpkrylov_b200 import synth
from cpkrylov_b200.ldl import ldl_superlu
from cpkrylov_b200.operators import KktSystem, opLDL2
base = synth.load_cvxqp1()
T = {"factor": [], "opLDL2": [], "upload_part": [], "system": []}
objs = []
for j in range(12):
    w = synth.ipm_batch_system(base, j)
    t0 = time.perf_counter(); fac = ldl_superlu(synth.kp_matrix(w)); t1 = time.perf_counter()
    M = opLDL2(w["G"], w["B"], -w["C"], factors=fac); t2 = time.perf_counter()
    S = KktSystem(w["H"], w["C"], M); t3 = time.perf_counter()
    T["factor"].append(t1 - t0); T["opLDL2"].append(t2 - t1); T["upload_part"].append(M.t_upload); T["system"].append(t3 - t2)
    objs.append(S)
for k, v in T.items():
    print("%-12s median %.2f ms  (first %.2f ms)" % (k, 1e3 * np.median(v[2:]), 1e3 * v[0]))
for S in objs: S.close()
