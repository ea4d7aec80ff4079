"""Sequence of IPM-like systems on the cvxqp1 pattern: what one more system costs with a fresh host
factorization + operator build, and with the in-place device refactorization."""
import os, sys, time, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); warnings.filterwarnings("ignore")
import numpy as np
import cpkrylov_b200 as cp
from cpkrylov_b200 import synth
from cpkrylov_b200.ldl import ldl_superlu, static_perm
from cpkrylov_b200.operators import KktSystem, opLDL2
base = synth.load_cvxqp1()
seq = [synth.ipm_batch_system(base, j) for j in range(10)]
opts = dict(atol=1e-6, rtol=1e-6, itmax=500, residual_update=True, nitref=1, force_itref=True)
# (a) every system from scratch
ta = []
for w in seq:
    t0 = time.perf_counter()
    fac = ldl_superlu(synth.kp_matrix(w)); M = opLDL2(w["G"], w["B"], -w["C"], factors=fac); S = KktSystem(w["H"], w["C"], M)
    t1 = time.perf_counter()
    x, st, fl = cp.reg_solve_on(S, "cpminres", w["rhs"], opts)
    t2 = time.perf_counter()
    ta.append((t1 - t0, t2 - t1, st["niters"]))
    S.close()
# (b) one operator, refactorized in place
t0 = time.perf_counter()
perm = static_perm(synth.kp_matrix(seq[0]))
w = seq[0]
M = opLDL2(w["G"], w["B"], -w["C"], factors="device", perm=perm); S = KktSystem(w["H"], w["C"], M)
t_first = time.perf_counter() - t0
tb = []
for j, w in enumerate(seq):
    t0 = time.perf_counter()
    if j:
        M.refactor(w["G"], w["B"], -w["C"]); S.update(w["H"], w["C"])
    t1 = time.perf_counter()
    x, st, fl = cp.reg_solve_on(S, "cpminres", w["rhs"], opts)
    t2 = time.perf_counter()
    tb.append((t1 - t0, t2 - t1, st["niters"]))
S.close()
med = lambda v: 1e3 * float(np.median(v))
print("from scratch : set-up %.2f ms (host LDL' + operator + system)   solve %.2f ms   iterations %s" % (med([t[0] for t in ta[2:]]), med([t[1] for t in ta[2:]]), [t[2] for t in ta]))
print("in place     : set-up %.2f ms (device LDL' + value refresh)     solve %.2f ms   iterations %s   (first system incl. plan: %.1f ms)"
      % (med([t[0] for t in tb[2:]]), med([t[1] for t in tb[2:]]), [t[2] for t in tb], 1e3 * t_first))
