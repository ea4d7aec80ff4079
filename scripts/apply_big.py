"""A few stand-alone opLDL2 applies (nitref = 0, then 3) and K_P products on the cfg-3 system; used under ncu."""
import os, sys, warnings, ctypes as ct
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); warnings.filterwarnings("ignore")
import numpy as np, torch
from cpkrylov_b200 import _lib, synth
from cpkrylov_b200.operators import opLDL2
g = int(sys.argv[1]) if len(sys.argv) > 1 else 100
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
w = synth.kkt_lap3d(g=g)
N = w["n"] + w["m"]
M = opLDL2(w["G"], w["B"], -w["C"])
L = _lib.lib()
x = torch.randn(N, dtype=torch.float64, device="cuda"); y = torch.empty_like(x)
for nitref in (0, 3):
    M.nitref = nitref
    for _ in range(reps):
        st = _lib.StatsStruct(); _lib.check(L.cpk_ldl2_apply(M.handle, x.data_ptr(), y.data_ptr(), 1, ct.byref(st)))
    print("apply nitref=%d us %.1f" % (nitref, st.t_solve_ms * 1e3))
for _ in range(reps):
    st = _lib.StatsStruct(); _lib.check(L.cpk_ldl2_matvec(M.handle, x.data_ptr(), y.data_ptr(), 1, ct.byref(st)))
print("K_P matvec us %.1f" % (st.t_solve_ms * 1e3))
