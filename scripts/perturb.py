import ctypes as ct, os, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); warnings.filterwarnings("ignore")
import numpy as np, torch
from cpkrylov_b200 import _lib, synth
from cpkrylov_b200.operators import opLDL2, KktSystem
from cpkrylov_b200.solvers import _fill_opts
w = synth.kkt_lap3d(g=100); n, m = w["n"], w["m"]; N = n + m
M = opLDL2(w["G"], w["B"], -w["C"]); S = KktSystem(w["H"], w["C"], M); L = _lib.lib()
b = torch.from_numpy(w["rhs"]).cuda(); x = torch.empty(N, dtype=torch.float64, device="cuda"); torch.cuda.synchronize()
for nitref in (3, 0):
    M.nitref = nitref
    for prof in (0, 1, 2, 3):
        sid, o = _fill_opts("cpcg", dict(atol=0.0, rtol=1e-30, itmax=40), n, m); o.profile = prof
        cap = int(L.cpk_hist_capacity(sid, ct.byref(o))); hist = np.zeros((3, cap)); ts = []
        for _ in range(6):
            st = _lib.StatsStruct()
            rc = L.cpk_reg_solve(S.handle, sid, b.data_ptr(), ct.byref(o), x.data_ptr(), 1, ct.byref(st), hist.ctypes.data, cap)
            ts.append(st.t_solve_ms)
        print("nitref", nitref, "profile", prof, "rc", rc, "iters", st.niters, "us/iter %.1f" % (1e3 * np.median(ts[2:]) / max(1, st.niters)))
