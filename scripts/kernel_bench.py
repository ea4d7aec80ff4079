"""Per-kernel timings on the cfg-3 system: H*v, K_P*v, opLDL2 apply (stand-alone
launches of the same device phase functions the persistent solver uses)."""
import argparse, ctypes as ct, json, os, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); warnings.filterwarnings("ignore")
import numpy as np, torch
from cpkrylov_b200 import _lib, synth
from cpkrylov_b200.operators import opLDL2, KktSystem

ap = argparse.ArgumentParser()
ap.add_argument("--g", type=int, default=100); ap.add_argument("--k", type=int, default=2)
ap.add_argument("--window", type=int, default=0); ap.add_argument("--reps", type=int, default=30)
ap.add_argument("--workload", default="kkt_lap3d")
a = ap.parse_args()
w = (synth.kkt_lap3d if a.workload == "kkt_lap3d" else synth.kkt_convdiff)(g=a.g, k=a.k, window=a.window)
n, m = w["n"], w["m"]; N = n + m
M = opLDL2(w["G"], w["B"], -w["C"]); info = M.info()
S = KktSystem(w["H"], w["C"], M)
L = _lib.lib()
x = torch.randn(N, dtype=torch.float64, device="cuda"); y = torch.empty_like(x)
torch.cuda.synchronize()
def timeit(fn):
    ts = []
    for _ in range(a.reps):
        st = _lib.StatsStruct(); _lib.check(fn(ct.byref(st))); ts.append(st.t_solve_ms * 1e3)
    ts = np.array(ts[3:]); return float(np.median(ts)), float(ts.min())
spmv = lambda nnz, r, c: 12 * nnz + 4 * (r + 1) + 8 * c + 8 * r
res = {}
med, mn = timeit(lambda st: L.cpk_system_matvec(S.handle, 0, x.data_ptr(), y.data_ptr(), 1, st))
res["spmv_H"] = dict(us=med, us_min=mn, GBs=spmv(w["H"].nnz, n, n) / med / 1e3)
nnzKP = w["G"].nnz + 2 * w["B"].nnz + w["C"].nnz
med, mn = timeit(lambda st: L.cpk_ldl2_matvec(M.handle, x.data_ptr(), y.data_ptr(), 1, st))
res["spmv_KP"] = dict(us=med, us_min=mn, GBs=spmv(nnzKP, N, N) / med / 1e3)
B_ldl = 24 * info["nnz_L_off"] + 48 * N
for nitref in (0, 3):
    M.nitref = nitref
    med, mn = timeit(lambda st: L.cpk_ldl2_apply(M.handle, x.data_ptr(), y.data_ptr(), 1, st))
    b = B_ldl + (spmv(nnzKP, N, N) + 8 * N if nitref else 0)
    res["apply_nitref%d" % nitref] = dict(us=med, us_min=mn, GBs=b / med / 1e3)
res["info"] = info
print(json.dumps(res))
# debug: per-level cycles of the level walk (CTA 0, thread 0)
M.nitref = 0
_lib.check(L.cpk_ldl2_set_track_rnorm(M.handle, 2))
acc = np.zeros(8); ms = 0.0
for _ in range(10):
    st = _lib.StatsStruct(); _lib.check(L.cpk_ldl2_apply(M.handle, x.data_ptr(), y.data_ptr(), 1, ct.byref(st)))
    acc += np.array([st.phase_cycles[i] for i in range(8)]); ms += st.t_solve_ms
print("level walk cycles (work0, bar0, work1, bar1, ...):", (acc / 10).round(0).tolist(), "kernel us", ms * 100)
