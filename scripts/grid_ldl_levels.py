"""cfg 3: stand-alone M*z (nitref=0) on the grid team and the per-level work/wait cycles of CTA 0."""
import os, sys, warnings, ctypes as ct
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); warnings.filterwarnings("ignore")
import numpy as np
import cpkrylov_b200 as cp
from cpkrylov_b200 import _lib, synth
from cpkrylov_b200.ldl import ldl_superlu
os.environ["CPK_VERBOSE"] = "1"
g = int(sys.argv[1]) if len(sys.argv) > 1 else 100
w = synth.kkt_lap3d(g=g)
fac = ldl_superlu(synth.kp_matrix(w))
M = cp.opLDL2(w["G"], w["B"], -w["C"], factors=fac)
M.nitref = 0
N = w["n"] + w["m"]
z = np.random.default_rng(0).standard_normal(N); y = np.empty(N)
import torch
dz = torch.from_numpy(z).cuda(); dy = torch.empty_like(dz)
L = _lib.lib()
for mode in (0, 2):
    _lib.check(L.cpk_ldl2_set_track_rnorm(M.handle, mode))
    ts = []
    for _ in range(6):
        stt = _lib.StatsStruct(); _lib.check(L.cpk_ldl2_apply(M.handle, dz.data_ptr(), dy.data_ptr(), 1, ct.byref(stt))); ts.append(stt.t_solve_ms * 1e3)
    print("mode", mode, "apply us (device-resident)", min(ts), "cycles [work0 wait0 work1 wait1 ...]", [int(stt.phase_cycles[i]) for i in range(8)], M.info())
