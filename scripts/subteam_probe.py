"""Mid-size systems (N = 80 000, cfg-5 larger variant): sub-teams of the cooperative grid, several
systems per launch, against one whole-grid launch per system."""
import os, sys, time, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); warnings.filterwarnings("ignore")
import numpy as np
import cpkrylov_b200 as cp
from cpkrylov_b200 import synth
from cpkrylov_b200.batch import BatchSolver
from cpkrylov_b200.ldl import ldl_superlu
g = int(sys.argv[1]) if len(sys.argv) > 1 else 40
cnt = int(sys.argv[2]) if len(sys.argv) > 2 else 14
opts = dict(atol=1e-6, rtol=1e-6, itmax=500, residual_update=True, nitref=1, force_itref=True)
systems = [synth.ipm_batch_lap3d(g, j) for j in range(cnt)]
facs = [ldl_superlu(synth.kp_matrix(w)) for w in systems]
rhs = [w["rhs"] for w in systems]
for tc in ("", "4", "8", "10", "16", "20", "37", "74", "148"):
    if tc: os.environ["CPK_TEAM_CTAS"] = tc
    else: os.environ.pop("CPK_TEAM_CTAS", None)
    bs = BatchSolver(systems, facs, opts)
    for _ in range(2): bs.solve("cpminres", rhs, opts)
    t0 = time.perf_counter(); xs, st = bs.solve("cpminres", rhs, opts); wall = time.perf_counter() - t0
    its = sum(d["niters"] for d in st)
    err = max(np.linalg.norm(x - w["xstar"]) / np.linalg.norm(w["xstar"]) for x, w in zip(xs, systems))
    print("team_ctas=%-7s launches %d  iters %d  device %.2f ms  wall %.2f ms  %.0f it/s (device)  max err vs x* %.1e"
          % (tc or "auto", bs.last_launches, its, bs.last_ms, 1e3 * wall, its / (1e-3 * bs.last_ms), err))
    bs.close()
