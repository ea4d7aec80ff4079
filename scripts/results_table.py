"""Round-end results table: every solver on the synthetic configs (device-resident
solve, CUDA-event time of the one launch), with algorithmic GB/s."""
import ctypes as ct, json, os, sys, time, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); warnings.filterwarnings("ignore")
import numpy as np, torch
from cpkrylov_b200 import _lib, synth
from cpkrylov_b200.operators import opLDL2, KktSystem
from cpkrylov_b200.solvers import _fill_opts, apply_opts_to_M
sys.argv = [sys.argv[0]] + sys.argv[1:]
import bench

def run(w, solver, opts, reps=4, mem=0, restart=0):
    n, m = w["n"], w["m"]; N = n + m
    M = opLDL2(w["G"], w["B"], -w["C"]); S = KktSystem(w["H"], w["C"], M); apply_opts_to_M(M, opts)
    info = M.info(); L = _lib.lib()
    sid, o = _fill_opts(solver, opts, n, m)
    cap = int(L.cpk_hist_capacity(sid, ct.byref(o))); hist = np.zeros((3, cap))
    b = torch.from_numpy(w["rhs"]).cuda(); x = torch.empty(N, dtype=torch.float64, device="cuda"); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        st = _lib.StatsStruct()
        rc = L.cpk_reg_solve(S.handle, sid, b.data_ptr(), ct.byref(o), x.data_ptr(), 1, ct.byref(st), hist.ctypes.data, cap)
        if rc: return dict(solver=solver, error=_lib.last_error())
        ts.append(st.t_solve_ms)
    d = _lib.stats_to_dict(st)
    ms = float(np.median(ts[1:]))
    by, parts = bench.algorithmic_bytes(dict(H=w["H"], C=w["C"], G=w["G"], B=w["B"], n=n, m=m), info, solver, d, mem=mem, restart=restart)
    err = float(np.linalg.norm(x.cpu().numpy() - w["xstar"]) / np.linalg.norm(w["xstar"]))
    S.close()
    return dict(solver=solver, opts=opts, iters=d["niters"], solved=d["solved"], ms=ms, it_per_s=1e3 * d["niters"] / ms,
                GBs=by / ms / 1e6, frac=by / ms / 1e6 / 6549.4, relerr=err, napply=d["napply"], nldlsolve=d["nldlsolve"], nresid=d["nresid"])

out = []
only = os.environ.get("CPK_RESULTS_ONLY", "")
w3 = synth.kkt_lap3d(g=100)
for solver, opts in [] if only == "cfg4" else [("cpcg", {}), ("cpcg", {"nitref": 0}), ("cpcglanczos", {}), ("cpminres", {}), ("cpsymmlq", {}),
                     ("cpgmres", {"restart": 20}), ("cpdqgmres", {"mem": 20})]:
    r = run(w3, solver, dict(opts, atol=1e-6, rtol=1e-6), mem=opts.get("mem", 0), restart=opts.get("restart", 0)); r["config"] = "cfg3 kkt_lap3d g=100"; out.append(r); print(json.dumps(r), flush=True)
del w3
w4 = synth.kkt_convdiff(g=126)
for solver, opts in [("cpdqgmres", {"mem": 20}), ("cpdqgmres", {"mem": 50}), ("cpgmres", {"restart": 50})]:
    o = dict(opts, atol=1e-6, rtol=1e-6, itmax=500, nitref=1, force_itref=True)
    r = run(w4, solver, o, reps=3, mem=opts.get("mem", 0), restart=opts.get("restart", 0)); r["config"] = "cfg4 kkt_convdiff g=126"; out.append(r); print(json.dumps(r), flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", (os.environ.get("CPK_RESULTS_TAG", "r2")) + "_results_table.json"), "w"), indent=1)
