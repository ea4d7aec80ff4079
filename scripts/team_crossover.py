"""One-CTA team vs cooperative grid on mid-size systems (where CPK_CTA_MAX_N draws the line)."""
import os, sys, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, warnings, json
sys.path.insert(0, ROOT_DIR); warnings.filterwarnings("ignore")
import numpy as np
import cpkrylov_b200 as cp
from cpkrylov_b200 import synth
from cpkrylov_b200.ldl import ldl_superlu
out = {}
cases = [("lap3d g=%d k=2" % g, dict(g=g)) for g in (10, 12, 15, 18, 21, 24, 27)]
cases += [("lap3d g=%d k=6 window=64" % g, dict(g=g, k=6, window=64, seed_B=2)) for g in (12, 18, 24)]
for label, kw in cases:
    w = synth.kkt_lap3d(**kw); fac = ldl_superlu(synth.kp_matrix(w))
    o = dict(atol=1e-6, rtol=1e-6, itmax=500, residual_update=True, nitref=1, force_itref=True)
    best = None
    for rep in range(3):
        x, st, fl, S = cp.reg_cpkrylov("cpminres", w["rhs"], w["H"], w["B"], w["C"], w["G"], o, factors=fac, return_system=True)
        info = S.M.info(); S.close()
        best = st["gpu"]["t_solve_ms"] if best is None else min(best, st["gpu"]["t_solve_ms"])
    out[label] = dict(N=w["n"] + w["m"], iters=int(st["niters"]), ms=best, us_per_iter=1e3 * best / max(st["niters"], 1), levels=[info["levels_fwd"], info["levels_bwd"]])
print("RESULT " + json.dumps(out))
'''.replace('ROOT_DIR', repr(ROOT))
for team in ("cta", "grid", "auto"):
    env = dict(os.environ)
    env.pop("CPK_TEAM", None)
    if team != "auto":
        env["CPK_TEAM"] = team
    p = subprocess.run([sys.executable, "-c", CHILD], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    for line in p.stdout.splitlines():
        if line.startswith("RESULT "):
            for k, v in json.loads(line[7:]).items():
                print("%-5s %-28s N=%d levels %s iters %d  %.2f ms  %.1f us/iteration" % (team, k, v["N"], v["levels"], v["iters"], v["ms"], v["us_per_iter"]))
            break
    else:
        print(team, p.stdout[-1500:])
