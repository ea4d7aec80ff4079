"""One-CTA team: compact walk (CPK_LDL_COMPACT=1, default) vs level walk on the fixture systems.

Run on the GPU box:  python scripts/compact_probe.py
Every configuration runs in its own process (the walk is chosen when the operator is created).
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import os, sys, warnings, ctypes as ct, json
ROOT = %r
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); warnings.filterwarnings("ignore")
import numpy as np
import cpkrylov_b200 as cp
from cpkrylov_b200 import _lib
from helpers import load_system, load_factors, load_oracle, EX_OPTS, relerr
out = {}
for name, method, extra in (("cvxqp1_m", "cpminres", {}), ("cvxqp1_m", "cpcg", {}), ("cvxqp2_s", "cpgmres", {"restart": 100}),
                            ("cvxqp2_s", "cpdqgmres", {"mem": 100})):
    s = load_system(name); fac = load_factors(name, "superlu"); orc = load_oracle(name)
    opts = dict(EX_OPTS, **extra)
    ts = []
    for rep in range(4):
        x, st, fl, S = cp.reg_cpkrylov(method, s["rhs"], s["Q"], s["A"], s["C"], s["G"], opts, factors=fac, return_system=True)
        ts.append(st["gpu"]["t_solve_ms"])
        S.close()
    key = "superlu/" + method + ("_restart100" if method == "cpgmres" else "_mem100" if method == "cpdqgmres" else "") + "/x"
    err = relerr(x, orc[key]) if key in orc.files else relerr(x, orc["x_direct"])
    out[name + ":" + method] = dict(iters=int(st["niters"]), solved=bool(fl["solved"]), ms=min(ts), us_per_iter=1e3 * min(ts) / max(st["niters"], 1), err_vs_oracle=err)
# stand-alone apply
s = load_system("cvxqp1_m"); fac = load_factors("cvxqp1_m", "superlu")
M = cp.opLDL2(s["G"], s["A"], -s["C"], factors=fac)
M.nitref = 0
z = np.random.default_rng(0).standard_normal(s["N"])
L = _lib.lib(); y = np.empty(s["N"]); ts = []
for _ in range(6):
    stt = _lib.StatsStruct(); _lib.check(L.cpk_ldl2_apply(M.handle, z.ctypes.data, y.ctypes.data, 0, ct.byref(stt))); ts.append(stt.t_solve_ms * 1e3)
from helpers import kp_of
out["apply_us"] = min(ts); out["apply_resid"] = float(np.linalg.norm(kp_of(s) @ y - z) / np.linalg.norm(z))
_lib.check(L.cpk_ldl2_set_track_rnorm(M.handle, 2))
for _ in range(2):
    stt = _lib.StatsStruct(); _lib.check(L.cpk_ldl2_apply(M.handle, z.ctypes.data, y.ctypes.data, 0, ct.byref(stt)))
out["walk_cycles(gather,ring wait,steps,scatter)"] = [int(stt.phase_cycles[i]) for i in range(4)]
_lib.check(L.cpk_ldl2_set_track_rnorm(M.handle, 0))
print("RESULT " + json.dumps(out))
'''


def run(env_extra):
    env = dict(os.environ, **env_extra)
    p = subprocess.run([sys.executable, "-c", CHILD.replace("%r", repr(ROOT), 1)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    res = None
    for line in p.stdout.splitlines():
        if line.startswith("RESULT "):
            res = json.loads(line[7:])
        elif "[cpk]" in line and ("compact" in line or "LDL sweep" in line):
            print("   ", line.strip())
    if res is None:
        print(p.stdout[-3000:])
    return res


if __name__ == "__main__":
    allres = {}
    configs = (("default", {"CPK_VERBOSE": "1"}), ("rows <= 768", {"CPK_LDL_TAIL_MAXLEN": "768"}), ("slack 50k", {"CPK_LDL_TAIL_SLACK": "50000"}),
               ("slack 400k", {"CPK_LDL_TAIL_SLACK": "400000"}))
    if "--quick" in sys.argv:
        configs = (("default", {"CPK_VERBOSE": "1"}),)
    for label, env in configs:
        r = run(env)
        allres[label] = r
        print(label, json.dumps(r, indent=1))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(allres, open(os.path.join(ROOT, "gpurun_out", "compact_probe.json"), "w"), indent=1)
