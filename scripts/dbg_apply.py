import os, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); warnings.filterwarnings("ignore")
import numpy as np
import cpkrylov_b200 as cp
from oracle import cpk_oracle as orc
from helpers import *
from cpkrylov_b200.ldl import ldl_superlu, ldl_dense_bk
s = small_kkt(300, 90, seed=11); KP = kp_of(s)
for kind in ("superlu","densebk"):
    fac = (ldl_superlu if kind == "superlu" else ldl_dense_bk)(KP)
    zs = np.random.default_rng(5).standard_normal((3, s["N"]))
    for nitref, force, ru, stateful in [(0, False, False, False), (3, False, False, False), (2, True, False, False), (1, True, True, False), (1, True, True, True), (0, False, True, True), (3, False, True, True)]:
        Mo = orc.OpLDL2(s["G"], s["A"], -s["C"], *fac, ru_stateful=stateful)
        Mg = cp.opLDL2(s["G"], s["A"], -s["C"], factors=fac)
        for M in (Mo, Mg):
            M.nitref, M.force_itref, M.residual_update = nitref, force, ru
        Mg.ru_stateful = stateful; Mg.set_track_rnorm(True)
        for z in zs:
            n0 = Mo.nsolve
            yo, yg = Mo @ z, Mg @ z
            print(kind, nitref, force, ru, stateful, "rel %.2e" % relerr(yg, yo), "nsolve gpu/orc", Mg.last_stats["nldlsolve"], Mo.nsolve - n0,
                  "rnorm gpu/orc", Mg.rNorm if nitref else None, Mo.rNorm, "res %.1e" % relerr(KP @ yg, z), "|z| %.2e" % np.linalg.norm(z))
        Mg.close()
