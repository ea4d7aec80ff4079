"""Regenerates tests/golden/*.npz.  Runs ONLY in the build container (reads
/root/reference/examples/*.mat, which does not exist on the GPU box).

  python tests/golden/make_golden.py

What is stored
  <fixture>_system.npz   the saddle-point system of the reference's example
                         (examples/cpk_exprog1.m:45-64, cpk_exprog2.m:47-66): K in
                         COO form, rhs, n, m.  Data (c) cpkrylov authors, LGPLv3 --
                         see tests/golden/README.md.
  <fixture>_factors_<f>.npz  the LDL' factors (L, d, e, perm) handed to BOTH the
                         oracle and the CUDA path.
  <fixture>_oracle.npz   outputs of oracle/cpk_oracle.py (NOT of MATLAB: parity is
                         unpinned, see the oracle header) for every solver choice
                         the example scripts list, plus K\\rhs from a direct solve.
"""
import os
import sys
import warnings

import numpy as np
import scipy.io as sio
import scipy.sparse as sp
import scipy.sparse.linalg as spla

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

from oracle import cpk_oracle as orc                      # noqa: E402
from cpkrylov_b200.ldl import ldl_superlu, ldl_dense_bk   # noqa: E402

REF = "/root/reference/examples"
EX_OPTS = dict(print=False, atol=1e-6, rtol=1e-6, itmax=500,                     # cpk_exprog1.m:79-90
               residual_update=True, nitref=1, force_itref=True, itref_tol=1e-8)

FIXTURES = {
    "cvxqp1_m": dict(file="cvxqp1_m_2x2_symm_iter10.mat", n=lambda d: int(d["nH"][0, 0]),
                     runs=[("cpminres", {}), ("cpcg", {}), ("cpcglanczos", {}), ("cpdqgmres", {"mem": 2}),
                           ("cpsymmlq", {}), ("cpgmres", {"restart": 50})]),     # cpk_exprog1.m:67-74 (+symmlq, gmres)
    "cvxqp2_s": dict(file="cvxqp2_s_3x3_nonsymm_perm_iter10.mat",
                     n=lambda d: int(d["nH"][0, 0] + d["nZ"][0, 0]),
                     runs=[("cpgmres", {"restart": 100}), ("cpdqgmres", {"mem": 100}),   # cpk_exprog2.m:69-74
                           ("cpgmres", {"restart": 20}), ("cpdqgmres", {"mem": 10})]),
}


def blocks(K, n):
    K = sp.csc_matrix(K)
    Q = K[:n, :n]
    G = sp.diags(Q.diagonal())
    A = K[n:, :n]
    C = -K[n:, n:]
    return Q, A, C, G


def main():
    for name, fx in FIXTURES.items():
        d = sio.loadmat(os.path.join(REF, fx["file"]))
        K = sp.coo_matrix(d["K"])
        n = fx["n"](d)
        m = int(d["nJ"][0, 0])
        rhs = d["rhs"].ravel().astype(np.float64)
        np.savez_compressed(os.path.join(HERE, name + "_system.npz"),
                            row=K.row.astype(np.int32), col=K.col.astype(np.int32), val=K.data,
                            N=K.shape[0], n=n, m=m, rhs=rhs)
        Q, A, C, G = blocks(K, n)
        KP = sp.bmat([[G, A.T], [A, -C]], format="csc")
        out = {"x_direct": spla.spsolve(sp.csc_matrix(K), rhs)}
        facs = [("superlu", ldl_superlu)]
        if K.shape[0] < 1000:       # Bunch-Kaufman L of the larger fixture is 7 MB: recomputed live in the tests
            facs.append(("densebk", ldl_dense_bk))
        for fname, fac in facs:
            L, dd, ee, perm = fac(KP)
            Lc = sp.coo_matrix(L)
            np.savez_compressed(os.path.join(HERE, "%s_factors_%s.npz" % (name, fname)),
                                Lrow=Lc.row.astype(np.int32), Lcol=Lc.col.astype(np.int32), Lval=Lc.data,
                                d=dd, e=ee, perm=perm.astype(np.int32))
            for meth, extra in fx["runs"]:
                o = dict(EX_OPTS); o.update(extra)
                tag = "%s/%s%s" % (fname, meth, "".join("_%s%d" % kv for kv in sorted(extra.items())))
                x, st, fl = orc.reg_cpkrylov(meth, rhs, Q, A, C, G, o, factor=lambda _K: (L, dd, ee, perm))
                out[tag + "/x"] = x
                out[tag + "/niters"] = st["niters"]
                out[tag + "/solved"] = fl["solved"]
                for hk in ("residHistory", "cgresidHistory", "lqresidHistory", "qrresidHistory"):
                    if hk in st:
                        out[tag + "/" + hk] = st[hk]
                print(name, tag, st["niters"], fl["solved"],
                      "relerr vs direct %.2e" % (np.linalg.norm(x - out["x_direct"]) / np.linalg.norm(out["x_direct"])))
        np.savez_compressed(os.path.join(HERE, name + "_oracle.npz"), **out)


if __name__ == "__main__":
    main()
