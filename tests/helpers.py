"""Shared test helpers: fixture loading, synthetic small KKT systems."""
import os

import numpy as np
import scipy.sparse as sp

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

EX_OPTS = dict(print=False, atol=1e-6, rtol=1e-6, itmax=500,          # examples/cpk_exprog1.m:79-90
               residual_update=True, nitref=1, force_itref=True, itref_tol=1e-8)


def load_system(name):
    d = np.load(os.path.join(GOLDEN, name + "_system.npz"))
    N, n, m = int(d["N"]), int(d["n"]), int(d["m"])
    K = sp.csc_matrix((d["val"], (d["row"], d["col"])), shape=(N, N))
    Q = K[:n, :n].tocsc()
    G = sp.diags(Q.diagonal()).tocsc()          # examples/cpk_exprog1.m:60-61
    A = K[n:, :n].tocsc()
    C = (-K[n:, n:]).tocsc()
    return dict(K=K, rhs=d["rhs"], n=n, m=m, N=N, Q=Q, A=A, C=C, G=G)


def load_factors(name, kind):
    d = np.load(os.path.join(GOLDEN, "%s_factors_%s.npz" % (name, kind)))
    N = d["d"].size
    L = sp.csr_matrix((d["Lval"], (d["Lrow"], d["Lcol"])), shape=(N, N))
    return L, d["d"], d["e"], d["perm"].astype(np.int64)


def load_oracle(name):
    return np.load(os.path.join(GOLDEN, name + "_oracle.npz"))


def kp_of(s):
    return sp.bmat([[s["G"], s["A"].T], [s["A"], -s["C"]]], format="csc")


def small_kkt(n=60, m=20, seed=0, nonsym=False, density=0.08, creg=1e-2):
    """Random small regularized saddle-point system with SPD (1,1) block."""
    rng = np.random.default_rng(seed)
    R = sp.random(n, n, density=density, random_state=rng, data_rvs=rng.standard_normal)
    H = (R @ R.T + sp.identity(n) * (1.0 + rng.random())).tocsc()
    if nonsym:
        S = sp.random(n, n, density=density / 2, random_state=rng, data_rvs=rng.standard_normal)
        H = (H + 0.3 * (S - S.T)).tocsc()
    B = sp.random(m, n, density=0.15, random_state=rng, data_rvs=rng.standard_normal).tocsc()
    B = (B + sp.csc_matrix((np.ones(m), (np.arange(m), rng.permutation(n)[:m])), shape=(m, n))).tocsc()
    C = (creg * sp.identity(m)).tocsc()
    G = sp.diags(H.diagonal()).tocsc()
    xs = rng.standard_normal(n + m)
    K = sp.bmat([[H, B.T], [B, -C]], format="csc")
    return dict(K=K, rhs=K @ xs, n=n, m=m, N=n + m, Q=H, A=B, C=C, G=G, xs=xs)


def relerr(a, b):
    nb = np.linalg.norm(b)
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / (nb if nb > 0 else 1.0)
