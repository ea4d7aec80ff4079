"""Seeded synthetic regularized KKT systems of the BASELINE configs (SURVEY.md
section 8d).  Host-side input generation only (numpy/scipy); every generator
returns a dict with H (n x n), B (m x n), C (m x m), G = diag(H), rhs = K x*,
xstar and the parameters that produced it (written into every result file).
"""
from __future__ import annotations

import os

import numpy as np
import scipy.sparse as sp


def _lap1d(g):
    return sp.diags([-np.ones(g - 1), 2.0 * np.ones(g), -np.ones(g - 1)], [-1, 0, 1], format="csr")


def _kron3(T, g):
    I = sp.identity(g, format="csr")
    return (sp.kron(sp.kron(T, I), I) + sp.kron(sp.kron(I, T), I) + sp.kron(sp.kron(I, I), T)).tocsr()


def _random_B(m, n, k, seed, window=0):
    """m x n, k entries per row, values N(0,1).  window=0: uniform-random columns;
    window=w: columns drawn in a w-wide window centred at column (n/m)*i."""
    rng = np.random.default_rng(seed)
    if window:
        centre = (np.arange(m) * (n // m)).reshape(-1, 1)
        cols = (centre + rng.integers(-window // 2, window // 2, size=(m, k))) % n
    else:
        cols = rng.integers(0, n, size=(m, k))
    rows = np.repeat(np.arange(m), k)
    B = sp.csr_matrix((rng.standard_normal(m * k), (rows, cols.ravel())), shape=(m, n))
    B.sum_duplicates()
    return B


def _finish(H, B, creg, seed_x, params):
    n, m = H.shape[0], B.shape[0]
    C = (creg * sp.identity(m)).tocsr()
    G = sp.diags(H.diagonal()).tocsr()
    rng = np.random.default_rng(seed_x)
    xstar = rng.standard_normal(n + m)
    b = np.concatenate([H @ xstar[:n] + B.T @ xstar[n:], B @ xstar[:n] - C @ xstar[n:]])
    nnzK = H.nnz + 2 * B.nnz + C.nnz
    params = dict(params, n=n, m=m, nnz_H=int(H.nnz), nnz_B=int(B.nnz), nnz_K=int(nnzK))
    return dict(H=H.tocsc(), B=B.tocsc(), C=C.tocsc(), G=G.tocsc(), rhs=b, xstar=xstar, n=n, m=m, params=params)


def kkt_lap3d(g=100, k=2, window=0, seed_B=1, seed_x=3, creg=1e-6):
    """cfg 3: H = 7-point Dirichlet Laplacian on a g^3 grid (diag 6, off -1),
    B = random sparse m x n with m = n/4, C = creg*I, G = diag(H)."""
    n = g ** 3
    H = _kron3(_lap1d(g), g)
    B = _random_B(n // 4, n, k, seed_B, window)
    return _finish(H, B, creg, seed_x, dict(workload="kkt_lap3d", g=g, k=k, window=window,
                                           seed_B=seed_B, seed_x=seed_x, creg=creg))


def kkt_convdiff(g=126, k=2, window=0, seed_B=4, seed_x=3, creg=1e-6, peclet=0.5):
    """cfg 4: H = -Laplace + central-difference convection, cell Peclet number
    `peclet` in each direction (nonsymmetric, positive-definite symmetric part)."""
    n = g ** 3
    T = sp.diags([(-1.0 - peclet) * np.ones(g - 1), 2.0 * np.ones(g), (-1.0 + peclet) * np.ones(g - 1)],
                 [-1, 0, 1], format="csr")
    H = _kron3(T, g)
    B = _random_B(n // 4, n, k, seed_B, window)
    return _finish(H, B, creg, seed_x, dict(workload="kkt_convdiff", g=g, k=k, window=window,
                                           seed_B=seed_B, seed_x=seed_x, creg=creg, peclet=peclet))


def ipm_batch_system(base, j, nZ=2000):
    """cfg 5: system j of an IPM-like sequence sharing the pattern of ``base``
    (a dict with K, rhs, n, m as returned by tests.helpers.load_system or
    ``load_cvxqp1``): the trailing nZ diagonal entries of the (1,1) block (the
    S^-1 Z + rho I part, examples/cpk_exprog1.m:13-16) are scaled by 10^u,
    u ~ U(-2, 2), seed 100 + j."""
    n, m = base["n"], base["m"]
    rng = np.random.default_rng(100 + j)
    K = sp.csc_matrix(base["K"]).copy()
    H = K[:n, :n].tolil()
    dg = H.diagonal()
    scale = np.ones(n)
    scale[n - nZ:] = 10.0 ** rng.uniform(-2.0, 2.0, size=nZ)
    H.setdiag(dg * scale)
    H = H.tocsc()
    B = K[n:, :n].tocsc()
    C = (-K[n:, n:]).tocsc()
    G = sp.diags(H.diagonal()).tocsc()
    return dict(H=H, B=B, C=C, G=G, rhs=np.asarray(base["rhs"], dtype=np.float64).copy(), n=n, m=m,
                params=dict(workload="ipm_batch", j=j, nZ=nZ, seed=100 + j))


def ipm_batch_lap3d(g, j, k=2):
    """cfg 5, larger variant (g=40, n=64 000, m=16 000): system j of an
    IPM-like sequence on the cfg-3 pattern.  H_j = Laplacian + diag(rho), rho = 10^u,
    u ~ U(-2, 2), seed 100 + j (the positive diagonal S^-1 Z + rho I of an interior-point
    iteration, examples/cpk_exprog1.m:13-16); B, C and the right-hand side recipe are shared."""
    base = kkt_lap3d(g=g, k=k)
    n = base["n"]
    rng = np.random.default_rng(100 + j)
    H = (base["H"] + sp.diags(10.0 ** rng.uniform(-2.0, 2.0, size=n))).tocsc()
    w = _finish(H, base["B"], 1e-6, 3, dict(workload="ipm_batch_lap3d", g=g, k=k, j=j, seed=100 + j))
    return w


def load_cvxqp1():
    """The reference's example system (tests/golden/cvxqp1_m_system.npz)."""
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    d = np.load(os.path.join(here, "tests", "golden", "cvxqp1_m_system.npz"))
    N, n, m = int(d["N"]), int(d["n"]), int(d["m"])
    K = sp.csc_matrix((d["val"], (d["row"], d["col"])), shape=(N, N))
    return dict(K=K, rhs=d["rhs"], n=n, m=m, N=N)


def kkt_matrix(s):
    return sp.bmat([[s["H"], s["B"].T], [s["B"], -s["C"]]], format="csc")


def kp_matrix(s):
    return sp.bmat([[s["G"], s["B"].T], [s["B"], -s["C"]]], format="csc")
