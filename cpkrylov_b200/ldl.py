"""Host-side symmetric-indefinite factorization  P' K_P P = L D L'  (untimed setup).

Plays the role of MATLAB's ``[L,D,P] = ldl(K_P)`` (HSL MA57) inside the opLDL2
constructor (reference ops/opLDL2.m:82).  The factors are *inputs* to the GPU
operator handle: they are computed once on the host, uploaded once, and the time
is reported separately as ``stats.ptime`` (reference reg_cpkrylov.m:128-132).

Two back ends, both returning ``(L, d, e, perm)``:
  L    scipy.sparse CSR, unit lower triangular (unit diagonal stored)
  d    D(i,i)
  e    D(i+1,i); e[i] != 0 marks a 2x2 pivot that starts at i (len N, e[-1]=0)
  perm (P' x)[k] = x[perm[k]]   (column k of P is the unit vector e_perm[k])

* ``ldl_superlu``  SuperLU in symmetric mode with no diagonal pivoting: a sparse
  1x1-pivot LDL' with a fill-reducing symmetric ordering.  Valid whenever every
  symmetric permutation of K_P has an LDL' with 1x1 pivots -- true for the
  symmetric quasi-definite K_P = [G B'; B -C], G > 0, C > 0 of every BASELINE config.
* ``ldl_dense_bk``  LAPACK Bunch-Kaufman (scipy.linalg.ldl) for small systems;
  produces 2x2 pivots and so exercises that branch of the D-solve.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp
import scipy.sparse.linalg as spla


def ldl_superlu(K, check=True):
    K = sp.csc_matrix(K)
    N = K.shape[0]
    lu = spla.splu(K, permc_spec="MMD_AT_PLUS_A", diag_pivot_thresh=0.0,
                   options={"SymmetricMode": True})
    if not np.array_equal(lu.perm_r, lu.perm_c):
        raise RuntimeError("SuperLU pivoted off the diagonal: K_P has no 1x1-pivot LDL' "
                           "in this ordering; use ldl_dense_bk or supply MA57 factors")
    d = lu.U.diagonal().copy()
    if np.any(d == 0):
        raise RuntimeError("zero pivot in LDL'")
    L = sp.csr_matrix(lu.L)
    L.sort_indices()
    if check:
        # U must equal D L' (symmetry of the elimination); loose check on the scale of U
        R = (lu.U - sp.diags(d) @ lu.L.T).tocoo()
        scale = np.abs(lu.U.data).max()
        if R.nnz and np.abs(R.data).max() > 1e-6 * scale:
            raise RuntimeError("SuperLU factors are not symmetric (U != D L')")
    # Pr K Pc = L U with Pr = I[perm_r -> rows]: (Pr K Pc)[pr[i], pr[j]] = K[i, j]
    perm = np.argsort(lu.perm_r).astype(np.int64)
    return L, d, np.zeros(N), perm


def ldl_dense_bk(K):
    Kd = K.toarray() if sp.issparse(K) else np.asarray(K, dtype=np.float64)
    N = Kd.shape[0]
    lu, D, perm = sla.ldl(Kd, lower=True)
    Lp = lu[perm, :]
    assert np.allclose(np.triu(Lp, 1), 0.0)
    L = sp.csr_matrix(np.tril(Lp))
    L.sort_indices()
    d = np.diag(D).copy()
    e = np.concatenate([np.diag(D, -1), [0.0]])
    return L, d, e, np.asarray(perm, dtype=np.int64)


def ldl_factor(K, method="auto"):
    """Dispatch: 'superlu', 'dense_bk', or 'auto' (superlu)."""
    if method in ("auto", "superlu"):
        return ldl_superlu(K)
    if method == "dense_bk":
        return ldl_dense_bk(K)
    raise ValueError("unknown LDL method %r" % (method,))


def static_perm(K):
    """Fill-reducing symmetric permutation for the device factorization of a symmetric
    quasi-definite matrix (``opLDL2(..., factors="device", perm=...)``): the ordering a
    SuperLU symmetric-mode factorization of the pattern chooses (MMD on A'+A).  It depends on
    the pattern only, so one call serves a whole sequence.  perm[k] = original index of row k."""
    return ldl_superlu(K, check=False)[3]
