"""Loader for the reference's example data format and a driver for sequences of
saddle-point systems.

The reference ships its systems as MATLAB ``.mat`` files holding ``K`` (the whole
saddle-point matrix), ``rhs``, and the block sizes ``nH``, ``nJ``, ``nZ``
(examples/cpk_exprog1.m:45-50, examples/cpk_exprog2.m:47-51).  Two layouts occur:

* ``"2x2"``  -- ``K = [Q A'; A -C]`` with ``n = nH`` (cpk_exprog1.m:48);
* ``"3x3"``  -- the permuted 3x3 interior-point form folded into the same 2x2
  partition with ``n = nH + nZ`` (cpk_exprog2.m:49).

``load_mat_system`` restates the block extraction of the example scripts
(cpk_exprog1.m:59-64): ``Q = K(1:n,1:n)``, ``G = diag(diag(Q))``,
``A = K(n+1:end,1:n)``, ``C = -K(n+1:end,n+1:end)``.

``solve_sequence`` solves a list of such systems (e.g. the iterations of one
interior-point run) with ``reg_cpkrylov`` semantics: systems small enough for the
one-CTA team go through ONE batched launch per device (``BatchSolver``), the others
one launch each.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from .batch import BatchSolver
from .ldl import static_perm
from .operators import KktSystem, opLDL2
from .solvers import reg_cpkrylov, reg_solve_on


def system_from_K(K, rhs, n, params=None):
    """Blocks of ``K = [Q A'; A -C]`` as the example scripts take them (cpk_exprog1.m:59-64)."""
    K = sp.csc_matrix(K)
    N = K.shape[0]
    if K.shape[0] != K.shape[1]:
        raise ValueError("K must be square")
    if not 0 < n <= N:
        raise ValueError("n = %d out of range for a %d x %d system" % (n, N, N))
    rhs = np.asarray(rhs, dtype=np.float64).reshape(-1)
    if rhs.size != N:
        raise ValueError("rhs has %d entries, K is %d x %d" % (rhs.size, N, N))
    Q = K[:n, :n].tocsc()
    G = sp.diags(Q.diagonal()).tocsc()
    A = K[n:, :n].tocsc()
    C = (-K[n:, n:]).tocsc()
    return dict(K=K, rhs=rhs, n=int(n), m=int(N - n), N=int(N), H=Q, B=A, C=C, G=G,
                Q=Q, A=A, params=dict(params or {}))


def load_mat_system(path, layout="auto"):
    """Reads one ``.mat`` file of the reference's examples.

    layout: "2x2" (n = nH), "3x3" (n = nH + nZ) or "auto" (3x3 when the file name says
    so, as the two shipped files do: ``*_2x2_*`` / ``*_3x3_*``)."""
    import scipy.io as sio
    d = sio.loadmat(path)
    for key in ("K", "rhs", "nH", "nJ"):
        if key not in d:
            raise ValueError("%s: no variable %r (expected K, rhs, nH, nJ, nZ)" % (path, key))
    nH, nJ = int(np.ravel(d["nH"])[0]), int(np.ravel(d["nJ"])[0])
    nZ = int(np.ravel(d["nZ"])[0]) if "nZ" in d else 0
    if layout == "auto":
        layout = "3x3" if "3x3" in str(path) else "2x2"
    if layout not in ("2x2", "3x3"):
        raise ValueError("layout must be '2x2', '3x3' or 'auto'")
    n = nH + nZ if layout == "3x3" else nH
    s = system_from_K(d["K"], d["rhs"], n, params=dict(file=str(path), layout=layout, nH=nH, nJ=nJ, nZ=nZ))
    if s["m"] != nJ:
        raise ValueError("%s: nJ = %d but K leaves %d constraint rows for n = %d" % (path, nJ, s["m"], n))
    return s


def solve_sequence(method, systems, opts=None, factors=None, device=0, batch_max_n=None):
    """Solves every system of the list; returns (xs, stats) in input order.

    Systems with N <= batch_max_n (default: the one-CTA team's limit, 24576) are solved
    in one batched launch, the rest one after the other."""
    if batch_max_n is None:
        batch_max_n = 24576
    xs = [None] * len(systems)
    stats = [None] * len(systems)
    small = [i for i, s in enumerate(systems) if s["N"] <= batch_max_n]
    large = [i for i in range(len(systems)) if i not in set(small)]
    if small:
        bs = BatchSolver([systems[i] for i in small], factors=None if factors is None else [factors[i] for i in small],
                         opts=opts, device=device)
        try:
            sol, st = bs.solve(method, [systems[i]["rhs"] for i in small], opts)
        finally:
            bs.close()
        for k, i in enumerate(small):
            xs[i], stats[i] = sol[k], st[k]
    for i in large:
        s = systems[i]
        x, st, fl = reg_cpkrylov(method, s["rhs"], s["H"], s["B"], s["C"], s["G"], opts,
                                 factors=None if factors is None else factors[i], device=device)
        st = dict(st, solved=fl["solved"])
        xs[i], stats[i] = x, st
    return xs, stats


def _same_pattern(a, b):
    a, b = sp.csc_matrix(a), sp.csc_matrix(b)
    a.sort_indices(); b.sort_indices()
    return a.shape == b.shape and np.array_equal(a.indptr, b.indptr) and np.array_equal(a.indices, b.indices)


def solve_ipm_sequence(method, systems, opts=None, device=0, perm=None):
    """The systems of ONE interior-point run, in order (each may depend on the previous
    solution, so they are solved one after the other): same sparsity patterns, new values.
    The operator is built once -- numeric LDL' on the GPU with a static fill-reducing
    permutation (SQD systems, opLDL2.m:81-82 does a fresh `ldl` per system) -- and refreshed in
    place for every further system (`opLDL2.refactor`, `KktSystem.update`).

    `systems` may be any iterable (e.g. a generator fed by the optimizer).  Yields
    (x, stats, flag) per system; stats["t_setup"] is the set-up time of that system."""
    import time
    M = S = None
    first = None
    try:
        for s in systems:
            t0 = time.perf_counter()
            if S is None:
                first = s
                if perm is None:
                    perm = static_perm(sp.bmat([[s["G"], s["B"].T], [s["B"], -sp.csc_matrix(s["C"])]], format="csc"))
                M = opLDL2(s["G"], s["B"], -sp.csc_matrix(s["C"]), factors="device", perm=perm, device=device)
                S = KktSystem(s["H"], s["C"], M)
            else:
                for key in ("H", "B", "C", "G"):
                    if not _same_pattern(first[key], s[key]):
                        raise ValueError("solve_ipm_sequence: block %s changed its sparsity pattern" % key)
                M.refactor(s["G"], s["B"], -sp.csc_matrix(s["C"]))
                S.update(s["H"], s["C"])
            t_setup = time.perf_counter() - t0
            x, st, fl = reg_solve_on(S, method, s["rhs"], opts)
            st["t_setup"] = t_setup
            yield x, st, fl
    finally:
        if S is not None:
            S.close()
        elif M is not None:
            M.close()
