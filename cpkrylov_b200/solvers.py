"""The reference's solver entry points and driver, same names and argument
meaning, running on the GPU through libcpk_b200.so.

    [x, y, stats, flag] = method(b1, A, C, M, opts)     kernels/cp*.m
    [x, stats, flag]    = reg_cpkrylov(method, b, A, B, C, G, opts)   reg_cpkrylov.m:1

``stats``/``flag`` are dicts with the reference's field names.
"""
from __future__ import annotations

import ctypes as ct
import time

import numpy as np
import scipy.sparse as sp

from . import _lib
from .operators import KktSystem, opLDL2, _vec


class SolverError(RuntimeError):
    """error()/MException raised inside a reference solver (indefinite
    preconditioner etc.).  ``identifier`` is the MException id where the
    reference sets one (cpcglanczos.m:161)."""

    def __init__(self, msg, identifier="", code=0):
        super().__init__(msg)
        self.identifier = identifier
        self.code = code


_STATUS_TEXT = {0: "maximum number of iterations attained",                        # cpcglanczos.m:315
                1: "residual small compared to initial residual",                  # :318
                2: "backward error small"}                                         # :323


def _fill_opts(name, opts, n, m):
    sid = _lib.SOLVER_IDS[name]
    o = _lib.OptsStruct()
    _lib.lib().cpk_opts_default(ct.byref(o), sid, n, m)
    opts = opts or {}
    for f in ("atol", "rtol", "btol"):
        if f in opts:
            setattr(o, f, float(opts[f]))
    if "itmax" in opts:
        o.itmax = int(opts["itmax"])
    if "restart" in opts:
        o.restart = int(opts["restart"])
    if "mem" in opts:
        o.mem = int(opts["mem"])
    if opts.get("profile"):
        o.profile = 1
    return sid, o


def _finish(name, st, hist, cap, rc, opts):
    if rc in (_lib.CPK_ERR_INDEFINITE, _lib.CPK_ERR_BREAKDOWN):
        ident = "CPCGLanczos:IndefiniteError" if (name == "cpcglanczos" and rc == _lib.CPK_ERR_INDEFINITE) else ""
        raise SolverError(_lib.last_error(), identifier=ident, code=rc)
    _lib.check(rc)
    d = _lib.stats_to_dict(st)
    L = d["hist_len"]
    stats = {"niters": d["niters"]}
    if name == "cpsymmlq":                                                          # cpsymmlq.m:363-366
        stats["cgresidHistory"] = hist[0, :L].copy()
        stats["lqresidHistory"] = hist[1, :L].copy()
        stats["qrresidHistory"] = hist[2, :L].copy()
    else:
        stats["residHistory"] = hist[0, :L].copy()
    if name == "cpcglanczos":
        stats["status"] = _STATUS_TEXT[d["status"]]
    stats["gpu"] = d
    flag = {"solved": d["solved"]}
    if (opts or {}).get("print", False):
        _print_history(name, stats)
    return stats, flag


def _print_history(name, stats):
    # opts.print: the reference prints one line per iteration from inside the loop
    # (e.g. cpminres.m:167-173,239-241); the device loop cannot, so the table is
    # printed from the returned history.
    h = stats.get("residHistory", stats.get("cgresidHistory"))
    print("\n**** %s on B200 ****\n" % name)
    print("%5s  %9s" % ("iter", "|resid|"))
    for k, v in enumerate(h):
        print("%5d  %9.2e" % (k, v))
    print()


def _content_key(X):
    """Identity of a sparse matrix by CONTENT (shape, pattern and value checksums), cheap enough
    to evaluate on every call: an `id()` would hit a stale device copy after an in-place change
    of the values, or after the interpreter reuses the id of a collected temporary."""
    X = X if sp.issparse(X) else sp.csc_matrix(X)
    X = X.tocsc() if X.format not in ("csc", "csr") else X
    return (X.format, X.shape, int(X.nnz), int(X.indptr[-1]), int(np.bitwise_xor.reduce(X.indices.astype(np.int64) * 2654435761 % (1 << 31))) if X.nnz else 0,
            float(X.data.sum()), float(np.abs(X.data).sum()), float(X.data[::7].sum()))


def _system_of(A, Cm, M):
    """The device system (A, C, M) behind ``method(b1, A, C, M, opts)``.  One opLDL2 may be
    used with any A, C (the reference passes M next to them, e.g. cpminres.m:1): the device copy
    of (A, C) is cached on M by content and replaced -- without touching M -- when another
    pair arrives."""
    if isinstance(A, KktSystem):
        return A
    # an operator has no content to hash: it is keyed by identity (the system keeps a reference,
    # so the id cannot be reused while the cache entry lives)
    key = (_content_key(A) if (sp.issparse(A) or isinstance(A, np.ndarray)) else ("operator", id(A)), _content_key(Cm))
    sysobj = getattr(M, "_system", None)
    if sysobj is not None and (sysobj.handle is None or sysobj._key != key):
        sysobj.close()                  # frees the device copy of the old (A, C); M stays (it is the caller's)
        sysobj = None
    if sysobj is None:
        sysobj = KktSystem(A, Cm, M, owns_M=False)
        sysobj._key = key
        M._system = sysobj
    return sysobj


def _make_solver(name):
    def solver(b, A, Cm, M, opts=None):
        S = _system_of(A, Cm, M)
        n, m = S.n, S.m
        b = _vec(b, n, "b")
        sid, o = _fill_opts(name, opts, n, m)
        cap = int(_lib.lib().cpk_hist_capacity(sid, ct.byref(o)))
        hist = np.zeros((3, cap))
        x = np.empty(n); y = np.empty(m)
        st = _lib.StatsStruct()
        rc = _lib.lib().cpk_solve(S.handle, sid, b.ctypes.data, ct.byref(o), x.ctypes.data, y.ctypes.data,
                                  _lib.MEM_HOST, ct.byref(st), hist.ctypes.data, cap)
        S.check_callback()
        stats, flag = _finish(name, st, hist, cap, rc, opts)
        return x, y, stats, flag
    solver.__name__ = name
    solver.__doc__ = "[x, y, stats, flag] = %s(b, A, C, M, opts) -- reference kernels/%s.m on the GPU." % (name, name)
    solver.cpk_name = name
    return solver


cpcg = _make_solver("cpcg")
cpcglanczos = _make_solver("cpcglanczos")
cpminres = _make_solver("cpminres")
cpsymmlq = _make_solver("cpsymmlq")
cpgmres = _make_solver("cpgmres")
cpdqgmres = _make_solver("cpdqgmres")
SOLVERS = {f.cpk_name: f for f in (cpcg, cpcglanczos, cpminres, cpsymmlq, cpgmres, cpdqgmres)}


def _method_name(method):
    if isinstance(method, str):
        name = method.lstrip("@")
    else:
        name = getattr(method, "cpk_name", getattr(method, "__name__", None))
    if name not in SOLVERS:
        raise ValueError("reg_cpkrylov: unknown method %r" % (method,))
    return name


def apply_opts_to_M(M, opts):
    """reg_cpkrylov.m:135-148."""
    if opts is None:
        return
    if "nitref" in opts:
        M.nitref = opts["nitref"]
    if "itref_tol" in opts:
        M.itref_tol = opts["itref_tol"]
    if "residual_update" in opts:
        M.residual_update = opts["residual_update"]
    if "force_itref" in opts:
        M.force_itref = opts["force_itref"]
    if "ru_stateful" in opts:       # extension: Spot handle-class reading of residual_update
        M.ru_stateful = opts["ru_stateful"]


def reg_solve_on(S, method, b, opts=None):
    """The solve part of reg_cpkrylov (reg_cpkrylov.m:150-178) on an existing device system
    ``S`` (KktSystem): rhs shift, Krylov loop, un-shift, in one launch.  Used for sequences
    whose systems are refreshed in place (``S.update`` + ``S.M.refactor``)."""
    name = _method_name(method)
    n, m = S.n, S.m
    apply_opts_to_M(S.M, opts)
    tstarts = time.perf_counter()
    b = _vec(b, n + m, "b")
    sid, o = _fill_opts(name, opts, n, m)
    cap = int(_lib.lib().cpk_hist_capacity(sid, ct.byref(o)))
    hist = np.zeros((3, cap))
    x = np.empty(n + m)
    st = _lib.StatsStruct()
    rc = _lib.lib().cpk_reg_solve(S.handle, sid, b.ctypes.data, ct.byref(o), x.ctypes.data, _lib.MEM_HOST,
                                  ct.byref(st), hist.ctypes.data, cap)
    S.check_callback()
    stats, flag = _finish(name, st, hist, cap, rc, opts)
    stats["stime"] = time.perf_counter() - tstarts
    return x, stats, flag


def reg_cpkrylov(method, b, A, B, Cm, G, opts=None, factors=None, ldl_method="auto", device=0,
                 return_system=False, perm=None):
    """[x, stats, flag] = reg_cpkrylov(method, b, A, B, C, G, opts)   (reg_cpkrylov.m:1).

    ``method``: one of this module's solver functions (or its name).  Extra
    keyword arguments are not in the reference: ``factors=(L,d,e,perm)`` supplies
    the LDL' of [G B'; B -C] (e.g. MATLAB's ldl/MA57 output) instead of
    factorizing here, ``factors="device", perm=p`` factorizes on the GPU with the static
    permutation p (symmetric quasi-definite K_P); ``device`` selects the GPU.
    """
    if method is None or b is None or A is None or B is None or Cm is None or G is None:
        raise ValueError("reg_cpkrylov: not enough inputs")                         # reg_cpkrylov.m:122-125
    name = _method_name(method)
    tstartp = time.perf_counter()
    B = sp.csc_matrix(B); Cm = sp.csc_matrix(Cm); G = sp.csc_matrix(G)
    n, m = A.shape[0], B.shape[0]
    M = opLDL2(G, B, -Cm, factors=factors, ldl_method=ldl_method, device=device, perm=perm)    # :131
    S = KktSystem(A, Cm, M, owns_M=True)
    ptime = time.perf_counter() - tstartp
    apply_opts_to_M(M, opts)

    tstarts = time.perf_counter()
    b = _vec(b, n + m, "b")
    sid, o = _fill_opts(name, opts, n, m)
    cap = int(_lib.lib().cpk_hist_capacity(sid, ct.byref(o)))
    hist = np.zeros((3, cap))
    x = np.empty(n + m)
    st = _lib.StatsStruct()
    rc = _lib.lib().cpk_reg_solve(S.handle, sid, b.ctypes.data, ct.byref(o), x.ctypes.data, _lib.MEM_HOST,
                                  ct.byref(st), hist.ctypes.data, cap)
    try:
        S.check_callback()
        stats, flag = _finish(name, st, hist, cap, rc, opts)
    except Exception:
        S.close()
        raise
    stime = time.perf_counter() - tstarts
    stats["ptime"] = ptime                                                          # :177-178
    stats["stime"] = stime
    stats["t_factor"] = M.t_factor
    stats["t_upload"] = M.t_upload
    if return_system:
        return x, stats, flag, S
    S.close()
    return x, stats, flag
