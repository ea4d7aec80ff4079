"""
oracle/cpk_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU (NumPy/SciPy + oracle/kernels.c) restatement of cpkrylov's per-iteration hot
path, statement for statement, so that the CUDA path in ``cpkrylov_b200`` can be
checked against it on identical inputs *including identical LDL' factors*.

Reference lines followed (all relative to /root/reference):
    util/SymGivens.m:1-29            -> sym_givens
    ops/opLDL2.m:60-92,161-195       -> OpLDL2
    kernels/cpcg.m:95-193            -> cpcg
    kernels/cpcglanczos.m:108-325    -> cpcglanczos
    kernels/cpminres.m:91-252        -> cpminres
    kernels/cpsymmlq.m:98-367        -> cpsymmlq
    kernels/cpgmres.m:99-269         -> cpgmres
    kernels/cpdqgmres.m:98-280       -> cpdqgmres
    reg_cpkrylov.m:122-178           -> reg_cpkrylov

Third-party arithmetic that is NOT under /root/reference and is restated here:
  * Spot toolbox (github.com/mpf/spot, "master", no pinned version;
    README.md:42-52): only operator dispatch -- P*inv(L')*inv(D)*inv(L)*P'
    evaluated right to left (ops/opLDL2.m:86).  Its value-class semantics decide
    whether ``residual_update`` keeps state between applies; both readings are
    implemented (``ru_stateful``; default False = value class = no state).
  * MATLAB builtins ldl (HSL MA57), sparse mtimes, sparse "\\", dot, norm
    (tested under MATLAB 2018b, README.md:47).  The factorization is an INPUT
    here (L, d, e, perm); see cpkrylov_b200/ldl.py for how the host obtains it.

PARITY STATUS: **UNPINNED**.  The reference has no tests, golden vectors or
recorded outputs, and neither MATLAB nor Octave is available in the build
container, so this restatement cannot be checked against a run of the
reference.  It is anchored instead on (i) the reference's own self-check
``norm(x - K\\rhs)/norm(x)`` for its two example systems
(examples/cpk_exprog1.m:101-104, examples/cpk_exprog2.m:100-103), (ii) the
operator identity K_P*(M*z) = z, (iii) solver invariants listed in
tests/test_oracle.py.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline leg and
``--impl reference``) may import this module.
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess

import numpy as np
import scipy.sparse as sp

_HERE = os.path.dirname(os.path.abspath(__file__))
_EPS = float(np.finfo(np.float64).eps)

# --------------------------------------------------------------------------
# Working precision.  The restatement runs in IEEE fp64 like the reference.  As an ARBITER
# for parity tests (which of two fp64 runs is closer to the exactly-rounded result of the same
# algorithm?) every vector and scalar can instead be carried in the C `long double` of the host
# (x87 80-bit on x86-64: 64-bit mantissa): ``with extended_precision(): ...``.  Matrices,
# factors and thresholds (eps) stay fp64 data; only the arithmetic widens.
# --------------------------------------------------------------------------
_RT = np.float64


class extended_precision:
    def __enter__(self):
        global _RT
        self._old = _RT
        _RT = np.longdouble
        return self

    def __exit__(self, *a):
        global _RT
        _RT = self._old


def is_extended():
    return _RT is np.longdouble


def _f(v):
    return _RT(v)


def _zeros(shape, order="C"):
    return np.zeros(shape, dtype=_RT, order=order)


def _empty(shape):
    return np.empty(shape, dtype=_RT)


def _sqrt(v):
    return math.sqrt(v) if _RT is np.float64 else np.sqrt(_RT(v))


def _hypot(a, b):
    return math.hypot(a, b) if _RT is np.float64 else np.hypot(_RT(a), _RT(b))


# --------------------------------------------------------------------------
# C kernels (oracle/kernels.c) with a SciPy/NumPy fallback of the same loops
# --------------------------------------------------------------------------
def build_c_kernels(force: bool = False) -> str:
    """Compile oracle/kernels.c -> oracle/_build/liborc.so (gcc -O2)."""
    out_dir = os.path.join(_HERE, "_build")
    so = os.path.join(out_dir, "liborc.so")
    src = os.path.join(_HERE, "kernels.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        os.makedirs(out_dir, exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", so, src])
    return so


_lib = None


def _c():
    global _lib
    if _lib is None:
        try:
            _lib = ctypes.CDLL(build_c_kernels())
        except Exception:  # pragma: no cover - gcc is present in the image
            _lib = False
    return _lib


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


class _Csr:
    """CSR with int64 indices, contiguous, for the C kernels."""

    def __init__(self, A):
        A = sp.csr_matrix(A)
        A.sort_indices()
        self.shape = A.shape
        self.rowptr = np.ascontiguousarray(A.indptr, dtype=np.int64)
        self.col = np.ascontiguousarray(A.indices, dtype=np.int64)
        self.val = np.ascontiguousarray(A.data, dtype=np.float64)
        self._sp = A

    def matvec(self, x):
        x = np.ascontiguousarray(x, dtype=_RT)
        lib = _c()
        if not lib:
            return self._sp @ x
        y = _empty(self.shape[0])
        if _RT is np.float64:
            lib.orc_csr_matvec(ctypes.c_int64(self.shape[0]), _p(self.rowptr, ctypes.c_int64),
                               _p(self.col, ctypes.c_int64), _p(self.val, ctypes.c_double),
                               _p(x, ctypes.c_double), _p(y, ctypes.c_double))
        else:
            lib.orc_csr_matvec_ld(ctypes.c_int64(self.shape[0]), _p(self.rowptr, ctypes.c_int64),
                                  _p(self.col, ctypes.c_int64), _p(self.val, ctypes.c_double),
                                  _p(x, ctypes.c_longdouble), _p(y, ctypes.c_longdouble))
        return y


class _Matrix:
    """What the solvers see as ``A`` / ``C``: supports ``A @ v`` and shape
    (MATLAB sparse mtimes; the kernels only use size(A,1) and A*v)."""

    def __init__(self, A):
        self._csr = _Csr(A)
        self.shape = self._csr.shape

    def __matmul__(self, v):
        return self._csr.matvec(v)


def as_operator(A):
    if hasattr(A, "__matmul__") and not sp.issparse(A) and not isinstance(A, np.ndarray):
        return A  # already an operator (matrix-free A, reg_cpkrylov.m:40)
    return _Matrix(A)


# --------------------------------------------------------------------------
# util/SymGivens.m:1-29
# --------------------------------------------------------------------------
def _sign(v):
    return 1.0 if v > 0 else (-1.0 if v < 0 else 0.0)  # MATLAB sign: sign(0) = 0


def sym_givens(a, b):
    """[c,s,d] with [c s; s -c][a;b] = [d;0] (util/SymGivens.m:1-29)."""
    if b == 0:
        if a == 0:
            c = 1.0
        else:
            c = _f(_sign(a))
        s = 0.0
        d = abs(a)
    elif a == 0:
        c = 0.0
        s = _f(_sign(b))
        d = abs(b)
    elif abs(b) > abs(a):
        t = a / b
        s = _sign(b) / _sqrt(1 + t * t)
        c = s * t
        d = b / s
    else:
        t = b / a
        c = _sign(a) / _sqrt(1 + t * t)
        s = c * t
        d = a / c
    return c, s, d


# --------------------------------------------------------------------------
# ops/opLDL2.m
# --------------------------------------------------------------------------
class OpLDL2:
    """Operator for K_P^{-1}, K_P = [A B'; B C] (ops/opLDL2.m:60-92).

    The factorization P'*K_P*P = L*D*L' (opLDL2.m:82, MATLAB ldl) is passed in:
      L    : scipy sparse, unit lower triangular (diagonal may be stored or not)
      d, e : D(i,i) and D(i+1,i) (e[i] != 0 <=> 2x2 pivot starting at i)
      perm : int array, column k of P is the unit vector e_perm[k], i.e.
             (P'*x)[k] = x[perm[k]]  and  (P*w)[perm[k]] = w[k].
    """

    def __init__(self, A, B, C, L, d, e, perm, ru_stateful=False):
        A = sp.csr_matrix(A)
        B = sp.csr_matrix(B)
        C = sp.csr_matrix(C)
        nA, nC = A.shape[0], C.shape[0]
        if nA != A.shape[1] or nC != C.shape[1]:
            raise ValueError("First and last arguments must be square.")       # opLDL2.m:68-70
        if B.shape[1] != nA or B.shape[0] != nC:
            raise ValueError("Incompatible dimensions.")                        # opLDL2.m:73-75
        self.nA, self.nC, self.n = nA, nC, nA + nC
        self.shape = (self.n, self.n)
        self.K = _Csr(sp.bmat([[A, B.T], [B, C]], format="csr"))                # opLDL2.m:81
        self._K12 = _Csr(B.T)                                                   # op.A(1:n,n+1:n+m), :170
        self._K22 = _Csr(C)                                                     # op.A(n+1:n+m,n+1:n+m), :171
        Ls = sp.tril(sp.csr_matrix(L), k=-1, format="csr")
        self.Lstrict = _Csr(Ls)
        self.d = np.ascontiguousarray(d, dtype=np.float64)
        e = np.asarray(e, dtype=np.float64)
        self.e = np.ascontiguousarray(np.concatenate([e, np.zeros(self.n - e.size)]))
        self.perm = np.ascontiguousarray(perm, dtype=np.int64)
        assert self.d.size == self.n and self.perm.size == self.n
        # public properties, defaults from opLDL2.m:45-50
        self.nitref = 3
        self.itref_tol = 1.0e-8
        self.force_itref = False
        self.residual_update = False
        self.ru_stateful = ru_stateful
        self.Aty = _zeros(nA)                                                 # opLDL2.m:90
        self.Cy = _zeros(nC)                                                  # opLDL2.m:91
        self.rNorm = None
        self.napply = 0
        self.nsolve = 0

    # op.LDL = P * inv(L') * inv(D) * inv(L) * P'   (opLDL2.m:86), right to left
    def ldl_solve(self, x):
        self.nsolve += 1
        w = np.ascontiguousarray(x[self.perm], dtype=_RT)               # P' * x
        lib = _c()
        Ls = self.Lstrict
        if lib:
            i64, f64 = ctypes.c_int64, ctypes.c_double
            ext = _RT is not np.float64
            wt = ctypes.c_longdouble if ext else f64
            fwd = lib.orc_unit_lower_solve_ld if ext else lib.orc_unit_lower_solve
            dsl = lib.orc_block_diag_solve_ld if ext else lib.orc_block_diag_solve
            bwd = lib.orc_unit_lower_transpose_solve_ld if ext else lib.orc_unit_lower_transpose_solve
            fwd(i64(self.n), _p(Ls.rowptr, i64), _p(Ls.col, i64), _p(Ls.val, f64), _p(w, wt))    # L \ .
            dsl(i64(self.n), _p(self.d, f64), _p(self.e, f64), _p(w, wt))                        # D \ .
            bwd(i64(self.n), _p(Ls.rowptr, i64), _p(Ls.col, i64), _p(Ls.val, f64), _p(w, wt))    # L' \ .
        else:  # pragma: no cover - same loops in Python, small cases only
            w = _py_ldl(self.n, Ls, self.d, self.e, w)
        y = _empty(self.n)
        y[self.perm] = w                                                        # P * .
        return y

    def __matmul__(self, x):
        return self.multiply(np.asarray(x, dtype=_RT))

    def multiply(self, x):
        """ops/opLDL2.m:161-188."""
        self.napply += 1
        n, m = self.nA, self.nC
        if self.residual_update:
            y = self.ldl_solve(np.concatenate([x[:n] - self.Aty, x[n:] - self.Cy]))   # :164-165
        else:
            y = self.ldl_solve(x)                                               # :167
        if self.residual_update and self.ru_stateful:
            # :169-172.  Under Spot's value-class semantics these assignments are
            # lost when multiply returns (ru_stateful=False, the default).
            self.Aty = self._K12.matvec(y[n:])
            self.Cy = self._K22.matvec(y[n:])
        if self.nitref > 0:                                                     # :174
            r = x - self.K.matvec(y)
            rNorm = _f(np.linalg.norm(r))
            xNorm = _f(np.linalg.norm(x))
            nit = 0
            while nit < self.nitref and (rNorm >= self.itref_tol * xNorm or self.force_itref):
                dy = self.ldl_solve(r)
                y = y + dy
                r = x - self.K.matvec(y)
                rNorm = _f(np.linalg.norm(r))
                nit += 1
            self.rNorm = rNorm
        return y

    def divide(self, b):
        """ops/opLDL2.m:193-195: M \\ b = K_P * b."""
        return self.K.matvec(np.asarray(b, dtype=_RT))


def _py_ldl(n, Ls, d, e, w):  # pragma: no cover
    for i in range(n):
        for k in range(Ls.rowptr[i], Ls.rowptr[i + 1]):
            w[i] -= Ls.val[k] * w[Ls.col[k]]
    i = 0
    while i < n:
        if i + 1 < n and e[i] != 0:
            a, b, c = d[i], e[i], d[i + 1]
            det = a * c - b * b
            w0, w1 = w[i], w[i + 1]
            w[i] = (c * w0 - b * w1) / det
            w[i + 1] = (a * w1 - b * w0) / det
            i += 2
        else:
            w[i] = w[i] / d[i]
            i += 1
    for i in range(n - 1, -1, -1):
        for k in range(Ls.rowptr[i], Ls.rowptr[i + 1]):
            w[Ls.col[k]] -= Ls.val[k] * w[i]
    return w


# --------------------------------------------------------------------------
# option parsing shared by the solvers (isfield logic at the top of each .m)
# --------------------------------------------------------------------------
def _opt(opts, name, default):
    if opts is not None and name in opts:
        return opts[name]
    return default


class SolverError(RuntimeError):
    """error(...) / MException thrown inside a solver; ``identifier`` mirrors
    the MException id where the reference sets one."""

    def __init__(self, msg, identifier="", iteration=0, value=_f("nan")):
        super().__init__(msg)
        self.identifier = identifier
        self.iteration = iteration
        self.value = value


def _indef(k, beta, what, second=False, ident=""):
    where = "Iter %d, " % k + ("2nd Lanczos vec, " if second else "")
    return SolverError(where + "beta (before sqrt) = %.5g : %s" % (beta, what),
                       identifier=ident, iteration=k, value=beta)


_SPD_MSG = "preconditioner does not behave as a spd matrix."
_SOS_MSG = "preconditioner not second-order sufficient"


# --------------------------------------------------------------------------
# kernels/cpcg.m
# --------------------------------------------------------------------------
def cpcg(b, A, C, M, opts=None):
    A, C = as_operator(A), as_operator(C)
    n, m = A.shape[0], C.shape[0]
    atol = _opt(opts, "atol", 1.0e-6)
    rtol = _opt(opts, "rtol", 1.0e-6)
    itmax = _opt(opts, "itmax", n)

    x = _zeros(n)
    a = _zeros(m)
    w = _zeros(m)
    g = -np.asarray(b, dtype=_RT)

    ru = M @ np.concatenate([g, w]); r = ru[:n]; u = ru[n:]                     # :125
    p = -r
    q = -u

    residNorm2 = _f(g @ r)                                                   # :130
    residNorm = _sqrt(residNorm2) if residNorm2 >= 0 else _f("nan")
    stopTol = atol + rtol * residNorm
    residHistory = [residNorm]
    itn = 0

    while residNorm > stopTol and itn < itmax:                                  # :147
        itn += 1
        Ap = A @ p; pAp = _f(p @ Ap)
        Cq = C @ q; qCq = _f(q @ Cq)
        alpha = residNorm2 / (pAp + qCq)

        x = x + alpha * p
        a = a + alpha * q
        g = g + alpha * Ap
        w = w + alpha * Cq

        ru = M @ np.concatenate([g, w]); r = ru[:n]; u = ru[n:]                 # :166
        t = a + u
        residNorm2_new = _f(g @ r) + _f(t @ w)
        beta = residNorm2_new / residNorm2

        p = -r + beta * p
        q = -t + beta * q

        residNorm2 = residNorm2_new
        residNorm = _sqrt(residNorm2) if residNorm2 >= 0 else _f("nan")  # :175 (complex in MATLAB)
        residHistory.append(residNorm)

    flag = {"solved": bool(residNorm <= stopTol)}
    stats = {"niters": itn, "residHistory": np.array(residHistory)}
    return x, a, stats, flag


# --------------------------------------------------------------------------
# shared Lanczos start (cpcglanczos.m:153-171, cpminres.m:131-148, cpsymmlq.m:137-154)
# --------------------------------------------------------------------------
def _lanczos_start(b, n, m, M, msg, ident=""):
    u = np.asarray(b, dtype=_RT)
    t = _zeros(m)
    vprec = M @ np.concatenate([u, t])
    vkp1 = vprec[:n].copy()
    qkp1 = -vprec[n:]
    beta = _f(u @ vkp1)
    eps100 = 100 * _EPS
    if beta < -eps100:
        raise _indef(0, beta, msg, ident=ident)
    beta = _sqrt(abs(beta))
    if beta > 0:
        vkp1 = vkp1 / beta
        qkp1 = qkp1 / beta
    return vkp1, qkp1, beta


def _lanczos_step(A, C, M, n, vk, qk, vkm1, qkm1, beta, k, msg, ident="", second=False):
    """u=A*vk ... normalise (cpminres.m:187-206 and twins).  ``vkm1 is None``
    drops the beta*vkm1 term (cpsymmlq.m:202-204, second Lanczos vector)."""
    u = A @ vk
    t = C @ qk
    alpha = _f(u @ vk) + _f(t @ qk)
    vprec = M @ np.concatenate([u, -t])
    if vkm1 is None:
        vkp1 = vprec[:n] - alpha * vk
        qkp1 = qk - vprec[n:]
        qkp1 = qkp1 - alpha * qk
    else:
        vkp1 = vprec[:n] - alpha * vk - beta * vkm1
        qkp1 = qk - vprec[n:]
        qkp1 = qkp1 - alpha * qk - beta * qkm1
    beta = _f(u @ vkp1) + _f(t @ qkp1)
    if beta < -100 * _EPS:
        raise _indef(k, beta, msg, second=second, ident=ident)
    beta = _sqrt(abs(beta))
    if beta > 0:
        vkp1 = vkp1 / beta
        qkp1 = qkp1 / beta
    return u, t, alpha, vkp1, qkp1, beta


# --------------------------------------------------------------------------
# kernels/cpcglanczos.m
# --------------------------------------------------------------------------
def cpcglanczos(b, A, C, M, opts=None):
    A, C = as_operator(A), as_operator(C)
    n, m = A.shape[0], C.shape[0]
    atol = _opt(opts, "atol", 1.0e-6)
    rtol = _opt(opts, "rtol", 1.0e-6)
    btol = _opt(opts, "btol", 0.0)
    itmax = _opt(opts, "itmax", n)
    ident = "CPCGLanczos:IndefiniteError"

    x = _zeros(n)
    y = _zeros(m)
    vk = _zeros(n)
    qk = _zeros(m)
    oldbeta = 0.0
    opNorm2 = 0.0

    vkp1, qkp1, beta = _lanczos_start(b, n, m, M, _SOS_MSG, ident)
    wv = vkp1
    wq = qkp1
    beta1 = beta
    residNorm = beta1
    residHistory = [residNorm]

    k = 0
    dg = 0.0
    low = 1.0
    eta = beta
    rhobar = 1.0
    xxNorm2 = 0.0
    xNorm = 0.0
    tau = 0.0
    delta = 0.0

    stopTol = atol + rtol * residNorm
    bstopTol = btol * beta1

    while residNorm > stopTol and residNorm > bstopTol and k < itmax:           # :221
        k += 1
        vkm1, qkm1 = vk, qk
        vk, qk = vkp1, qkp1

        # the x/y update (:238-239) needs alpha only; keep statement order
        u = A @ vk
        t = C @ qk
        alpha = _f(u @ vk) + _f(t @ qk)
        dg = alpha - low * low * dg
        zeta = eta / dg
        x = x + zeta * wv
        y = y - zeta * wq

        vprec = M @ np.concatenate([u, -t])                                     # :242
        vkp1 = vprec[:n] - alpha * vk - beta * vkm1
        qkp1 = qk - vprec[n:]
        qkp1 = qkp1 - alpha * qk - beta * qkm1
        beta = _f(u @ vkp1) + _f(t @ qkp1)
        if beta < -100 * _EPS:
            raise _indef(k, beta, _SOS_MSG, ident=ident)
        beta = _sqrt(abs(beta))
        if beta > 0:
            vkp1 = vkp1 / beta
            qkp1 = qkp1 / beta

        low = beta / dg
        eta = -low * eta
        wv = vkp1 - low * wv
        wq = qkp1 - low * wq

        if btol > 0:                                                            # :271-291
            rho = _sqrt(rhobar * rhobar + low * low)
            cs = rhobar / rho
            sn = low / rho
            num = zeta - delta * tau
            taubar = num / rhobar
            tau = num / rho
            xNorm = _sqrt(xxNorm2 + taubar * taubar)
            xxNorm2 = xxNorm2 + tau * tau
            delta = sn
            rhobar = -cs
            opNorm2 = opNorm2 + alpha * alpha + beta * beta + oldbeta * oldbeta
            opNorm = _sqrt(opNorm2)
            bkerr = opNorm * xNorm + beta1
            bstopTol = btol * bkerr

        residNorm = beta * abs(zeta)
        residHistory.append(residNorm)
        oldbeta = beta

    stats = {"niters": k, "residHistory": np.array(residHistory)}
    flag = {"solved": False}
    stats["status"] = "maximum number of iterations attained"
    if residNorm <= stopTol:
        flag["solved"] = True
        stats["status"] = "residual small compared to initial residual"
    if btol > 0:
        if residNorm <= bstopTol:
            flag["solved"] = True
            stats["status"] = "backward error small"
    return x, y, stats, flag


# --------------------------------------------------------------------------
# kernels/cpminres.m
# --------------------------------------------------------------------------
def cpminres(b, A, C, M, opts=None):
    A, C = as_operator(A), as_operator(C)
    n, m = A.shape[0], C.shape[0]
    atol = _opt(opts, "atol", 1.0e-6)
    rtol = _opt(opts, "rtol", 1.0e-6)
    itmax = _opt(opts, "itmax", n)

    x = _zeros(n)
    y = _zeros(m)
    vk = _zeros(n)
    qk = _zeros(m)

    vkp1, qkp1, beta = _lanczos_start(b, n, m, M, _SPD_MSG)
    wv = vkp1
    wq = qkp1
    wv2 = _zeros(n)
    wq2 = _zeros(m)
    residNorm = beta
    residHistory = [residNorm]

    k = 0
    deltabar = 0.0
    epsln = 0.0
    taubar = beta
    cs = -1.0
    sn = 0.0
    stopTol = atol + rtol * residNorm

    while residNorm > stopTol and k < itmax:                                    # :176
        k += 1
        vkm1, qkm1 = vk, qk
        vk, qk = vkp1, qkp1

        u, t, alpha, vkp1, qkp1, beta = _lanczos_step(
            A, C, M, n, vk, qk, vkm1, qkm1, beta, k, _SPD_MSG)                  # :187-206

        oldeps = epsln                                                          # :211-215
        delta = cs * deltabar + sn * alpha
        gammabar = sn * deltabar - cs * alpha
        epsln = sn * beta
        deltabar = -cs * beta

        gamma = _hypot(gammabar, beta)                                      # norm([gammabar beta]) :218
        cs = gammabar / gamma
        sn = beta / gamma
        tau = cs * taubar
        taubar = sn * taubar

        wv1, wv2 = wv2, wv                                                      # :225-228
        wq1, wq2 = wq2, wq
        wv = (vk - oldeps * wv1 - delta * wv2) / gamma
        wq = (qk - oldeps * wq1 - delta * wq2) / gamma
        x = x + tau * wv
        y = y - tau * wq

        residNorm = taubar
        residHistory.append(residNorm)

    stats = {"niters": k, "residHistory": np.array(residHistory)}
    flag = {"solved": bool(residNorm <= stopTol)}
    return x, y, stats, flag


# --------------------------------------------------------------------------
# kernels/cpsymmlq.m
# --------------------------------------------------------------------------
def cpsymmlq(b, A, C, M, opts=None):
    A, C = as_operator(A), as_operator(C)
    n, m = A.shape[0], C.shape[0]
    atol = _opt(opts, "atol", 1.0e-6)
    rtol = _opt(opts, "rtol", 1.0e-6)
    itmax = _opt(opts, "itmax", n)
    b = np.asarray(b, dtype=_RT)

    x = _zeros(n)
    y = _zeros(m)
    wv = _zeros(n)
    wq = _zeros(m)
    k = 0

    vkp1, qkp1, beta1 = _lanczos_start(b, n, m, M, _SPD_MSG)
    cgresidNorm = beta1
    stopTol = atol + rtol * cgresidNorm

    if cgresidNorm <= stopTol:                                                  # :161-170
        lqresidNorm = beta1
        qrresidNorm = beta1
        cgresidHistory = [cgresidNorm]
        lqresidHistory = [lqresidNorm]
        qrresidHistory = [qrresidNorm]
    else:
        cgresidHistory, lqresidHistory, qrresidHistory = [], [], []

    done = cgresidNorm <= stopTol
    if not done:
        vk, qk = vkp1, qkp1
        u, t, alpha, vkp1, qkp1, beta = _lanczos_step(
            A, C, M, n, vk, qk, None, None, 0.0, 0, _SPD_MSG, second=True)      # :198-216

        gammabar = alpha                                                        # :219-225
        deltabar = beta
        epsdelzeta = beta1
        epsilonzeta = 0.0
        bstep = 0.0
        snprod = 1.0
        matnorm2 = alpha * alpha + beta * beta

        while cgresidNorm > stopTol and k < itmax:                              # :229
            matnorm = _sqrt(matnorm2)
            epsmat = matnorm * _EPS
            den = gammabar
            if den == 0:
                den = epsmat
            lqresidNorm = _hypot(epsdelzeta, epsilonzeta)
            qrresidNorm = snprod * beta1
            cgresidNorm = qrresidNorm * beta / abs(den)
            lqresidHistory.append(lqresidNorm)
            qrresidHistory.append(qrresidNorm)
            cgresidHistory.append(cgresidNorm)

            k += 1
            zetabar = epsdelzeta / den                                          # :255-256 (zeta is dead)

            vkm1, qkm1 = vk, qk
            vk, qk = vkp1, qkp1
            betaold = beta
            u, t, alpha, vkp1, qkp1, beta = _lanczos_step(
                A, C, M, n, vk, qk, vkm1, qkm1, beta, k, _SPD_MSG)              # :266-285

            matnorm2 = matnorm2 + alpha * alpha + beta * beta + betaold * betaold

            gamma = _hypot(gammabar, betaold)                               # :291-297
            cs = gammabar / gamma
            sn = betaold / gamma
            delta = cs * deltabar + sn * alpha
            gammabar = sn * deltabar - cs * alpha
            epsilon = sn * beta
            deltabar = -cs * beta

            zeta = epsdelzeta / gamma                                           # :300-306
            zcs = zeta * cs
            zsn = zeta * sn
            x = x + zcs * wv + zsn * vk
            y = y - zcs * wq - zsn * qk
            wv = sn * wv - cs * vk
            wq = sn * wq - cs * qk

            bstep = bstep + snprod * cs * zeta                                  # :310-313
            snprod = snprod * sn
            epsdelzeta = epsilonzeta - delta * zeta
            epsilonzeta = -epsilon * zeta

        matnorm = _sqrt(matnorm2)                                           # :318-327
        epsmat = matnorm * _EPS
        den = gammabar
        if den == 0:
            den = epsmat
        lqresidNorm = _hypot(epsdelzeta, epsilonzeta)
        qrresidNorm = snprod * beta1
        lqresidHistory.append(lqresidNorm)
        qrresidHistory.append(qrresidNorm)
        cgresidHistory = [beta1] + cgresidHistory                               # :331

        if cgresidNorm < lqresidNorm:                                           # :334-339
            zetabar = epsdelzeta / den
            bstep = bstep + snprod * zetabar
            x = x + zetabar * wv
            y = y - zetabar * wq

        vprec = M @ np.concatenate([b, _zeros(m)])                            # :342-347
        vk = vprec[:n]
        qk = -vprec[n:]
        bstep = bstep / beta1
        x = x + bstep * vk
        y = y - bstep * qk

    stats = {"niters": k,
             "lqresidHistory": np.array(lqresidHistory),
             "qrresidHistory": np.array(qrresidHistory),
             "cgresidHistory": np.array(cgresidHistory)}
    flag = {"solved": bool(cgresidNorm <= stopTol)}
    return x, y, stats, flag


# --------------------------------------------------------------------------
# kernels/cpgmres.m
# --------------------------------------------------------------------------
def _sqrt_real(v):
    # sqrt(real(.)) fallback of cpgmres.m:174-176,220-222: MATLAB's sqrt of a
    # negative real is complex and real(sqrt(v)) = 0; the guard recomputes
    # sqrt(real(v)), which is the same complex number, so the value kept is
    # purely imaginary.  A negative P-inner product is a breakdown; return NaN
    # so that the caller's tests fail the same way (NaN > stopTol is false).
    return _sqrt(v) if v >= 0 else _f("nan")


def cpgmres(b, A, C, M, opts=None):
    A, C = as_operator(A), as_operator(C)
    n, m = A.shape[0], C.shape[0]
    atol = _opt(opts, "atol", 1.0e-6)
    rtol = _opt(opts, "rtol", 1.0e-6)
    restart = int(_opt(opts, "restart", 50))
    itmax = _opt(opts, "itmax", n + m)
    b = np.asarray(b, dtype=_RT)

    g = _zeros(restart + 1)
    V = _zeros((n, restart + 1), order="F")
    Q = _zeros((m, restart + 1), order="F")
    H = _zeros((restart + 1, restart))
    c = _zeros(restart)
    s = _zeros(restart)

    x = _zeros(n)
    y = _zeros(m)

    finished = False
    outer = 0
    outermax = int(math.ceil(itmax / restart))
    residHistory = []
    residNorm = _f("nan")
    stopTol = _f("nan")
    k = 0

    while (not finished) and outer < outermax:                                  # :155
        outer += 1
        q = _zeros(m)
        if outer == 1:
            u = b
            t = _zeros(m)
            w = M @ np.concatenate([u, -t])
            V[:, 0] = w[:n]
            Q[:, 0] = -w[n:]
        else:
            u = b - A @ x
            t = C @ y
            w = M @ np.concatenate([u, -t])
            V[:, 0] = w[:n]
            Q[:, 0] = y - w[n:]
        residNorm = _sqrt_real(_f(u @ V[:, 0]) + _f(t @ Q[:, 0]))         # :173
        if residNorm != 0:
            V[:, 0] = V[:, 0] / residNorm
            Q[:, 0] = Q[:, 0] / residNorm
        if outer == 1:
            stopTol = atol + rtol * residNorm
            residHistory = [residNorm]

        k = 0
        g[0] = residNorm

        while residNorm > stopTol and k < restart:                              # :203
            k += 1
            u = A @ V[:, k - 1]
            t = C @ Q[:, k - 1]
            w = M @ np.concatenate([u, -t])
            V[:, k] = w[:n]
            Q[:, k] = Q[:, k - 1] - w[n:]
            for j in range(1, k + 1):
                H[j - 1, k - 1] = _f(V[:, j - 1] @ u) + _f(Q[:, j - 1] @ t)
                V[:, k] = V[:, k] - H[j - 1, k - 1] * V[:, j - 1]
                Q[:, k] = Q[:, k] - H[j - 1, k - 1] * Q[:, j - 1]
            H[k, k - 1] = _sqrt_real(_f(u @ V[:, k]) + _f(t @ Q[:, k]))   # :219
            if H[k, k - 1] != 0:
                V[:, k] = V[:, k] / H[k, k - 1]
                Q[:, k] = Q[:, k] / H[k, k - 1]

            for j in range(1, k):                                               # :229-234
                Hjk = c[j - 1] * H[j - 1, k - 1] + s[j - 1] * H[j, k - 1]
                H[j, k - 1] = s[j - 1] * H[j - 1, k - 1] - c[j - 1] * H[j, k - 1]
                H[j - 1, k - 1] = Hjk

            c[k - 1], s[k - 1], H[k - 1, k - 1] = sym_givens(H[k - 1, k - 1], H[k, k - 1])
            H[k, k - 1] = 0
            g[k] = s[k - 1] * g[k - 1]
            g[k - 1] = c[k - 1] * g[k - 1]
            residNorm = abs(g[k])
            residHistory.append(residNorm)

        # z = H(1:k,1:k) \ g(1:k): upper-triangular back substitution (:257)
        z = _zeros(k)
        for i in range(k - 1, -1, -1):
            acc = g[i]
            for j in range(i + 1, k):
                acc -= H[i, j] * z[j]
            z[i] = acc / H[i, i]
        x = x + V[:, :k] @ z
        q = q + Q[:, :k] @ z
        y = y - q

        finished = residNorm <= stopTol

    stats = {"niters": (outer - 1) * restart + k, "residHistory": np.array(residHistory)}
    flag = {"solved": bool(residNorm <= stopTol)}
    return x, y, stats, flag


# --------------------------------------------------------------------------
# kernels/cpdqgmres.m
# --------------------------------------------------------------------------
def cpdqgmres(b, A, C, M, opts=None):
    A, C = as_operator(A), as_operator(C)
    n, m = A.shape[0], C.shape[0]
    atol = _opt(opts, "atol", 1.0e-6)
    rtol = _opt(opts, "rtol", 1.0e-6)
    itmax = int(_opt(opts, "itmax", n + m))
    mem = 50
    if opts is not None and "mem" in opts:
        mem = max(1, int(opts["mem"]))
    b = np.asarray(b, dtype=_RT)

    mem = max(1, min(mem, itmax))                                               # :125 (itmax=0 would give empty arrays)
    g = _zeros(mem + 1)
    V = _zeros((n, mem + 1), order="F")
    Q = _zeros((m, mem + 1), order="F")
    PV = _zeros((n, mem + 1), order="F")
    PQ = _zeros((m, mem + 1), order="F")
    c = _zeros(max(mem, 1))
    s = _zeros(max(mem, 1))
    # H(j, 2+k-j) band storage (cpdqgmres.m:133); rows grown on demand instead
    # of zeros(itmax, mem+2) -- same values, row index j is absolute.
    Hrows = {}

    def Hget(j, kk):
        row = Hrows.get(j)
        return 0.0 if row is None else _f(row[kk])

    def Hset(j, kk, v):
        if j not in Hrows:
            Hrows[j] = _zeros(mem + 3)
        Hrows[j][kk] = v

    x = _zeros(n)
    y = _zeros(m)
    u = b
    t = _zeros(m)

    w = M @ np.concatenate([u, t])                                              # :151
    V[:, 0] = w[:n]
    Q[:, 0] = -w[n:]
    residNorm = _sqrt_real(_f(u @ V[:, 0]))
    if residNorm != 0:
        V[:, 0] = V[:, 0] / residNorm
        Q[:, 0] = Q[:, 0] / residNorm

    k = 0
    g[0] = residNorm
    stopTol = atol + rtol * residNorm
    residHistory = [residNorm]

    while residNorm > stopTol and k < itmax:                                    # :194
        k += 1
        kpos = (k - 1) % (mem + 1)          # 0-based slot of k      (:199)
        kp1pos = k % (mem + 1)              # 0-based slot of k+1    (:200)
        rotpos = (k - 1) % mem              # 0-based rotation slot  (:201)

        u = A @ V[:, kpos]
        t = C @ Q[:, kpos]
        w = M @ np.concatenate([u, -t])
        V[:, kp1pos] = w[:n]
        Q[:, kp1pos] = Q[:, kpos] - w[n:]
        for j in range(max(1, k - mem + 1), k + 1):                             # :210-216
            jpos = (j - 1) % (mem + 1)
            kk = 2 + k - j
            h = _f(V[:, jpos] @ u) + _f(Q[:, jpos] @ t)
            Hset(j, kk, h)
            V[:, kp1pos] = V[:, kp1pos] - h * V[:, jpos]
            Q[:, kp1pos] = Q[:, kp1pos] - h * Q[:, jpos]
        hk1 = _sqrt_real(_f(u @ V[:, kp1pos]) + _f(t @ Q[:, kp1pos]))     # :218
        Hset(k, 1, hk1)
        if hk1 != 0:
            V[:, kp1pos] = V[:, kp1pos] / hk1
            Q[:, kp1pos] = Q[:, kp1pos] / hk1

        for j in range(max(1, k - mem), k):                                     # :228-235
            jrotpos = (j - 1) % mem
            kk = k - j + 1
            kk1 = kk + 1
            Hjk = c[jrotpos] * Hget(j, kk1) + s[jrotpos] * Hget(j + 1, kk)
            Hset(j + 1, kk, s[jrotpos] * Hget(j, kk1) - c[jrotpos] * Hget(j + 1, kk))
            Hset(j, kk1, Hjk)

        c[rotpos], s[rotpos], hk2 = sym_givens(Hget(k, 2), Hget(k, 1))          # :243
        Hset(k, 2, hk2)
        Hset(k, 1, 0.0)
        g[kp1pos] = s[rotpos] * g[kpos]
        g[kpos] = c[rotpos] * g[kpos]

        PV[:, kpos] = V[:, kpos]                                                # :253-265
        PQ[:, kpos] = Q[:, kpos]
        for j in range(max(1, k - mem), k):
            jpos = (j - 1) % (mem + 1)
            kk = 2 + k - j
            PV[:, kpos] = PV[:, kpos] - Hget(j, kk) * PV[:, jpos]
            PQ[:, kpos] = PQ[:, kpos] - Hget(j, kk) * PQ[:, jpos]
        PV[:, kpos] = PV[:, kpos] / Hget(k, 2)
        PQ[:, kpos] = PQ[:, kpos] / Hget(k, 2)
        x = x + g[kpos] * PV[:, kpos]
        y = y - g[kpos] * PQ[:, kpos]

        residNorm = abs(g[kp1pos])
        residHistory.append(residNorm)
        Hrows.pop(k - mem - 2, None)        # rows older than the band are dead

    stats = {"niters": k, "residHistory": np.array(residHistory)}
    flag = {"solved": bool(residNorm <= stopTol)}
    return x, y, stats, flag


SOLVERS = {
    "cpcg": cpcg,
    "cpcglanczos": cpcglanczos,
    "cpminres": cpminres,
    "cpsymmlq": cpsymmlq,
    "cpgmres": cpgmres,
    "cpdqgmres": cpdqgmres,
}


# --------------------------------------------------------------------------
# reg_cpkrylov.m:122-178
# --------------------------------------------------------------------------
def reg_cpkrylov(method, b, A, B, C, G, opts=None, factor=None, ru_stateful=False):
    """[x, stats, flag] = reg_cpkrylov(method, b, A, B, C, G, opts).

    ``factor(K_P) -> (L, d, e, perm)`` plays the role of MATLAB's ldl inside the
    opLDL2 constructor (ops/opLDL2.m:82); it is an argument because the test
    must hand the *same* factors to the oracle and to the CUDA path.
    """
    import time
    if callable(method):
        solver = method
    else:
        solver = SOLVERS[method]
    b = np.asarray(b, dtype=_RT).ravel()
    Aop = as_operator(A)
    B = sp.csr_matrix(B)
    C = sp.csr_matrix(C)
    G = sp.csr_matrix(G)

    t0 = time.perf_counter()
    n = Aop.shape[0]
    m = B.shape[0]
    KP = sp.bmat([[G, B.T], [B, -C]], format="csc")
    L, d, e, perm = factor(KP)
    M = OpLDL2(G, B, -C, L, d, e, perm, ru_stateful=ru_stateful)                # :131
    ptime = time.perf_counter() - t0

    if opts is not None:                                                        # :135-148
        if "nitref" in opts:
            M.nitref = max(0, int(round(opts["nitref"])))
        if "itref_tol" in opts:
            M.itref_tol = opts["itref_tol"]
        if "residual_update" in opts:
            M.residual_update = bool(opts["residual_update"])
        if "force_itref" in opts:
            M.force_itref = bool(opts["force_itref"])

    t1 = time.perf_counter()
    shift = False
    if np.any(b[n:n + m]):                                                      # :154
        shift = True
        xy0 = M @ np.concatenate([_zeros(n), b[n:n + m]])
        BT = _Csr(B.T)
        b1 = b[:n] - Aop @ xy0[:n] - BT.matvec(xy0[n:])
    else:
        b1 = b[:n]

    dx, dy, stats, flag = solver(b1, Aop, as_operator(C), M, opts)              # :163

    if shift:
        x1 = xy0[:n] + dx
        x2 = xy0[n:] + dy
    else:
        x1, x2 = dx, dy
    x = np.concatenate([x1, x2])
    stats["ptime"] = ptime
    stats["stime"] = time.perf_counter() - t1
    stats["napply"] = M.napply
    stats["nsolve"] = M.nsolve
    return x, stats, flag


def ldl_static_perm(K, perm):
    """Reference for the device-side factorization (cpk_ldl2_create_sqd): LDL' of
    K[perm][:, perm] WITHOUT pivoting (natural order, diagonal pivots only) through
    SuperLU's symmetric mode, as `ldl_superlu` does for its own ordering.  Returns
    (L unit lower triangular CSC, d).  TEST INFRASTRUCTURE, like the rest of this file."""
    import scipy.sparse.linalg as spla
    perm = np.asarray(perm, dtype=np.int64)
    Kp = sp.csc_matrix(K)[perm][:, perm].tocsc()
    lu = spla.splu(Kp, permc_spec="NATURAL", diag_pivot_thresh=0.0, options={"SymmetricMode": True})
    N = Kp.shape[0]
    if not (np.array_equal(lu.perm_r, np.arange(N)) and np.array_equal(lu.perm_c, np.arange(N))):
        raise RuntimeError("SuperLU left the natural order: the matrix is not strongly factorizable in this ordering")
    d = lu.U.diagonal().copy()
    return sp.csc_matrix(lu.L), d
