"""Host-side mirrors of the reference's operator objects, holding device handles.

opLDL2     <-> ops/opLDL2.m  (constructor :60-92, multiply :161-188, divide :193-195,
                              public properties :45-50, transpose/ctranspose :120-136,
                              double :138-149)
KktSystem  <-> the (A, C, M) triple every solver receives (e.g. kernels/cpminres.m:1)
"""
from __future__ import annotations

import ctypes as ct
import time

import numpy as np
import scipy.sparse as sp

from . import _lib
from .ldl import ldl_factor


def _vec(x, n, name):
    x = np.ascontiguousarray(np.asarray(x, dtype=np.float64).reshape(-1))
    if x.size != n:
        raise ValueError("%s must have %d entries, got %d" % (name, n, x.size))
    return x


class opLDL2:
    """Operator for multiplication by inv([A B'; B C]) through its LDL' factors,
    with optional residual update and iterative refinement (ops/opLDL2.m).

    ``opLDL2(A, B, C)`` factorizes on the host (untimed setup, like MATLAB's
    ``ldl`` in opLDL2.m:82) unless ``factors=(L, d, e, perm)`` are given, and
    uploads everything once.  ``M @ z`` / ``M * z`` runs on the GPU.
    """

    def __init__(self, A, B, C, factors=None, ldl_method="auto", device=0, perm=None):
        A = sp.csc_matrix(A); B = sp.csc_matrix(B); C_ = sp.csc_matrix(C)
        nA, nC = A.shape[0], C_.shape[0]
        if nA != A.shape[1] or nC != C_.shape[1]:
            raise ValueError("First and last arguments must be square.")        # opLDL2.m:68-70
        if B.shape[1] != nA or B.shape[0] != nC:
            raise ValueError("Incompatible dimensions.")                         # opLDL2.m:73-75
        self.nA, self.nC, self.n = nA, nC, nA + nC
        self.shape = (self.n, self.n)
        t0 = time.perf_counter()
        if isinstance(factors, str) and factors == "device":
            # symmetric quasi-definite K_P, static permutation: numeric LDL' on the device
            # (cpk_ldl2_create_sqd); `refactor` then serves the next systems of the sequence
            if perm is None:
                raise ValueError("factors='device' needs the fill-reducing permutation `perm` "
                                 "(e.g. cpkrylov_b200.ldl.static_perm(K_P))")
            self.t_factor = 0.0
            self.factors = None
            perm = np.ascontiguousarray(perm, dtype=np.int64)
            keep = [_lib.Csc(A), _lib.Csc(B), _lib.Csc(C_)]
            h = ct.c_uint64(0)
            _lib.check(_lib.lib().cpk_ldl2_create_sqd(ct.byref(h), *[k.ref() for k in keep],
                                                      perm.ctypes.data_as(ct.POINTER(ct.c_int64)), int(device)))
            self.t_upload = time.perf_counter() - t0
            self.handle = h
            self.device = device
            self._owned_by_system = False
            self._keep = None
            self._nitref, self._itref_tol, self._force_itref, self._residual_update = 3, 1.0e-8, False, False
            self._ru_stateful = False
            self.last_stats = None
            return
        if factors is None:
            K = sp.bmat([[A, B.T], [B, C_]], format="csc")                       # opLDL2.m:81
            factors = ldl_factor(K, ldl_method)                                  # opLDL2.m:82
        self.t_factor = time.perf_counter() - t0
        L, d, e, perm = factors
        self.factors = factors
        N = self.n
        d = np.asarray(d, dtype=np.float64); e = np.asarray(e, dtype=np.float64)
        e = np.concatenate([e, np.zeros(N - e.size)])[:N]
        D = sp.diags([d, e[:N - 1], e[:N - 1]], [0, -1, 1], shape=(N, N), format="csc")
        D.eliminate_zeros()
        t1 = time.perf_counter()
        self._keep = [_lib.Csc(A), _lib.Csc(B), _lib.Csc(C_), _lib.Csc(L), _lib.Csc(D)]
        perm = np.ascontiguousarray(perm, dtype=np.int64)
        h = ct.c_uint64(0)
        _lib.check(_lib.lib().cpk_ldl2_create(ct.byref(h), *[k.ref() for k in self._keep],
                                              perm.ctypes.data_as(ct.POINTER(ct.c_int64)), int(device)))
        self._keep = None
        self.t_upload = time.perf_counter() - t1
        self.handle = h
        self.device = device
        self._owned_by_system = False
        # public properties (opLDL2.m:45-50)
        self._nitref, self._itref_tol, self._force_itref, self._residual_update = 3, 1.0e-8, False, False
        self._ru_stateful = False
        self.last_stats = None

    # -- properties with the reference's setter semantics ----------------------
    @property
    def nitref(self):
        return self._nitref

    @nitref.setter
    def nitref(self, val):
        self._nitref = max(0, int(round(val)))                                   # opLDL2.m:97-99
        _lib.check(_lib.lib().cpk_ldl2_set_nitref(self.handle, float(val)))

    @property
    def itref_tol(self):
        return self._itref_tol

    @itref_tol.setter
    def itref_tol(self, val):
        self._itref_tol = float(val)     # opLDL2.m:101 ("sef.") never installs a setter
        _lib.check(_lib.lib().cpk_ldl2_set_itref_tol(self.handle, float(val)))

    @property
    def force_itref(self):
        return self._force_itref

    @force_itref.setter
    def force_itref(self, val):
        self._force_itref = bool(val) if val in (0, 1, True, False) else False   # opLDL2.m:105-111
        _lib.check(_lib.lib().cpk_ldl2_set_force_itref(self.handle, int(self._force_itref)))

    @property
    def residual_update(self):
        return self._residual_update

    @residual_update.setter
    def residual_update(self, val):
        self._residual_update = bool(val)
        _lib.check(_lib.lib().cpk_ldl2_set_residual_update(self.handle, int(bool(val))))

    @property
    def ru_stateful(self):
        return self._ru_stateful

    @ru_stateful.setter
    def ru_stateful(self, val):
        self._ru_stateful = bool(val)
        _lib.check(_lib.lib().cpk_ldl2_set_ru_stateful(self.handle, int(bool(val))))

    def set_track_rnorm(self, val):
        _lib.check(_lib.lib().cpk_ldl2_set_track_rnorm(self.handle, int(bool(val))))

    @property
    def rNorm(self):
        v = ct.c_double(0.0)
        _lib.check(_lib.lib().cpk_ldl2_get_rnorm(self.handle, ct.byref(v)))
        return v.value

    def info(self):
        vals = [ct.c_int64(0) for _ in range(4)]
        _lib.check(_lib.lib().cpk_ldl2_info(self.handle, *[ct.byref(v) for v in vals]))
        return dict(zip(["nnz_L_off", "levels_fwd", "levels_bwd", "n_2x2"], [v.value for v in vals]))

    # -- operator contract ------------------------------------------------------
    def multiply(self, z):
        """y = M*z (opLDL2.m:161-188)."""
        z = _vec(z, self.n, "z")
        y = np.empty(self.n)
        st = _lib.StatsStruct()
        _lib.check(_lib.lib().cpk_ldl2_apply(self.handle, z.ctypes.data, y.ctypes.data, _lib.MEM_HOST, ct.byref(st)))
        self.last_stats = _lib.stats_to_dict(st)
        return y

    __matmul__ = multiply

    def __mul__(self, z):
        return self.multiply(z)

    def divide(self, b):
        """M \\ b = K_P * b (opLDL2.m:193-195)."""
        b = _vec(b, self.n, "b")
        y = np.empty(self.n)
        st = _lib.StatsStruct()
        _lib.check(_lib.lib().cpk_ldl2_matvec(self.handle, b.ctypes.data, y.ctypes.data, _lib.MEM_HOST, ct.byref(st)))
        self.last_stats = _lib.stats_to_dict(st)
        return y

    def transpose(self):                                                         # opLDL2.m:120-122
        return self

    ctranspose = transpose                                                       # real data: opLDL2.m:134-136
    conj = transpose

    @property
    def T(self):
        return self

    def double(self):                                                            # opLDL2.m:138-149
        X = np.zeros((self.n, self.n))
        e = np.zeros(self.n)
        for i in range(self.n):
            e[i] = 1.0
            X[:, i] = self.multiply(e)
            e[i] = 0.0
        return X

    # -- sequences with a fixed pattern ---------------
    def refactor(self, A, B, C):
        """New values, same sparsity patterns: numeric LDL' on the device, in place.
        Only for operators created with ``factors="device"``."""
        keep = [_lib.Csc(sp.csc_matrix(A)), _lib.Csc(sp.csc_matrix(B)), _lib.Csc(sp.csc_matrix(C))]
        t0 = time.perf_counter()
        _lib.check(_lib.lib().cpk_ldl2_refactor(self.handle, *[k.ref() for k in keep]))
        self.t_refactor = time.perf_counter() - t0

    def device_factor(self):
        """(L, d): unit lower triangular L (CSC, pattern with fill) and diag(D) as the device holds them."""
        L_ = _lib.lib()
        nnz = ct.c_int64(0)
        _lib.check(L_.cpk_ldl2_get_factor(self.handle, ct.byref(nnz), None, None, None, None))
        N = self.n
        colptr = np.zeros(N + 1, dtype=np.int64); rowind = np.zeros(max(nnz.value, 1), dtype=np.int64)
        val = np.zeros(max(nnz.value, 1)); d = np.zeros(N)
        _lib.check(L_.cpk_ldl2_get_factor(self.handle, ct.byref(nnz), colptr.ctypes.data_as(ct.POINTER(ct.c_int64)),
                                          rowind.ctypes.data_as(ct.POINTER(ct.c_int64)),
                                          val.ctypes.data_as(ct.POINTER(ct.c_double)), d.ctypes.data_as(ct.POINTER(ct.c_double))))
        L = sp.csc_matrix((val[:nnz.value], rowind[:nnz.value], colptr), shape=(N, N)) + sp.identity(N, format="csc")
        return L, d

    def close(self):
        cached = getattr(self, "_system", None)
        if cached is not None and not getattr(cached, "owns_M", True):
            self._system = None
            cached.close()              # the device copy of (A, C) a solver call cached on this operator
        if getattr(self, "handle", None) is not None and not self._owned_by_system:
            try:
                _lib.lib().cpk_destroy(self.handle)
            except Exception:
                pass
            self.handle = None

    def __del__(self):
        self.close()


class KktSystem:
    """Device-resident (A, C, M) of ``method(b1, A, C, M, opts)``."""

    def __init__(self, A, Cm, M: opLDL2, owns_M=True):
        """``owns_M``: closing the system also destroys M (the driver and the batch solver create
        M themselves); a solver called as ``method(b1, A, C, M, opts)`` passes False -- M is the
        caller's and outlives the system."""
        if not isinstance(M, opLDL2):
            raise TypeError("M must be a cpkrylov_b200.opLDL2 (GPU operator handle)")
        Cm = sp.csc_matrix(Cm)
        self.matrix_free = not (sp.issparse(A) or isinstance(A, np.ndarray))
        self.n, self.m = int(A.shape[0]), Cm.shape[0]
        self.N = self.n + self.m
        self.M = M
        c = _lib.Csc(Cm)
        h = ct.c_uint64(0)
        if self.matrix_free:
            # reg_cpkrylov.m:40: "A may be a matrix or a linear operator" -- anything with
            # A @ v (or A.matvec / A * v as a Spot operator has it); the product is a host
            # callback answered from inside the solve call (cpk_system_create_op)
            self._Aop = A
            self._cb_error = None
            self.n_products = 0

            def _cb(_ctx, v_ptr, u_ptr, n):
                try:
                    v = np.ctypeslib.as_array(v_ptr, shape=(n,))
                    u = np.ctypeslib.as_array(u_ptr, shape=(n,))
                    if hasattr(A, "matvec"):
                        r = A.matvec(v)
                    elif hasattr(A, "__matmul__"):
                        r = A @ v
                    else:
                        r = A * v
                    u[:] = np.asarray(r, dtype=np.float64).reshape(n)
                    self.n_products += 1
                    return 0
                except BaseException as ex:        # never unwind through the C frames
                    self._cb_error = ex
                    return 1
            self._cb = _lib.MATVEC_FN(_cb)          # keep the thunk alive as long as the system
            _lib.check(_lib.lib().cpk_system_create_op(ct.byref(h), self.n, self._cb, None, c.ref(), M.handle))
        else:
            a = _lib.Csc(sp.csc_matrix(A))
            _lib.check(_lib.lib().cpk_system_create(ct.byref(h), a.ref(), c.ref(), M.handle))
        self.handle = h
        self.owns_M = bool(owns_M)
        M._owned_by_system = True

    def check_callback(self):
        """re-raises what the A*v callback raised during the last solve (the C side only sees 'failed')"""
        ex = getattr(self, "_cb_error", None)
        if ex is not None:
            self._cb_error = None
            raise ex

    def update(self, A, Cm):
        """New values of A (= H) and C, same patterns (next system of a sequence)."""
        a, c = _lib.Csc(sp.csc_matrix(A)), _lib.Csc(sp.csc_matrix(Cm))
        _lib.check(_lib.lib().cpk_system_update(self.handle, a.ref(), c.ref()))

    def matvec(self, which, x):
        nn = self.n if which == 0 else self.m
        x = _vec(x, nn, "x")
        y = np.empty(nn)
        st = _lib.StatsStruct()
        _lib.check(_lib.lib().cpk_system_matvec(self.handle, which, x.ctypes.data, y.ctypes.data, _lib.MEM_HOST, ct.byref(st)))
        self.last_stats = _lib.stats_to_dict(st)
        return y

    def close(self):
        if getattr(self, "handle", None) is not None:
            try:
                _lib.lib().cpk_destroy(self.handle)
                self.M._owned_by_system = False
                if self.owns_M:
                    self.M.close()
            except Exception:
                pass
            self.handle = None

    def __del__(self):
        self.close()
