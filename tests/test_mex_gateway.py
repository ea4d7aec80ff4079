"""The MEX gateway (matlab/cpk_b200_mex.cpp) meets a compiler and is driven end to end.

No MATLAB exists in the build image, so the gateway is compiled against the prototype header
matlab/stub/mex.h and linked with the toy runtime matlab/stub/mex_stub.cpp into
tests/_build/libcpk_mex_stub.so; the test then builds mxArrays through the runtime's C helpers
and calls `mexFunction` the way MATLAB would (reg_cpkrylov_gpu.m is the caller in production).

CPU: syntax + link + error paths (the reference's texts).  GPU (-m gpu): opLDL2 apply and a
cpsymmlq reg_solve through the gateway, all three histories against the ctypes path
(reg_cpkrylov.m:1,109-117, kernels/cpsymmlq.m:363-366).
"""
import ctypes as ct
import os
import subprocess

import numpy as np
import pytest
import scipy.sparse as sp

from helpers import EX_OPTS, kp_of, load_factors, load_system

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "tests", "_build")
SO = os.path.join(BUILD, "libcpk_mex_stub.so")


def _build():
    from cpkrylov_b200 import _lib
    os.makedirs(BUILD, exist_ok=True)
    src = [os.path.join(ROOT, "matlab", "cpk_b200_mex.cpp"), os.path.join(ROOT, "matlab", "stub", "mex_stub.cpp")]
    inc = ["-I" + os.path.join(ROOT, "matlab", "stub"), "-I" + os.path.join(ROOT, "include")]
    libdir = os.path.dirname(_lib.LIB_PATH)
    newest = max(os.path.getmtime(f) for f in src + [os.path.join(ROOT, "matlab", "stub", "mex.h")])
    if not os.path.exists(SO) or os.path.getmtime(SO) < newest:
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-fPIC", "-shared", "-o", SO] + src + inc +
                              ["-L" + libdir, "-l:" + os.path.basename(_lib.LIB_PATH), "-Wl,-rpath," + libdir])
    _lib.lib()                                  # the C ABI library first (same process-wide registry)
    L = ct.CDLL(SO)
    P = ct.c_void_p
    L.stub_dense.restype = P; L.stub_dense.argtypes = [ct.c_size_t, ct.c_size_t, P]
    L.stub_sparse.restype = P; L.stub_sparse.argtypes = [ct.c_size_t, ct.c_size_t, P, P, P]
    L.stub_string.restype = P; L.stub_string.argtypes = [ct.c_char_p]
    L.stub_rows.restype = ct.c_size_t; L.stub_rows.argtypes = [P]
    L.stub_cols.restype = ct.c_size_t; L.stub_cols.argtypes = [P]
    L.stub_data.restype = ct.POINTER(ct.c_double); L.stub_data.argtypes = [P]
    L.stub_free.argtypes = [P]
    L.stub_last_id.restype = ct.c_char_p; L.stub_last_msg.restype = ct.c_char_p
    L.stub_call.restype = ct.c_int; L.stub_call.argtypes = [ct.c_int, ct.POINTER(P), ct.c_int, ct.POINTER(P)]
    return L


class Mex:
    """[out...] = cpk_b200_mex(cmd, args...)"""

    def __init__(self):
        self.L = _build()

    def _in(self, a):
        L = self.L
        if isinstance(a, str):
            return L.stub_string(a.encode())
        if sp.issparse(a):
            a = sp.csc_matrix(a); a.sort_indices()
            jc = np.ascontiguousarray(a.indptr, dtype=np.int64); ir = np.ascontiguousarray(a.indices, dtype=np.int64)
            pr = np.ascontiguousarray(a.data, dtype=np.float64)
            if pr.size == 0:
                ir = np.zeros(1, dtype=np.int64); pr = np.zeros(1)
            return L.stub_sparse(a.shape[0], a.shape[1], jc.ctypes.data, ir.ctypes.data, pr.ctypes.data)
        v = np.ascontiguousarray(np.atleast_1d(np.asarray(a, dtype=np.float64)))
        return L.stub_dense(v.size, 1, v.ctypes.data)

    def call(self, nlhs, cmd, *args):
        L = self.L
        ins = [self._in(cmd)] + [self._in(a) for a in args]
        prhs = (ct.c_void_p * len(ins))(*ins)
        plhs = (ct.c_void_p * max(nlhs, 1))()
        rc = L.stub_call(nlhs, plhs, len(ins), prhs)
        for p in ins:
            L.stub_free(p)
        if rc:
            raise RuntimeError("%s|%s" % (L.stub_last_id().decode(), L.stub_last_msg().decode()))
        outs = []
        for i in range(nlhs):
            m, n = L.stub_rows(plhs[i]), L.stub_cols(plhs[i])
            outs.append(np.ctypeslib.as_array(L.stub_data(plhs[i]), shape=(m * n,)).copy().reshape((n, m)).T if m * n else np.zeros((m, n)))
            L.stub_free(plhs[i])
        return outs


def test_gateway_compiles_against_the_stub_header():
    src = os.path.join(ROOT, "matlab", "cpk_b200_mex.cpp")
    subprocess.check_call(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "matlab", "stub"),
                           "-I" + os.path.join(ROOT, "include"), src])


def test_gateway_error_paths_without_a_gpu():
    mex = Mex()
    with pytest.raises(RuntimeError, match="unknown command"):
        mex.call(0, "no_such_command")
    with pytest.raises(RuntimeError, match="Invalid number of arguments."):          # opLDL2.m:61-63
        mex.call(1, "ldl2_create", sp.identity(3, format="csc"))
    with pytest.raises(RuntimeError, match="unknown handle"):
        mex.call(0, "destroy", 12345.0)
    with pytest.raises(RuntimeError, match="expected a real sparse double matrix"):
        mex.call(1, "system_create", np.ones(3), np.ones(3), 1.0)


@pytest.mark.gpu
def test_gateway_reg_solve_cpsymmlq_three_histories():
    import cpkrylov_b200 as cp
    s = load_system("cvxqp2_s")
    L, d, e, perm = load_factors("cvxqp2_s", "superlu")
    N = s["N"]
    D = sp.diags([d], [0], shape=(N, N), format="csc")
    P = sp.csc_matrix((np.ones(N), (perm, np.arange(N))), shape=(N, N))         # column k has its 1 in row perm[k]
    mex = Mex()
    hM, = mex.call(1, "ldl2_create", s["G"], s["A"], -s["C"], sp.csc_matrix(L), D, P)
    for name in ("nitref", "itref_tol", "residual_update", "force_itref"):      # reg_cpkrylov.m:135-148
        mex.call(0, "ldl2_set", float(hM[0, 0]), name, float(EX_OPTS[name]))
    z = np.random.default_rng(0).standard_normal(N)
    y, = mex.call(1, "ldl2_apply", float(hM[0, 0]), z)
    Mg = cp.opLDL2(s["G"], s["A"], -s["C"], factors=(L, d, e, perm))
    Mg.nitref, Mg.itref_tol, Mg.residual_update, Mg.force_itref = (EX_OPTS[k] for k in ("nitref", "itref_tol", "residual_update", "force_itref"))
    assert np.allclose(y.ravel(), Mg @ z, rtol=1e-13, atol=1e-13)
    Mg.close()
    hS, = mex.call(1, "system_create", s["Q"], s["C"], float(hM[0, 0]))
    ov = [EX_OPTS["atol"], EX_OPTS["rtol"], np.nan, float(EX_OPTS["itmax"]), np.nan, np.nan]
    x, niters, solved, status, hist, tdev = mex.call(6, "reg_solve", float(hS[0, 0]), 3.0, s["rhs"], ov, [float(s["n"]), float(s["m"])])
    xr, sr, fr = cp.reg_cpkrylov("cpsymmlq", s["rhs"], s["Q"], s["A"], s["C"], s["G"], dict(EX_OPTS), factors=(L, d, e, perm))
    assert int(niters[0, 0]) == sr["niters"] and bool(solved[0, 0]) == fr["solved"]
    assert np.allclose(x.ravel(), xr, rtol=1e-12, atol=0)
    assert hist.shape == (len(sr["cgresidHistory"]), 3)
    assert np.allclose(hist[:, 0], sr["cgresidHistory"], rtol=1e-12)
    assert np.allclose(hist[:, 1], sr["lqresidHistory"], rtol=1e-12)
    assert np.allclose(hist[:, 2], sr["qrresidHistory"], rtol=1e-12)
    assert tdev[0, 0] > 0
    mex.call(0, "destroy", float(hS[0, 0]))
    # the same solve with A as an operator (reg_cpkrylov.m:40): the gateway evaluates A*v through
    # mexCallMATLAB (the toy interpreter multiplies by the sparse matrix) while the kernel is resident
    mex.L.stub_callbacks.restype = ct.c_int
    c0 = mex.L.stub_callbacks()
    hS2, = mex.call(1, "system_create_op", float(s["n"]), s["Q"], s["C"], float(hM[0, 0]))
    x2, niters2, solved2 = mex.call(3, "reg_solve", float(hS2[0, 0]), 3.0, s["rhs"], ov, [float(s["n"]), float(s["m"])])
    assert int(niters2[0, 0]) == sr["niters"] and bool(solved2[0, 0]) == fr["solved"]
    assert mex.L.stub_callbacks() - c0 >= sr["niters"]
    assert np.allclose(x2.ravel(), xr, rtol=1e-7, atol=0)          # cpsymmlq amplifies rounding (profiles/r2_parity_table.jsonl)
    mex.call(0, "destroy", float(hS2[0, 0]))
    mex.call(0, "destroy", float(hM[0, 0]))
