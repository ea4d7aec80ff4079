"""Matrix-free A (reg_cpkrylov.m:40: "the argument A may be a matrix or a linear operator").

The GPU path keeps the whole solve in ONE persistent kernel; the product A*v is answered by a
host callback through a mailbox (include/cpk_b200.h: cpk_system_create_op).  Parity: the same
system solved with A as an operator and with A as a matrix gives the same flag, iteration
count and -- to rounding: the fused dot products visit the rows in another order -- the same
iterates; and both agree with the oracle run on the operator.
"""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from helpers import EX_OPTS, kp_of, load_system, relerr, small_kkt
from oracle import cpk_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cp():
    import cpkrylov_b200 as cp
    from cpkrylov_b200 import _lib
    assert _lib.lib().cpk_device_count() > 0, "GPU tests need a CUDA device"
    return cp


class SpotLike:
    """an operator in the sense of the Spot toolbox: size + mtimes, nothing else"""

    def __init__(self, A):
        self._A = sp.csr_matrix(A)
        self.shape = A.shape
        self.calls = 0

    def __matmul__(self, v):
        self.calls += 1
        return self._A @ v


def _with_team(team, fn):
    import os
    os.environ["CPK_TEAM"] = team
    try:
        return fn()
    finally:
        os.environ.pop("CPK_TEAM", None)


@pytest.mark.parametrize("team", ["cta", "grid"])
@pytest.mark.parametrize("meth,extra", [("cpcg", {}), ("cpcglanczos", {}), ("cpminres", {}), ("cpsymmlq", {}),
                                        ("cpgmres", {"restart": 30}), ("cpdqgmres", {"mem": 30})])
def test_operator_equals_matrix_exprog1(cp, meth, extra, team):
    """examples/cpk_exprog1.m system (N = 5500, rhs with a nonzero tail: the shift of
    reg_cpkrylov.m:154-158 also goes through the operator)"""
    from cpkrylov_b200.ldl import ldl_superlu
    s = load_system("cvxqp1_m")
    fac = ldl_superlu(kp_of(s))
    o = dict(EX_OPTS, **extra)
    op = SpotLike(s["Q"])
    xm, sm, fm = _with_team(team, lambda: cp.reg_cpkrylov(meth, s["rhs"], s["Q"], s["A"], s["C"], s["G"], o, factors=fac))
    xo, so, fo = _with_team(team, lambda: cp.reg_cpkrylov(meth, s["rhs"], op, s["A"], s["C"], s["G"], o, factors=fac))
    assert fo["solved"] == fm["solved"] and so["niters"] == sm["niters"]
    assert so["gpu"]["launches"] == 1                       # still one persistent kernel
    assert so["gpu"]["shifted"]
    # one product per iteration + the shift (+ cpsymmlq / restarts: a few more)
    assert so["niters"] + 1 <= op.calls <= so["niters"] + 8, (op.calls, so["niters"])
    tol = 1e-10 if meth in ("cpcg", "cpminres", "cpgmres", "cpdqgmres") else 2e-8     # see profiles/r2_parity_table.jsonl
    assert relerr(xo, xm) < tol, relerr(xo, xm)
    xr, sr, fr = orc.reg_cpkrylov(meth, s["rhs"], spla.aslinearoperator(s["Q"]), s["A"], s["C"], s["G"], o, factor=lambda K: fac)
    assert fr["solved"] == fo["solved"] and abs(sr["niters"] - so["niters"]) <= 2
    if sr["niters"] == so["niters"]:
        assert relerr(xo, xr) < max(tol, 1e-9)


@pytest.mark.parametrize("team", ["cta", "grid"])
def test_operator_nonsymmetric_exprog2(cp, team):
    """examples/cpk_exprog2.m system (nonsymmetric (1,1) block): A*v, never A'*v"""
    from cpkrylov_b200.ldl import ldl_superlu
    s = load_system("cvxqp2_s")
    fac = ldl_superlu(kp_of(s))
    o = dict(EX_OPTS, restart=100)
    op = SpotLike(s["Q"])
    xm, sm, fm = _with_team(team, lambda: cp.reg_cpkrylov("cpgmres", s["rhs"], s["Q"], s["A"], s["C"], s["G"], o, factors=fac))
    xo, so, fo = _with_team(team, lambda: cp.reg_cpkrylov("cpgmres", s["rhs"], op, s["A"], s["C"], s["G"], o, factors=fac))
    assert fo["solved"] == fm["solved"] and so["niters"] == sm["niters"]
    assert relerr(xo, xm) < 1e-9


def test_solver_signature_with_operator(cp):
    """method(b1, A, C, M, opts) with A an operator (cpminres.m:1), M the caller's"""
    from cpkrylov_b200.ldl import ldl_dense_bk
    s = small_kkt(80, 24, seed=5)
    fac = ldl_dense_bk(kp_of(s))
    M = cp.opLDL2(s["G"], s["A"], -s["C"], factors=fac)
    b1 = s["rhs"][:80]
    op = spla.aslinearoperator(s["Q"])
    x1, y1, st1, f1 = cp.cpminres(b1, s["Q"], s["C"], M, dict(print=False))
    x2, y2, st2, f2 = cp.cpminres(b1, op, s["C"], M, dict(print=False))
    assert st1["niters"] == st2["niters"] and f1["solved"] == f2["solved"]
    assert relerr(np.concatenate([x2, y2]), np.concatenate([x1, y1])) < 1e-10
    M.close()


def test_operator_failure_is_reported(cp):
    """an exception inside the callback stops the resident kernel and comes back as itself"""
    from cpkrylov_b200.ldl import ldl_dense_bk
    s = small_kkt(60, 16, seed=2)
    fac = ldl_dense_bk(kp_of(s))

    class Boom:
        shape = s["Q"].shape
        calls = 0

        def __matmul__(self, v):
            Boom.calls += 1
            if Boom.calls == 3:
                raise ZeroDivisionError("operator failed on purpose")
            return s["Q"] @ v

    with pytest.raises(ZeroDivisionError):
        cp.reg_cpkrylov("cpminres", s["rhs"], Boom(), s["A"], s["C"], s["G"], dict(print=False), factors=fac)
    # the library is still usable afterwards
    x, st, fl = cp.reg_cpkrylov("cpminres", s["rhs"], s["Q"], s["A"], s["C"], s["G"], dict(print=False), factors=fac)
    assert fl["solved"]
