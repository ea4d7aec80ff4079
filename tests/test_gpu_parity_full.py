"""GPU parity at the sizes the numbers are quoted on (-m gpu), plus the scalar corner cases.

* BASELINE cfg 3 (kkt_lap3d g=100, N = 1.25 M): all six solvers against the oracle run to
  completion on identical (L, D, p): flag, |d niters| <= 2, FULL residual history, solution
  through the arbitrated bar of tests/parity.py.
* BASELINE cfg 4 (kkt_convdiff g=126, N = 2.5 M): cpdqgmres(20) and cpgmres(50).
* BASELINE cfg 5: all 256 systems of one ipm_batch against the oracle.
* util/SymGivens.m on the device, every branch; lucky breakdown (beta == 0, H(k+1,k) == 0:
  cpminres.m:202, cpgmres.m:223); the CPK_ERR_BREAKDOWN sites of cpcg / cpgmres / cpdqgmres.
"""
import ctypes as ct
import os

import numpy as np
import pytest
import scipy.sparse as sp

import parity
from helpers import EX_OPTS, small_kkt
from oracle import cpk_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cp():
    import cpkrylov_b200 as cp
    from cpkrylov_b200 import _lib
    assert _lib.lib().cpk_device_count() > 0, "GPU tests need a CUDA device"
    return cp


# ---------------------------------------------------------------------------
# full-size synthetic configurations
# ---------------------------------------------------------------------------
class _Big:
    """system + factor + GPU operator, built once per module"""

    def __init__(self, cp, w, opts):
        from cpkrylov_b200 import synth
        from cpkrylov_b200.ldl import ldl_superlu
        from cpkrylov_b200.operators import KktSystem, opLDL2
        from cpkrylov_b200.solvers import apply_opts_to_M
        self.w = w
        self.fac = ldl_superlu(synth.kp_matrix(w))
        self.M = opLDL2(w["G"], w["B"], -w["C"], factors=self.fac)
        apply_opts_to_M(self.M, opts)
        self.S = KktSystem(w["H"], w["C"], self.M)

    def close(self):
        self.S.close()


@pytest.fixture(scope="module")
def cfg3(cp):
    from cpkrylov_b200 import synth
    b = _Big(cp, synth.kkt_lap3d(g=100), {})
    yield b
    b.close()


@pytest.fixture(scope="module")
def cfg4(cp):
    from cpkrylov_b200 import synth
    b = _Big(cp, synth.kkt_convdiff(g=126), dict(nitref=1, force_itref=True))
    yield b
    b.close()


def _full_size_case(cp, big, case, meth, o):
    from cpkrylov_b200.solvers import reg_solve_on
    w = big.w
    xo, so, fo = parity.oracle_run(meth, w, o, big.fac)
    xg, sg, fg = reg_solve_on(big.S, meth, w["rhs"], o)
    assert fg["solved"] == fo["solved"], (fg, fo)
    assert abs(sg["niters"] - so["niters"]) <= 2, (sg["niters"], so["niters"])
    assert sg["gpu"]["launches"] == 1
    dev = parity.check_history(sg, so)
    row = parity.check_solution(case, meth, w, o, big.fac, xg, sg, xo, so)
    parity.record(dict(case=case, solver=meth, history_max_rel_dev=dev))
    assert parity.relerr(xg, w["xstar"]) < 1e-3
    return row


@pytest.mark.parametrize("meth,extra", [("cpcg", {}), ("cpcglanczos", {}), ("cpminres", {}), ("cpsymmlq", {}),
                                        ("cpgmres", {"restart": 20}), ("cpdqgmres", {"mem": 20})])
def test_cfg3_full_size_all_solvers(cp, cfg3, meth, extra):
    """BASELINE cfg 3 at n = 10^6, reference-default options (nitref = 3, itref_tol = 1e-8)"""
    _full_size_case(cp, cfg3, "cfg3 kkt_lap3d g=100", meth, dict(atol=1e-6, rtol=1e-6, **extra))


@pytest.mark.parametrize("meth,extra", [("cpdqgmres", {"mem": 20}), ("cpgmres", {"restart": 50})])
def test_cfg4_full_size(cp, cfg4, meth, extra):
    """BASELINE cfg 4 at n = 2.0 * 10^6, nonsymmetric H, forced refinement"""
    _full_size_case(cp, cfg4, "cfg4 kkt_convdiff g=126", meth,
                    dict(atol=1e-6, rtol=1e-6, itmax=500, nitref=1, force_itref=True, **extra))


def test_cfg5_all_256_systems(cp):
    """BASELINE cfg 5: one whole batch of 256 IPM-like systems (cvxqp1 pattern), cpminres with
    the example options, every system against the oracle"""
    from cpkrylov_b200 import synth
    from cpkrylov_b200.batch import BatchSolver
    from cpkrylov_b200.ldl import ldl_superlu
    base = synth.load_cvxqp1()
    systems = [synth.ipm_batch_system(base, j) for j in range(256)]
    facs = [ldl_superlu(synth.kp_matrix(w)) for w in systems]
    o = dict(EX_OPTS)
    bs = BatchSolver(systems, facs, o)
    try:
        xs, stats = bs.solve("cpminres", [w["rhs"] for w in systems], o)
        worst, dn = 0.0, 0
        for j, (w, f, xg, st) in enumerate(zip(systems, facs, xs, stats)):
            xo, so, fo = parity.oracle_run("cpminres", w, o, f)
            assert st["solved"] == fo["solved"], j
            assert abs(st["niters"] - so["niters"]) <= 2, (j, st["niters"], so["niters"])
            dn = max(dn, abs(st["niters"] - so["niters"]))
            if fo["solved"] and st["niters"] == so["niters"]:
                e = parity.relerr(xg, xo)
                worst = max(worst, e)
                assert e <= 1e-8, (j, e)       # cpminres on this pattern: 1e-14 .. 1e-10 measured
        parity.record(dict(case="cfg5 ipm_batch x256", solver="cpminres", worst_gpu_vs_oracle=worst, max_diff_niters=dn))
    finally:
        bs.close()


# ---------------------------------------------------------------------------
# scalar corner cases
# ---------------------------------------------------------------------------
def test_sym_givens_device_branches(cp):
    """util/SymGivens.m:4-28 on the device, all five branches, bit for bit against the oracle"""
    from cpkrylov_b200 import _lib
    L = _lib.lib()
    fn = L.cpk_debug_sym_givens
    fn.restype = ct.c_int
    fn.argtypes = [ct.c_int, ct.c_int] + [ct.c_void_p] * 5
    a = np.array([0.0, 3.0, -3.0, 0.0, 0.0, 1.0, -1.0, 2.0, -2.0, 5.0, -5.0, 1e-300, 1e300, 7.0, -0.0, 4.0])
    b = np.array([0.0, 0.0, 0.0, 2.0, -2.0, 2.0, 2.0, -1.0, -1.0, 5.0, 5.0, 1e300, 1e-300, -7.0, 0.0, -0.0])
    n = a.size
    c, s, d = np.empty(n), np.empty(n), np.empty(n)
    _lib.check(fn(0, n, a.ctypes.data, b.ctypes.data, c.ctypes.data, s.ctypes.data, d.ctypes.data))
    for i in range(n):
        co, so_, do = orc.sym_givens(float(a[i]), float(b[i]))
        assert (c[i], s[i], d[i]) == (co, so_, do), (a[i], b[i], (c[i], s[i], d[i]), (co, so_, do))
        # [c s; s -c] [a; b] = [d; 0]
        assert abs(c[i] * a[i] + s[i] * b[i] - d[i]) <= 1e-15 * max(abs(d[i]), 1e-300)


def _exact_system():
    """Exactly representable data with H = G: the preconditioner is the exact inverse, so the
    Krylov space closes after one step and the recurrences hit beta == 0 / H(k+1,k) == 0."""
    n, m = 4, 2
    H = sp.csc_matrix(2.0 * np.eye(n))
    B = sp.csc_matrix(np.array([[1.0, 0, 0, 0], [0, 1.0, 0, 0]]))
    C = sp.csc_matrix(np.eye(m))
    K = sp.bmat([[H, B.T], [B, -C]], format="csc")
    return dict(H=H, B=B, C=C, G=H.copy(), n=n, m=m, K=K)


@pytest.mark.parametrize("meth", ["cpcg", "cpcglanczos", "cpminres", "cpsymmlq", "cpgmres", "cpdqgmres"])
def test_lucky_breakdown_exact(cp, meth):
    """beta == 0 (cpminres.m:202, cpcglanczos.m:257, cpsymmlq.m:281) and H(k+1,k) == 0
    (cpgmres.m:223, cpdqgmres.m:222): no normalisation, the loop goes on; same iterates as
    the oracle (all sums have one nonzero term: no reduction-order noise)."""
    from cpkrylov_b200.ldl import ldl_dense_bk
    s = _exact_system()
    fac = ldl_dense_bk(sp.bmat([[s["G"], s["B"].T], [s["B"], -s["C"]]], format="csc"))
    xs = np.array([0.0, 0, 8.0, 0, 0, 0])
    b = s["K"] @ xs
    o = dict(print=False, atol=0.0, rtol=0.0, itmax=3, restart=3, mem=3, nitref=0)
    xo, so, fo = orc.reg_cpkrylov(meth, b, s["H"], s["B"], s["C"], s["G"], o, factor=lambda K: fac)
    xg, sg, fg = cp.reg_cpkrylov(meth, b, s["H"], s["B"], s["C"], s["G"], o, factors=fac)
    assert sg["niters"] == so["niters"] and fg["solved"] == fo["solved"]
    key = "cgresidHistory" if meth == "cpsymmlq" else "residHistory"
    ho, hg = np.asarray(so[key]), np.asarray(sg[key])
    assert len(ho) == len(hg)
    assert np.array_equal(np.isnan(ho), np.isnan(hg))
    assert np.allclose(hg[~np.isnan(ho)], ho[~np.isnan(ho)], rtol=1e-14, atol=0.0)
    assert (0.0 in ho.tolist()) or meth in ("cpgmres", "cpdqgmres")       # the zero was really hit
    if not np.isnan(xo).any():
        assert np.allclose(xg, xo, rtol=1e-14, atol=1e-300)


@pytest.mark.parametrize("nneg,seed", [(60, 8), (3, 10), (8, 9)])
@pytest.mark.parametrize("meth,extra", [("cpcg", {}), ("cpgmres", {"restart": 4}), ("cpdqgmres", {"mem": 4})])
def test_breakdown_sites(cp, meth, extra, nneg, seed):
    """A preconditioner that is indefinite on the constraint null space makes the P-inner
    product negative under a square root: complex in MATLAB (cpcg.m:175, cpgmres.m:173,219,
    cpdqgmres.m:157,218), NaN in the oracle, CPK_ERR_BREAKDOWN with the iteration index on the
    GPU.  (60, 8): at the start; (3, 10), (8, 9): inside the loop."""
    from cpkrylov_b200 import _lib
    from cpkrylov_b200.ldl import ldl_dense_bk
    s = small_kkt(60, 16, seed=seed)
    g = s["G"].diagonal().copy()
    idx = np.random.default_rng(seed).permutation(60)[:nneg]
    g[idx] = -g[idx]
    G = sp.diags(g).tocsc()
    f = ldl_dense_bk(sp.bmat([[G, s["A"].T], [s["A"], -s["C"]]], format="csc"))
    o = dict(print=False, itmax=30, **extra)
    xo, so, fo = orc.reg_cpkrylov(meth, s["rhs"], s["Q"], s["A"], s["C"], G, o, factor=lambda K: f)
    nan = np.flatnonzero(np.isnan(so["residHistory"]))
    assert nan.size > 0
    with pytest.raises(cp.SolverError) as eg:
        cp.reg_cpkrylov(meth, s["rhs"], s["Q"], s["A"], s["C"], G, o, factors=f)
    assert eg.value.code == _lib.CPK_ERR_BREAKDOWN
    assert str(eg.value).startswith("Iter %d," % int(nan[0])), (str(eg.value), int(nan[0]))
