"""CPU tests (-m "not gpu"): the oracle against the committed golden vectors, the
reference's own self-check, operator identities and solver invariants."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from helpers import EX_OPTS, kp_of, load_factors, load_oracle, load_system, relerr, small_kkt
from oracle import cpk_oracle as orc
from cpkrylov_b200.ldl import ldl_dense_bk, ldl_superlu

CASES1 = [("cpminres", {}), ("cpcg", {}), ("cpcglanczos", {}), ("cpdqgmres", {"mem": 2}),
          ("cpsymmlq", {}), ("cpgmres", {"restart": 50})]
CASES2 = [("cpgmres", {"restart": 100}), ("cpdqgmres", {"mem": 100}), ("cpgmres", {"restart": 20}),
          ("cpdqgmres", {"mem": 10})]


def _tag(kind, meth, extra):
    return "%s/%s%s" % (kind, meth, "".join("_%s%d" % kv for kv in sorted(extra.items())))


@pytest.mark.parametrize("a,b,exp", [(0.0, 0.0, (1.0, 0.0, 0.0)), (-3.0, 0.0, (-1.0, 0.0, 3.0)),
                                     (0.0, -2.0, (0.0, -1.0, 2.0)), (3.0, 4.0, (0.6, 0.8, 5.0)),
                                     (4.0, 3.0, (0.8, 0.6, 5.0)), (-4.0, 3.0, (-0.8, 0.6, 5.0))])
def test_sym_givens_branches(a, b, exp):
    # util/SymGivens.m:4-28, every branch; [c s; s -c][a;b] = [d;0]
    c, s, d = orc.sym_givens(a, b)
    assert np.allclose([c, s, d], exp, atol=1e-15)
    assert abs(c * a + s * b - d) < 1e-14 and abs(s * a - c * b) < 1e-14


def test_c_kernels_match_scipy():
    rng = np.random.default_rng(0)
    A = sp.random(200, 150, density=0.05, random_state=rng, format="csr")
    x = rng.standard_normal(150)
    assert np.allclose(orc._Csr(A).matvec(x), A @ x, rtol=1e-14, atol=1e-14)


@pytest.mark.parametrize("factor", [ldl_superlu, ldl_dense_bk])
def test_opldl2_identity_and_refinement(factor):
    # K_P * (M*z) = z ; divide is K_P*b (opLDL2.m:193-195)
    s = small_kkt(80, 30, seed=1)
    KP = kp_of(s)
    L, d, e, p = factor(KP)
    if factor is ldl_dense_bk:
        assert np.count_nonzero(e) > 0          # exercises the 2x2 D-solve
    M = orc.OpLDL2(s["G"], s["A"], -s["C"], L, d, e, p)
    z = np.random.default_rng(2).standard_normal(s["N"])
    for nitref, force in ((0, False), (3, False), (2, True)):
        M.nitref, M.force_itref = nitref, force
        n0 = M.nsolve
        y = M @ z
        assert relerr(KP @ y, z) < 1e-11
        assert M.nsolve - n0 == (1 + (nitref if force else 0))
    assert np.allclose(M.divide(z), KP @ z)


def test_opldl2_dimension_errors():
    s = small_kkt(20, 5)
    L, d, e, p = ldl_superlu(kp_of(s))
    with pytest.raises(ValueError, match="Incompatible dimensions"):
        orc.OpLDL2(s["G"], s["A"][:, :10], -s["C"], L, d, e, p)
    with pytest.raises(ValueError, match="must be square"):
        orc.OpLDL2(s["G"][:, :10], s["A"], -s["C"], L, d, e, p)


def test_residual_update_stateless_vs_stateful_agree_after_refinement():
    # with nitref>=1 & force_itref both readings give K_P^-1 x to rounding
    s = small_kkt(60, 20, seed=3)
    f = ldl_superlu(kp_of(s))
    zs = np.random.default_rng(4).standard_normal((3, s["N"]))
    outs = []
    for stateful in (False, True):
        M = orc.OpLDL2(s["G"], s["A"], -s["C"], *f, ru_stateful=stateful)
        M.residual_update, M.nitref, M.force_itref = True, 1, True
        outs.append([M @ z for z in zs])
    for a, b in zip(*outs):
        assert relerr(a, b) < 1e-10


@pytest.mark.parametrize("meth,extra", CASES1)
def test_oracle_matches_golden_cvxqp1(meth, extra):
    s = load_system("cvxqp1_m")
    fac = load_factors("cvxqp1_m", "superlu")
    g = load_oracle("cvxqp1_m")
    o = dict(EX_OPTS, **extra)
    x, st, fl = orc.reg_cpkrylov(meth, s["rhs"], s["Q"], s["A"], s["C"], s["G"], o, factor=lambda K: fac)
    tag = _tag("superlu", meth, extra)
    assert st["niters"] == int(g[tag + "/niters"]) and fl["solved"] == bool(g[tag + "/solved"])
    assert relerr(x, g[tag + "/x"]) < 1e-9
    # the reference's own check: norm(x - K\rhs)/norm(x) (cpk_exprog1.m:101-104)
    assert relerr(x, g["x_direct"]) < 5e-6


@pytest.mark.parametrize("kind", ["superlu", "densebk"])
@pytest.mark.parametrize("meth,extra", CASES2)
def test_oracle_matches_golden_cvxqp2(kind, meth, extra):
    s = load_system("cvxqp2_s")
    fac = load_factors("cvxqp2_s", kind)
    g = load_oracle("cvxqp2_s")
    o = dict(EX_OPTS, **extra)
    x, st, fl = orc.reg_cpkrylov(meth, s["rhs"], s["Q"], s["A"], s["C"], s["G"], o, factor=lambda K: fac)
    tag = _tag(kind, meth, extra)
    assert abs(st["niters"] - int(g[tag + "/niters"])) <= 2 and fl["solved"] == bool(g[tag + "/solved"])
    if fl["solved"]:
        assert relerr(x, g[tag + "/x"]) < 1e-7
        assert relerr(x, g["x_direct"]) < 5e-4      # cpk_exprog2.m:100-103


def test_factor_backends_agree_on_fixture():
    s = load_system("cvxqp2_s")
    KP = kp_of(s)
    z = np.random.default_rng(0).standard_normal(s["N"])
    ys = []
    for f in (ldl_superlu, ldl_dense_bk):
        M = orc.OpLDL2(s["G"], s["A"], -s["C"], *f(KP))
        M.nitref = 0
        ys.append(M @ z)
    assert relerr(ys[0], ys[1]) < 1e-8
    assert relerr(KP @ ys[0], z) < 1e-8


@pytest.mark.parametrize("meth", ["cpcg", "cpcglanczos", "cpminres", "cpsymmlq", "cpgmres", "cpdqgmres"])
def test_solver_invariants_small(meth):
    s = small_kkt(70, 25, seed=5, nonsym=meth in ("cpgmres", "cpdqgmres"))
    f = ldl_superlu(kp_of(s))
    o = dict(atol=1e-7, rtol=1e-7, itmax=400, mem=60, restart=60)
    x, st, fl = orc.reg_cpkrylov(meth, s["rhs"], s["Q"], s["A"], s["C"], s["G"], o, factor=lambda K: f)
    assert fl["solved"]
    assert relerr(x, s["xs"]) < 1e-5
    h = st.get("residHistory", st.get("cgresidHistory"))
    assert len(h) == st["niters"] + 1                    # history length niters+1
    if meth == "cpminres":
        assert np.all(np.diff(h) <= 1e-12 * h[0])        # MINRES residual estimate is monotone


def test_edge_already_converged_and_itmax():
    s = small_kkt(40, 10, seed=6)
    f = ldl_superlu(kp_of(s))
    for meth in orc.SOLVERS:
        x, st, fl = orc.reg_cpkrylov(meth, np.zeros(s["N"]), s["Q"], s["A"], s["C"], s["G"], {}, factor=lambda K: f)
        assert st["niters"] == 0 and fl["solved"] and np.all(x == 0)
        x, st, fl = orc.reg_cpkrylov(meth, s["rhs"], s["Q"], s["A"], s["C"], s["G"],
                                     dict(itmax=3, atol=0, rtol=1e-14, restart=2, mem=2), factor=lambda K: f)
        assert not fl["solved"]
        assert st["niters"] == (4 if meth == "cpgmres" else 3)      # cpgmres.m:148 overshoots to a full cycle


def test_zero_b2_takes_no_shift_path():
    s = small_kkt(40, 10, seed=7)
    f = ldl_superlu(kp_of(s))
    b = s["rhs"].copy(); b[s["n"]:] = 0.0
    x, st, fl = orc.reg_cpkrylov("cpminres", b, s["Q"], s["A"], s["C"], s["G"], dict(atol=1e-8, rtol=1e-8), factor=lambda K: f)
    assert relerr(s["K"] @ x, b) < 1e-6
    assert st["napply"] == st["niters"] + 1              # no extra M*[0;b2] apply


def test_indefinite_preconditioner_raises():
    s = small_kkt(30, 8, seed=8)
    G = (-s["G"]).tocsc()                                # wrong sign: P-inner product negative
    KP = sp.bmat([[G, s["A"].T], [s["A"], -s["C"]]], format="csc")
    f = ldl_dense_bk(KP)
    for meth, ident in (("cpminres", ""), ("cpcglanczos", "CPCGLanczos:IndefiniteError"), ("cpsymmlq", "")):
        with pytest.raises(orc.SolverError) as ei:
            orc.reg_cpkrylov(meth, s["rhs"], s["Q"], s["A"], s["C"], G, {}, factor=lambda K: f)
        assert "beta (before sqrt)" in str(ei.value) and ei.value.identifier == ident


def test_cglanczos_btol_status():
    s = small_kkt(50, 15, seed=9)
    f = ldl_superlu(kp_of(s))
    x, st, fl = orc.reg_cpkrylov("cpcglanczos", s["rhs"], s["Q"], s["A"], s["C"], s["G"],
                                 dict(atol=0, rtol=1e-30, btol=1e-8, itmax=200), factor=lambda K: f)
    assert fl["solved"] and st["status"] == "backward error small"
