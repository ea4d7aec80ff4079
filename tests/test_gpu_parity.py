"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI (ctypes
mirror of the reference interface), against the oracle on identical inputs and
identical LDL' factors.

Bar (north star): same convergence flag, iteration counts within +-2, solutions
agreeing to 1e-10 relative.  The Krylov recurrences amplify rounding-level
differences (reduction order); where GPU and oracle differ by more than 1e-10 an
extended-precision run of the oracle arbitrates (tests/parity.py):
err(GPU vs extended) <= max(1e-10, 4 x err(oracle fp64 vs extended)), hard cap 1e-7.
Residual histories are compared over their FULL length.
"""
import os

import numpy as np
import pytest
import scipy.sparse as sp

import parity
from helpers import EX_OPTS, kp_of, load_factors, load_system, relerr, small_kkt
from oracle import cpk_oracle as orc

pytestmark = pytest.mark.gpu

CASES1 = [("cpminres", {}), ("cpcg", {}), ("cpcglanczos", {}), ("cpdqgmres", {"mem": 2}),
          ("cpsymmlq", {}), ("cpgmres", {"restart": 50})]
CASES2 = [("cpgmres", {"restart": 100}), ("cpdqgmres", {"mem": 100}), ("cpgmres", {"restart": 20})]


@pytest.fixture(scope="module")
def cp():
    import cpkrylov_b200 as cp
    from cpkrylov_b200 import _lib
    assert _lib.lib().cpk_device_count() > 0, "GPU tests need a CUDA device"
    return cp


def _compare(cp, s, fac, meth, o, team, check_hist=True, case="fixture"):
    os.environ["CPK_TEAM"] = team
    try:
        xo, so, fo = orc.reg_cpkrylov(meth, s["rhs"], s["Q"], s["A"], s["C"], s["G"], o, factor=lambda K: fac)
        xg, sg, fg = cp.reg_cpkrylov(meth, s["rhs"], s["Q"], s["A"], s["C"], s["G"], o, factors=fac)
    finally:
        os.environ.pop("CPK_TEAM", None)
    assert fg["solved"] == fo["solved"]
    assert abs(sg["niters"] - so["niters"]) <= 2
    assert sg["gpu"]["launches"] == 1                       # the whole loop is one launch
    if fo["solved"]:
        parity.check_solution("%s N=%d %s" % (case, s["N"], team), meth, s, o, fac, xg, sg, xo, so)
    if check_hist and fo["solved"]:
        parity.check_history(sg, so)
    return xg, sg, xo, so


@pytest.mark.parametrize("team", ["cta", "grid"])
@pytest.mark.parametrize("meth,extra", CASES1)
def test_exprog1_all_solvers(cp, meth, extra, team):
    # examples/cpk_exprog1.m with every solver choice it lists (+ cpsymmlq, cpgmres)
    s = load_system("cvxqp1_m")
    fac = load_factors("cvxqp1_m", "superlu")
    xg, sg, xo, so = _compare(cp, s, fac, meth, dict(EX_OPTS, **extra), team)
    assert sg["gpu"]["shifted"]                             # rhs tail is nonzero: reg_cpkrylov.m:154-157
    xd = np.load(os.path.join(os.path.dirname(__file__), "golden", "cvxqp1_m_oracle.npz"))["x_direct"]
    assert relerr(xg, xd) < 5e-6                            # the example's own check vs K\rhs


@pytest.mark.parametrize("team", ["cta", "grid"])
@pytest.mark.parametrize("kind", ["superlu", "densebk"])
@pytest.mark.parametrize("meth,extra", CASES2)
def test_exprog2_nonsymmetric(cp, meth, extra, kind, team):
    # examples/cpk_exprog2.m: nonsymmetric H (SpMV must use H, not H'), 2x2 pivots with densebk
    s = load_system("cvxqp2_s")
    fac = load_factors("cvxqp2_s", kind)
    _compare(cp, s, fac, meth, dict(EX_OPTS, **extra), team)


@pytest.mark.parametrize("team", ["cta", "grid"])
@pytest.mark.parametrize("env", [{"CPK_LDL_SYNCFREE": "1"}, {"CPK_LDL_NO_SHORTCUTS": "1"}, {"CPK_LDL_NO_TAIL": "1"},
                                 {"CPK_LDL_SYNCFREE": "1", "CPK_LDL_NO_TAIL": "1"}, {}, {"CPK_LDL_COMPACT": "1"},
                                 {"CPK_LDL_RC": "1"}, {"CPK_LDL_RC": "1", "CPK_LDL_NO_TAIL": "1"}, {"CPK_LDL_RC": "1", "CPK_LDL_NO_SHORTCUTS": "1"},
                                 {"CPK_LDL_COMPACT": "1", "CPK_CW_NO_CHAINS": "1"}, {"CPK_LDL_TAIL_FILL": "2", "CPK_LDL_TAIL_MAXLEN": "32"},
                                 {"CPK_LDL_DENSE_META": "1", "CPK_LDL_WIDE_ROW": "8"}])
def test_ldl_walk_variants_agree(cp, env, team):
    """The LDL' solve has four walks (level-synchronous and sync-free/tagged walks of the item
    list, row-class passes for shallow sweeps with a diagonal D -- CPK_LDL_RC=1 --, and the
    shared-memory compact walk of the one-CTA team) and
    setup shortcuts (trivial/fused
    rows, level merging); every combination must give the oracle's answer (the
    switches are read when the operator is created).  The global-memory walks of the
    one-CTA team are reached with CPK_LDL_COMPACT=0; CPK_LDL_COMPACT=1 forces the
    compact walk on the shallow cvxqp2 factor as well."""
    if team == "cta" and "CPK_LDL_COMPACT" not in env and env:
        env = dict(env, CPK_LDL_COMPACT="0")
    s = load_system("cvxqp1_m")
    fac = load_factors("cvxqp1_m", "superlu")
    os.environ.update(env)
    try:
        _compare(cp, s, fac, "cpminres", dict(EX_OPTS), team)
        s2 = load_system("cvxqp2_s")
        _compare(cp, s2, load_factors("cvxqp2_s", "densebk"), "cpgmres", dict(EX_OPTS, restart=100), team)
    finally:
        for k in env:
            os.environ.pop(k, None)


@pytest.mark.parametrize("team", ["grid", "cta"])
def test_exprog1_dense_bk_2x2_pivots(cp, team):
    """Dense Bunch-Kaufman factor of cvxqp1: many 2x2 pivots, rows of thousands of entries
    (one-CTA team: the compact walk splits them into ordered parts of 512)."""
    from cpkrylov_b200.ldl import ldl_dense_bk
    s = load_system("cvxqp1_m")
    fac = ldl_dense_bk(kp_of(s))
    assert np.count_nonzero(fac[2]) > 100                   # many 2x2 pivots
    _compare(cp, s, fac, "cpminres", dict(EX_OPTS), team)


@pytest.mark.parametrize("team", ["cta", "grid"])
@pytest.mark.parametrize("kind", ["superlu", "densebk"])
def test_opldl2_apply_modes(cp, kind, team):
    """M*z against the oracle for every option combination of opLDL2.m:161-188."""
    from cpkrylov_b200.ldl import ldl_dense_bk, ldl_superlu
    s = small_kkt(300, 90, seed=11)
    KP = kp_of(s)
    fac = (ldl_superlu if kind == "superlu" else ldl_dense_bk)(KP)
    rng = np.random.default_rng(5)
    zs = rng.standard_normal((3, s["N"]))
    os.environ["CPK_TEAM"] = team
    try:
        for nitref, force, ru, stateful in [(0, False, False, False), (3, False, False, False), (2, True, False, False),
                                            (1, True, True, False), (1, True, True, True), (0, False, True, True),
                                            (3, False, True, True)]:
            Mo = orc.OpLDL2(s["G"], s["A"], -s["C"], *fac, ru_stateful=stateful)
            Mg = cp.opLDL2(s["G"], s["A"], -s["C"], factors=fac)
            for M in (Mo, Mg):
                M.nitref, M.force_itref, M.residual_update = nitref, force, ru
            Mg.ru_stateful = stateful
            Mg.set_track_rnorm(True)
            for z in zs:                                    # consecutive applies: the stateful reading keeps Aty/Cy
                n0 = Mo.nsolve
                yo, yg = Mo @ z, Mg @ z
                assert relerr(yg, yo) < 1e-11, (nitref, force, ru, stateful)
                assert Mg.last_stats["nldlsolve"] == Mo.nsolve - n0          # same refinement decisions
                if force:
                    assert Mg.last_stats["nldlsolve"] == 1 + nitref
                if nitref > 0:
                    assert Mg.rNorm <= 1e-10 * np.linalg.norm(z) + 10 * Mo.rNorm      # op.rNorm (opLDL2.m:186)
            if not ru:
                assert relerr(KP @ yg, z) < 1e-10          # K_P (M z) = z
            assert relerr(Mg.divide(z), KP @ z) < 1e-13    # opLDL2.m:193-195
            assert Mg.T is Mg and Mg.shape == (s["N"], s["N"])     # opLDL2.m:120-136
            Mg.close()
    finally:
        os.environ.pop("CPK_TEAM", None)


def test_opldl2_double_small(cp):
    # double(M) densifies by N applies (opLDL2.m:138-149)
    from cpkrylov_b200.ldl import ldl_superlu
    s = small_kkt(24, 8, seed=12)
    KP = kp_of(s)
    M = cp.opLDL2(s["G"], s["A"], -s["C"], factors=ldl_superlu(KP))
    assert np.allclose(M.double() @ KP.toarray(), np.eye(s["N"]), atol=1e-9)


@pytest.mark.parametrize("team", ["cta", "grid"])
def test_matvec_uses_H_not_its_transpose(cp, team):
    from cpkrylov_b200.ldl import ldl_superlu
    s = small_kkt(500, 120, seed=13, nonsym=True)
    os.environ["CPK_TEAM"] = team
    try:
        M = cp.opLDL2(s["G"], s["A"], -s["C"], factors=ldl_superlu(kp_of(s)))
        S = cp.KktSystem(s["Q"], s["C"], M)
        x = np.random.default_rng(1).standard_normal(s["n"])
        q = np.random.default_rng(2).standard_normal(s["m"])
        assert relerr(S.matvec(0, x), s["Q"] @ x) < 1e-14
        assert relerr(S.matvec(0, x), s["Q"].T @ x) > 1e-3
        assert relerr(S.matvec(1, q), s["C"] @ q) < 1e-14
        S.close()
    finally:
        os.environ.pop("CPK_TEAM", None)


@pytest.mark.parametrize("team", ["cta", "grid"])
def test_edge_cases(cp, team):
    from cpkrylov_b200.ldl import ldl_superlu
    s = small_kkt(200, 60, seed=14)
    fac = ldl_superlu(kp_of(s))
    os.environ["CPK_TEAM"] = team
    try:
        for meth in cp.SOLVERS:
            # already converged start (Appendix D.1): zero rhs
            x, st, fl = cp.reg_cpkrylov(meth, np.zeros(s["N"]), s["Q"], s["A"], s["C"], s["G"], {}, factors=fac)
            assert st["niters"] == 0 and fl["solved"] and np.all(x == 0)
            assert len(st.get("residHistory", st.get("cgresidHistory"))) == 1
            # itmax reached (D.2); cpgmres overshoots to a full cycle (cpgmres.m:148)
            o = dict(itmax=3, atol=0, rtol=1e-14, restart=2, mem=2)
            xg, sg, fg = cp.reg_cpkrylov(meth, s["rhs"], s["Q"], s["A"], s["C"], s["G"], o, factors=fac)
            xo, so, fo = orc.reg_cpkrylov(meth, s["rhs"], s["Q"], s["A"], s["C"], s["G"], o, factor=lambda K: fac)
            assert not fg["solved"] and sg["niters"] == so["niters"] == (4 if meth == "cpgmres" else 3)
            assert relerr(xg, xo) < 1e-10
        # zero b2: no shift (D.3)
        b = s["rhs"].copy(); b[s["n"]:] = 0
        xg, sg, fg = cp.reg_cpkrylov("cpminres", b, s["Q"], s["A"], s["C"], s["G"], dict(atol=1e-6, rtol=1e-6), factors=fac)
        assert not sg["gpu"]["shifted"] and sg["gpu"]["napply"] == sg["niters"] + 1
        assert relerr(s["K"] @ xg, b) < 1e-4
        # cpcglanczos backward-error stop + status string (D.10)
        o = dict(atol=0, rtol=1e-30, btol=1e-8, itmax=200)
        xg, sg, fg = cp.reg_cpkrylov("cpcglanczos", s["rhs"], s["Q"], s["A"], s["C"], s["G"], o, factors=fac)
        xo, so, fo = orc.reg_cpkrylov("cpcglanczos", s["rhs"], s["Q"], s["A"], s["C"], s["G"], o, factor=lambda K: fac)
        assert fg["solved"] and sg["status"] == so["status"] == "backward error small"
        assert abs(sg["niters"] - so["niters"]) <= 2
    finally:
        os.environ.pop("CPK_TEAM", None)


def test_solver_signature_method_b1_A_C_M(cp):
    # [x, y, stats, flag] = method(b1, A, C, M, opts) with M a GPU opLDL2 (cpminres.m:1)
    from cpkrylov_b200.ldl import ldl_superlu
    s = small_kkt(150, 40, seed=15)
    fac = ldl_superlu(kp_of(s))
    b1 = s["rhs"][:s["n"]]
    Mg = cp.opLDL2(s["G"], s["A"], -s["C"], factors=fac)
    Mo = orc.OpLDL2(s["G"], s["A"], -s["C"], *fac)
    for meth in cp.SOLVERS:
        xg, yg, sg, fg = cp.SOLVERS[meth](b1, s["Q"], s["C"], Mg, dict(print=False))
        xo, yo, so, fo = orc.SOLVERS[meth](b1, s["Q"], s["C"], Mo, dict(print=False))
        assert fg["solved"] == fo["solved"] and abs(sg["niters"] - so["niters"]) <= 2
        assert relerr(np.r_[xg, yg], np.r_[xo, yo]) < 1e-8


def test_indefinite_preconditioner_errors(cp):
    from cpkrylov_b200.ldl import ldl_dense_bk
    s = small_kkt(60, 16, seed=8)
    G = (-s["G"]).tocsc()
    f = ldl_dense_bk(sp.bmat([[G, s["A"].T], [s["A"], -s["C"]]], format="csc"))
    for meth, ident, text in (("cpminres", "", "does not behave as a spd matrix"),
                              ("cpsymmlq", "", "does not behave as a spd matrix"),
                              ("cpcglanczos", "CPCGLanczos:IndefiniteError", "not second-order sufficient")):
        with pytest.raises(cp.SolverError) as eg:
            cp.reg_cpkrylov(meth, s["rhs"], s["Q"], s["A"], s["C"], G, {}, factors=f)
        with pytest.raises(orc.SolverError) as eo:
            orc.reg_cpkrylov(meth, s["rhs"], s["Q"], s["A"], s["C"], G, {}, factor=lambda K: f)
        assert text in str(eg.value) and eg.value.identifier == ident
        assert str(eg.value).split(",")[0] == str(eo.value).split(",")[0]      # same "Iter k"


@pytest.mark.parametrize("meth,extra", [("cpcg", {}), ("cpminres", {}), ("cpdqgmres", {"mem": 20}), ("cpgmres", {"restart": 30})])
def test_synthetic_lap3d_reduced_vs_oracle(cp, meth, extra):
    # BASELINE cfg 3 / 4 generators at g=24 (grid team), reference-default options
    from cpkrylov_b200 import synth
    from cpkrylov_b200.ldl import ldl_superlu
    gen = synth.kkt_convdiff if meth in ("cpdqgmres", "cpgmres") else synth.kkt_lap3d
    w = gen(g=24)
    s = dict(Q=w["H"], A=w["B"], C=w["C"], G=w["G"], rhs=w["rhs"], n=w["n"], m=w["m"], N=w["n"] + w["m"])
    fac = ldl_superlu(synth.kp_matrix(w))
    xg, sg, xo, so = _compare(cp, s, fac, meth, dict(print=False, **extra), "grid")
    assert relerr(xg, w["xstar"]) < 1e-3


def test_stress_deep_filled_factor(cp):
    """SURVEY 8d 'stress' variant (k=6 entries per row of B in a 64-wide window) at reduced
    size: L has fill and hundreds of dependency levels, rows of hundreds of entries; the grid
    team then walks the sweep sync-free (chosen at setup from the depth)."""
    from cpkrylov_b200 import synth
    from cpkrylov_b200.ldl import ldl_superlu
    w = synth.kkt_lap3d(g=24, k=6, window=64, seed_B=2)
    s = dict(Q=w["H"], A=w["B"], C=w["C"], G=w["G"], rhs=w["rhs"], n=w["n"], m=w["m"], N=w["n"] + w["m"])
    fac = ldl_superlu(synth.kp_matrix(w))
    for team in ("grid", "cta"):
        xg, sg, xo, so = _compare(cp, s, fac, "cpcg", dict(print=False), team)
        assert relerr(xg, w["xstar"]) < 1e-3
    M = cp.opLDL2(w["G"], w["B"], -w["C"], factors=fac)
    assert M.info()["levels_fwd"] > 50
    M.close()


def test_full_size_properties_cfg3(cp):
    """BASELINE cfg 3 at full size (n = 10^6): size-independent properties instead
    of an oracle run -- operator identity, linearity of the apply, true residual
    of the solve."""
    from cpkrylov_b200 import synth
    w = synth.kkt_lap3d(g=100)
    KP = synth.kp_matrix(w)
    K = synth.kkt_matrix(w)
    N = w["n"] + w["m"]
    M = cp.opLDL2(w["G"], w["B"], -w["C"])
    rng = np.random.default_rng(0)
    z1, z2 = rng.standard_normal(N), rng.standard_normal(N)
    M.nitref = 0
    y1, y2, y3 = M @ z1, M @ z2, M @ (2.5 * z1 + z2)
    assert relerr(KP @ y1, z1) < 1e-10                     # K_P (M z) = z
    assert relerr(y3, 2.5 * y1 + y2) < 1e-12               # linearity
    M.close()
    x, st, fl = cp.reg_cpkrylov(cp.cpcg, w["rhs"], w["H"], w["B"], w["C"], w["G"], dict(print=False))
    assert fl["solved"] and 5 <= st["niters"] <= 60
    assert relerr(x, w["xstar"]) < 1e-3
    # P-norm stop test implies a small true residual in the first block
    r = K @ x - w["rhs"]
    assert np.linalg.norm(r) <= 1e-4 * np.linalg.norm(w["rhs"])


def test_batch_of_ipm_systems_one_launch(cp):
    # BASELINE cfg 5 (reduced): independent perturbed systems, one CTA each, one launch
    import ctypes as ct
    from cpkrylov_b200 import _lib, synth
    from cpkrylov_b200.ldl import ldl_superlu
    from cpkrylov_b200.batch import BatchSolver
    base = synth.load_cvxqp1()
    systems = [synth.ipm_batch_system(base, j) for j in range(6)]
    facs = [ldl_superlu(synth.kp_matrix(w)) for w in systems]
    o = dict(EX_OPTS)
    bs = BatchSolver(systems, facs, o)
    xs, stats = bs.solve("cpminres", [w["rhs"] for w in systems], o)
    assert bs.last_launches == 1
    for w, f, xg, st in zip(systems, facs, xs, stats):
        xo, so, fo = orc.reg_cpkrylov("cpminres", w["rhs"], w["H"], w["B"], w["C"], w["G"], o, factor=lambda K: f)
        assert st["solved"] == fo["solved"] and abs(st["niters"] - so["niters"]) <= 2
        if fo["solved"]:
            assert relerr(xg, xo) < 1e-8
    bs.close()


@pytest.mark.gpu
@pytest.mark.parametrize("meth,extra", [("cpcg", {}), ("cpdqgmres", {"mem": 8}), ("cpgmres", {"restart": 6})])
def test_batch_of_mid_size_systems_in_sub_teams(cp, meth, extra):
    """Systems too large for one CTA: the cooperative grid is cut into sub-teams, one system
    each, several systems per launch; a small system in the same call still goes to the one-CTA
    launch.  Every solution must match the CPU restatement and the whole-grid solve."""
    from cpkrylov_b200 import synth
    from cpkrylov_b200.ldl import ldl_superlu
    from cpkrylov_b200.batch import BatchSolver
    systems = [synth.kkt_lap3d(g=32, k=2, seed_B=10 + j, seed_x=20 + j) for j in range(5)]      # N = 40 960 each
    small = synth.ipm_batch_system(synth.load_cvxqp1(), 0)
    systems.insert(2, small)
    facs = [ldl_superlu(synth.kp_matrix(w)) for w in systems]
    o = dict(EX_OPTS, **extra)          # the example options keep the ill-conditioned IPM system inside attainable accuracy
    bs = BatchSolver(systems, facs, o)
    try:
        xs, stats = bs.solve(meth, [w["rhs"] for w in systems], o)
        assert 2 <= bs.last_launches <= 3                   # one CTA launch + one or two sub-team waves
        for j, (w, f, xg, st) in enumerate(zip(systems, facs, xs, stats)):
            xo, so, fo = orc.reg_cpkrylov(meth, w["rhs"], w["H"], w["B"], w["C"], w["G"], o, factor=lambda K, f=f: f)
            assert st["solved"] == fo["solved"] and abs(st["niters"] - so["niters"]) <= 2, j
            assert relerr(xg, xo) < 1e-7, (j, relerr(xg, xo))
            x1, s1, f1 = cp.reg_cpkrylov(meth, w["rhs"], w["H"], w["B"], w["C"], w["G"], o, factors=f)
            assert relerr(xg, x1) < 1e-8 and s1["niters"] == st["niters"], j
    finally:
        bs.close()


def test_one_operator_many_systems(cp):
    """method(b1, A, C, M, opts) with ONE opLDL2 and different (A, C): the reference lets M be
    reused with any A (cpminres.m:1).  The device copy of (A, C) cached on M is keyed by content:
    a second matrix, and an in-place change of the first one's values, must both reach the GPU."""
    from cpkrylov_b200.ldl import ldl_superlu
    s = small_kkt(150, 40, seed=21)
    fac = ldl_superlu(kp_of(s))
    b1 = s["rhs"][:s["n"]]
    Mg = cp.opLDL2(s["G"], s["A"], -s["C"], factors=fac)
    Mo = orc.OpLDL2(s["G"], s["A"], -s["C"], *fac)
    A1 = s["Q"].copy()
    A2 = (s["Q"] + sp.identity(s["n"]) * 0.5).tocsc()
    for A in (A1, A2, A1):
        xg, yg, sg, fg = cp.cpminres(b1, A, s["C"], Mg, dict(print=False))
        xo, yo, so, fo = orc.cpminres(b1, A, s["C"], Mo, dict(print=False))
        assert fg["solved"] == fo["solved"] and abs(sg["niters"] - so["niters"]) <= 2
        assert relerr(np.r_[xg, yg], np.r_[xo, yo]) < 1e-8
    A1.data *= 1.25                                         # same object, new values
    xg, yg, sg, fg = cp.cpminres(b1, A1, s["C"], Mg, dict(print=False))
    xo, yo, so, fo = orc.cpminres(b1, A1, s["C"], Mo, dict(print=False))
    assert relerr(np.r_[xg, yg], np.r_[xo, yo]) < 1e-8
    y = Mg @ np.ones(s["N"])                                # M is still the caller's and alive
    assert np.isfinite(y).all()
    Mg.close()


@pytest.mark.parametrize("team", ["cta", "grid"])
@pytest.mark.parametrize("gen", ["kkt_lap3d", "kkt_convdiff"])
def test_packed_stencil_matvec(cp, gen, team):
    """H of the synthetic configurations has 2 resp. 7 distinct values and a band below 2^15: its SELL
    entries are packed to 4 bytes (column offset + value index).  Products and a whole solve must be
    bit-identical to the unpacked form (CPK_SELL_PACK=0): the packing is lossless."""
    from cpkrylov_b200 import synth
    from cpkrylov_b200.ldl import ldl_superlu
    w = getattr(synth, gen)(g=14)
    fac = ldl_superlu(synth.kp_matrix(w))
    x = np.random.default_rng(3).standard_normal(w["n"])
    out = {}
    os.environ["CPK_TEAM"] = team
    try:
        for pack in ("1", "0"):
            os.environ["CPK_SELL_PACK"] = pack
            M = cp.opLDL2(w["G"], w["B"], -w["C"], factors=fac)
            S = cp.KktSystem(w["H"], w["C"], M)
            from cpkrylov_b200.solvers import reg_solve_on
            # kkt_convdiff has a nonsymmetric H: the symmetric solvers rightly stop with the indefinite error
            meth = "cpminres" if gen == "kkt_lap3d" else "cpdqgmres"
            xs, st, fl = reg_solve_on(S, meth, w["rhs"], dict(print=False))
            out[pack] = (S.matvec(0, x), xs, st["niters"])
            S.close()
    finally:
        os.environ.pop("CPK_TEAM", None)
        os.environ.pop("CPK_SELL_PACK", None)
    assert relerr(out["1"][0], w["H"] @ x) < 1e-14
    assert np.array_equal(out["1"][0], out["0"][0])
    assert np.array_equal(out["1"][1], out["0"][1]) and out["1"][2] == out["0"][2]
