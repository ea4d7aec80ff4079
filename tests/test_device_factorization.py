"""Device-side LDL' for sequences with a fixed pattern.

CPU part: the factorization PLAN (symbolic pattern with fill, triple-product lists, level
schedule) is built on the host; `cpk_debug_sqd_plan` returns it without a device, and a numpy
walk of the plan -- exactly what kernel `k_sqd_factor` does -- must reproduce the no-pivot LDL'
of the permuted matrix (oracle: SuperLU symmetric mode in natural order).
GPU part: `opLDL2(..., factors="device")`, `refactor`, `KktSystem.update` against the oracle."""
import ctypes as ct

import numpy as np
import pytest
import scipy.sparse as sp

from cpkrylov_b200 import _lib
from cpkrylov_b200.ldl import static_perm
from helpers import EX_OPTS, kp_of, load_system, relerr, small_kkt
from oracle import cpk_oracle as orc


def _plan(G, B, Cneg, perm):
    lib = _lib.lib()
    fn = lib.cpk_debug_sqd_plan
    P64, PI = ct.POINTER(ct.c_int64), ct.POINTER(ct.c_int)
    fn.restype = ct.c_int
    fn.argtypes = [ct.POINTER(_lib.CscStruct)] * 3 + [P64, P64, P64, P64] + [PI] * 6
    a, b, c = _lib.Csc(G), _lib.Csc(B), _lib.Csc(Cneg)
    perm = np.ascontiguousarray(perm, dtype=np.int64)
    sizes = np.zeros(5, dtype=np.int64)
    p64 = lambda x: x.ctypes.data_as(P64)
    pi = lambda x: x.ctypes.data_as(PI)
    _lib.check(fn(a.ref(), b.ref(), c.ref(), p64(perm), p64(sizes), None, None, None, None, None, None, None, None))
    ne, nnzL, nops, nlev, nin = (int(v) for v in sizes)
    N = perm.size
    out = dict(colptr=np.zeros(N + 1, np.int64), rowind=np.zeros(max(nnzL, 1), np.int64), asrc=np.zeros(ne, np.int32),
               dk=np.zeros(ne, np.int32), exec=np.zeros(ne, np.int32), levptr=np.zeros(nlev + 1, np.int32),
               opptr=np.zeros(ne + 1, np.int32), ops=np.zeros(max(3 * nops, 1), np.int32))
    _lib.check(fn(a.ref(), b.ref(), c.ref(), p64(perm), p64(sizes), p64(out["colptr"]), p64(out["rowind"]), pi(out["asrc"]),
                  pi(out["dk"]), pi(out["exec"]), pi(out["levptr"]), pi(out["opptr"]), pi(out["ops"])))
    out.update(ne=ne, nnzL=nnzL, nops=nops, nlev=nlev, N=N,
               vals_in=np.concatenate([a.val[:G.nnz], b.val[:B.nnz], c.val[:Cneg.nnz]]))
    return out


def _walk_plan(p):
    """numpy restatement of k_sqd_factor; also checks that a level only reads earlier levels."""
    f = np.where(p["asrc"] >= 0, p["vals_in"][np.maximum(p["asrc"], 0)], 0.0)
    done = np.zeros(p["ne"], dtype=bool)
    for l in range(p["nlev"]):
        nodes = p["exec"][p["levptr"][l]:p["levptr"][l + 1]]
        new = {}
        for e in nodes:
            v = f[e]
            for t in range(p["opptr"][e], p["opptr"][e + 1]):
                a, b, c = p["ops"][3 * t:3 * t + 3]
                assert done[a] and done[b] and done[c]
                v = v - (f[a] * f[c]) * f[b]
            if p["dk"][e] >= 0:
                assert done[p["dk"][e]]
                v = v / f[p["dk"][e]]
            new[e] = v
        for e, v in new.items():
            f[e] = v
            done[e] = True
    assert done.all()
    N, nnzL = p["N"], p["nnzL"]
    L = sp.csc_matrix((f[:nnzL], p["rowind"][:nnzL], p["colptr"]), shape=(N, N)) + sp.identity(N, format="csc")
    return L, f[nnzL:]


@pytest.fixture(scope="module")
def cpk_lib():
    try:
        return _lib.lib()
    except _lib.CpkLibraryMissing:
        pytest.skip("libcpk_b200.so not built")


@pytest.mark.parametrize("case", ["cvxqp2_s", "random0", "random1"])
def test_plan_walk_reproduces_the_no_pivot_factorization(cpk_lib, case):
    s = load_system(case) if case.startswith("cvxqp") else small_kkt(n=80, m=30, seed=int(case[-1]))
    KP = kp_of(s)
    perm = static_perm(KP)
    p = _plan(s["G"], s["A"], -s["C"], perm)
    L, d = _walk_plan(p)
    Lo, do = orc.ldl_static_perm(KP, perm)
    assert np.abs(d - do).max() <= 1e-10 * np.abs(do).max()
    assert abs(L - Lo).max() <= 1e-10 * max(1.0, abs(Lo).max())
    # and it is a factorization of the permuted K_P
    Kp = KP[perm][:, perm]
    R = (L @ sp.diags(d) @ L.T - Kp).tocoo()
    assert np.abs(R.data).max() <= 1e-10 * np.abs(Kp.data).max()
    assert p["nlev"] < p["ne"]


def test_plan_rejects_a_structurally_zero_diagonal(cpk_lib):
    s = small_kkt(n=20, m=6, seed=3)
    Cz = sp.csc_matrix((6, 6))
    with pytest.raises(_lib.CpkError):
        _plan(s["G"], s["A"], Cz, np.arange(26))


@pytest.mark.gpu
def test_device_factor_matches_oracle_and_solves():
    import cpkrylov_b200 as cp
    s = load_system("cvxqp1_m")
    KP = kp_of(s)
    perm = static_perm(KP)
    M = cp.opLDL2(s["G"], s["A"], -s["C"], factors="device", perm=perm)
    L, d = M.device_factor()
    Lo, do = orc.ldl_static_perm(KP, perm)
    assert np.abs(d - do).max() <= 1e-9 * np.abs(do).max()
    assert abs(L - Lo).max() <= 1e-9 * max(1.0, abs(Lo).max())
    z = np.random.default_rng(0).standard_normal(s["N"])
    M.nitref = 0
    y = M @ z
    assert relerr(KP @ y, z) <= 1e-6
    M.close()


@pytest.mark.gpu
@pytest.mark.parametrize("meth", ["cpminres", "cpcg"])
def test_sequence_refactorized_in_place(meth):
    """An IPM-like sequence on the cvxqp1 pattern: one operator, refactorized on the device for
    every system, against the CPU restatement with its own host factorization of each system."""
    import cpkrylov_b200 as cp
    from cpkrylov_b200 import synth
    from cpkrylov_b200.ldl import ldl_superlu
    from cpkrylov_b200.operators import KktSystem
    base = synth.load_cvxqp1()
    seq = [synth.ipm_batch_system(base, j) for j in range(4)]
    perm = static_perm(synth.kp_matrix(seq[0]))
    w = seq[0]
    M = cp.opLDL2(w["G"], w["B"], -w["C"], factors="device", perm=perm)
    S = KktSystem(w["H"], w["C"], M)
    o = dict(EX_OPTS)
    try:
        for j, w in enumerate(seq):
            if j:
                M.refactor(w["G"], w["B"], -w["C"])
                S.update(w["H"], w["C"])
            x, st, fl = cp.reg_solve_on(S, meth, w["rhs"], o)
            fac = ldl_superlu(synth.kp_matrix(w))
            xo, so, fo = orc.reg_cpkrylov(meth, w["rhs"], w["H"], w["B"], w["C"], w["G"], o, factor=lambda K, f=fac: f)
            assert fl["solved"] == fo["solved"] and abs(st["niters"] - so["niters"]) <= 2, j
            assert relerr(x, xo) <= 1e-7, (j, relerr(x, xo))
    finally:
        S.close()


@pytest.mark.gpu
def test_refactor_error_paths():
    import cpkrylov_b200 as cp
    from cpkrylov_b200.ldl import ldl_superlu
    from cpkrylov_b200.operators import KktSystem
    s = load_system("cvxqp2_s")
    KP = kp_of(s)
    # an operator built from host factors has no plan
    Mh = cp.opLDL2(s["G"], s["A"], -s["C"], factors=ldl_superlu(KP))
    with pytest.raises(_lib.CpkError):
        Mh.refactor(s["G"], s["A"], -s["C"])
    Mh.close()
    perm = static_perm(KP)
    M = cp.opLDL2(s["G"], s["A"], -s["C"], factors="device", perm=perm)
    # another pattern is refused
    A2 = s["A"].tolil(); A2[0, :] = 1.0
    with pytest.raises(_lib.CpkError):
        M.refactor(s["G"], A2.tocsc(), -s["C"])
    # a singular pivot is reported, not propagated as NaN
    with pytest.raises(_lib.CpkError) as ei:
        Gz = s["G"].copy(); Gz.data[:] = 0.0           # same pattern, zero values: the first pivot vanishes
        M.refactor(Gz, s["A"], -s["C"])
    assert "pivot" in str(ei.value) or "pattern" in str(ei.value)
    # after a refactorization the factor only lives in the compact-walk stream: a solver whose
    # scratch leaves no room for it must fail loudly (checked below on the larger fixture)
    M.refactor(s["G"], s["A"], -s["C"])
    S = KktSystem(s["Q"], s["C"], M)
    try:
        x, st, fl = cp.reg_solve_on(S, "cpgmres", s["rhs"], dict(EX_OPTS, restart=100))
        xo, so, fo = orc.reg_cpkrylov("cpgmres", s["rhs"], s["Q"], s["A"], s["C"], s["G"], dict(EX_OPTS, restart=100),
                                      factor=lambda K: ldl_superlu(KP))
        assert fl["solved"] == fo["solved"] and abs(st["niters"] - so["niters"]) <= 2 and relerr(x, xo) <= 1e-7
    finally:
        S.close()
    # ... on the larger fixture: 92 KB of stream region + 141 KB of cpdqgmres(128) scratch do not fit
    s1 = load_system("cvxqp1_m")
    M1 = cp.opLDL2(s1["G"], s1["A"], -s1["C"], factors="device", perm=static_perm(kp_of(s1)))
    M1.refactor(s1["G"], s1["A"], -s1["C"])
    S1 = KktSystem(s1["Q"], s1["C"], M1)
    try:
        with pytest.raises(_lib.CpkError) as ei:
            cp.reg_solve_on(S1, "cpdqgmres", s1["rhs"], dict(EX_OPTS, mem=128))
        assert "compact-walk stream" in str(ei.value)
    finally:
        S1.close()
