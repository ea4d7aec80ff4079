"""Host-side check of the item list of the LDL' sweeps (cpk_device.cuh: DevSweep).

`build_sweeps` (cpk_host.cu) compiles the factors into one ordered item list: trivial and fused
rows compiled away, consecutive dependency levels MERGED into groups by substituting the in-group
dependencies (level merging), long rows as warp-rows.  `cpk_debug_sweep` returns the list without
touching a device.  This file walks it exactly like `ldl_solve_levels` does -- level after level,
nothing produced in a level may be read inside that level -- and compares with a direct solve of
P L D L' P' y = z (opLDL2.m:86).
"""
import ctypes as ct

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from cpkrylov_b200 import _lib, synth
from cpkrylov_b200.ldl import ldl_dense_bk, ldl_superlu
from helpers import kp_of, load_system, small_kkt

F_FWD, F_FUSED, F_STORE, F_WDIRECT, F_PARTNER, F_WARPROW = 1, 2, 4, 8, 16, 32


def _sweep(L, d, e, perm):
    lib = _lib.lib()
    fn = lib.cpk_debug_sweep
    fn.restype = ct.c_int
    PC = ct.POINTER(_lib.CscStruct)
    fn.argtypes = [PC, PC, ct.POINTER(ct.c_int64), ct.POINTER(ct.c_int64)] + [ct.c_void_p] * 13
    N = d.size
    Dm = sp.diags([d, e[:-1], e[:-1]], [0, -1, 1], shape=(N, N), format="csc")
    Lc, Dc = _lib.Csc(L), _lib.Csc(Dm)
    perm = np.ascontiguousarray(perm, dtype=np.int64)
    sizes = np.zeros(9, dtype=np.int64)
    pp = perm.ctypes.data_as(ct.POINTER(ct.c_int64))
    ps = sizes.ctypes.data_as(ct.POINTER(ct.c_int64))
    _lib.check(fn(Lc.ref(), Dc.ref(), pp, ps, *([None] * 13)))
    nitems, nfwd, nlev, nent = (int(v) for v in sizes[:4])
    i32 = lambda n: np.zeros(max(n, 1), dtype=np.int32)
    f64 = lambda n: np.zeros(max(n, 1))
    S = dict(nitems=nitems, nfwd=nfwd, nlev=nlev, lev_f=int(sizes[4]), lev_b=int(sizes[5]), merged_f=int(sizes[6]), merged_b=int(sizes[7]),
             levptr=i32(nlev + 1), sptr=i32(nitems + 1), col=i32(nent), val=f64(nent), rid=i32(nitems * 32), pidx=i32(nitems * 32),
             flags=i32(nitems * 32), d=f64(nitems * 32), partner=i32(nitems * 32), e=f64(nitems * 32), dp=f64(nitems * 32),
             nlone=int(sizes[8]), lone_pidx=i32(int(sizes[8])), lone_d=f64(int(sizes[8])))
    _lib.check(fn(Lc.ref(), Dc.ref(), pp, ps, *[S[k].ctypes.data for k in
                                                 ("levptr", "sptr", "col", "val", "rid", "pidx", "flags", "d", "partner", "e", "dp", "lone_pidx", "lone_d")]))
    return S


def _walk(S, N, z):
    """Level-synchronous walk of the item list (item_process<false> of cpk_kernels.cuh)."""
    wv = np.full(N, np.nan)
    yv = np.full(N, np.nan)
    out = np.full(N, np.nan)
    nwritten = np.zeros(N, dtype=np.int64)
    lanes = np.arange(32)
    # lone rows (ldl_lone_rows): y = z / d straight from the input vector
    lp = S["lone_pidx"][:S["nlone"]]
    out[lp] = z[lp] / S["lone_d"][:S["nlone"]]
    nwritten[lp] += 1
    for g in range(S["nlev"]):
        a, b = int(S["levptr"][g]), int(S["levptr"][g + 1])
        w_new, y_new = {}, {}
        for t in range(a, b):
            beg, end = int(S["sptr"][t]), int(S["sptr"][t + 1])
            sl = slice(t * 32, t * 32 + 32)
            rid, pidx, flg, dd = S["rid"][sl], S["pidx"][sl], S["flags"][sl], S["d"][sl]
            f0 = int(flg[0])
            isfwd, warprow = bool(f0 & F_FWD), bool(f0 & F_WARPROW)
            c = S["col"][beg:end].reshape(-1, 32)
            v = S["val"][beg:end].reshape(-1, 32)
            x = np.zeros(c.shape)
            m_in = c <= -2
            x[m_in] = z[-c[m_in] - 2]
            m_w = c >= N
            x[m_w] = wv[c[m_w] - N]
            m_s = (c >= 0) & (c < N)
            x[m_s] = (wv if isfwd else yv)[c[m_s]]
            assert not np.isnan(x[c != -1]).any(), "level %d item %d reads a value that is not there yet" % (g, t)
            x[c == -1] = 0.0
            sums = -(v * x).sum(axis=0)
            if warprow:
                sums = np.array([sums.sum()] + [0.0] * 31)
                live = lanes == 0
            else:
                live = rid >= 0
            for ln in np.flatnonzero(live):
                r, pi, fl = int(rid[ln]), int(pidx[ln]), int(flg[ln])
                if isfwd:
                    acc = z[pi]
                else:
                    w = z[pi] if fl & F_WDIRECT else wv[r]
                    assert not np.isnan(w)
                    if not fl & F_PARTNER:
                        acc = w / dd[ln]
                    else:
                        pr = int(S["partner"][t * 32 + ln])
                        ee, dp = S["e"][t * 32 + ln], S["dp"][t * 32 + ln]
                        acc = (dp * w - ee * wv[pr]) / (dd[ln] * dp - ee * ee)
                acc += sums[ln]
                if isfwd and not fl & F_FUSED:
                    w_new[r] = acc
                else:
                    if isfwd:
                        acc = acc / dd[ln]
                    if isfwd or fl & F_STORE:
                        y_new[r] = acc
                    out[pi] = acc
                    nwritten[pi] += 1
        for r, val in w_new.items():
            wv[r] = val
        for r, val in y_new.items():
            yv[r] = val
    assert (nwritten == 1).all(), "every element of the result is written exactly once"
    return out


def _direct(L, d, e, perm, z):
    N = d.size
    Dm = sp.diags([d, e[:-1], e[:-1]], [0, -1, 1], shape=(N, N), format="csc")
    Lc = sp.csc_matrix(L)
    w = spla.spsolve_triangular(sp.csr_matrix(Lc), z[perm], lower=True, unit_diagonal=True)
    u = spla.spsolve(Dm, w)
    y = spla.spsolve_triangular(sp.csr_matrix(Lc.T), u, lower=False, unit_diagonal=True)
    out = np.empty(N)
    out[perm] = y
    return out


def _check(L, d, e, perm, tol=1e-10, seeds=(0, 1)):
    S = _sweep(L, d, e, perm)
    N = d.size
    for seed in seeds:
        z = np.random.default_rng(seed).standard_normal(N)
        ref = _direct(L, d, e, perm, z)
        got = _walk(S, N, z)
        assert np.abs(got - ref).max() <= tol * np.abs(ref).max()
    return S


def _levels(L):
    Ls = sp.csr_matrix(sp.tril(L, -1))
    lev = np.zeros(Ls.shape[0], dtype=int)
    ip, ix = Ls.indptr, Ls.indices
    for i in range(Ls.shape[0]):
        if ip[i + 1] > ip[i]:
            lev[i] = lev[ix[ip[i]:ip[i + 1]]].max() + 1
    return int(lev.max()) + 1


def test_forest_factor_collapses_to_one_level_per_sweep():
    # BASELINE cfg 3 shape (G diagonal, k = 2): fill-free forest, 10+ levels -> 1 + 1
    w = synth.kkt_lap3d(g=12)
    K = sp.bmat([[w["G"], w["B"].T], [w["B"], -w["C"]]], format="csc")
    L, d, e, perm = ldl_superlu(K)
    S = _check(L, d, e, perm)
    assert _levels(L) > 2
    assert S["lev_f"] == 1 and S["lev_b"] == 1


def test_deep_factor_levels_are_merged_into_groups():
    # k = 6 windowed constraints: filled factor, a long chain of small levels
    w = synth.kkt_lap3d(g=12, k=6, window=64)
    K = sp.bmat([[w["G"], w["B"].T], [w["B"], -w["C"]]], format="csc")
    L, d, e, perm = ldl_superlu(K)
    S = _check(L, d, e, perm, tol=1e-9)
    nl = _levels(L)
    assert S["lev_f"] + S["lev_b"] <= nl, (S["lev_f"], S["lev_b"], nl)       # at least halved (two sweeps of nl levels each before)
    assert S["merged_f"] > 0 and S["merged_b"] > 0


@pytest.mark.parametrize("name", ["cvxqp1_m", "cvxqp2_s"])
def test_example_systems(name):
    s = load_system(name)
    L, d, e, perm = ldl_superlu(kp_of(s))
    _check(L, d, e, perm, tol=1e-9)


def test_2x2_pivots_close_a_group():
    s = load_system("cvxqp2_s")
    L, d, e, perm = ldl_dense_bk(kp_of(s))
    assert np.any(e != 0)
    _check(L, d, e, perm, tol=1e-8)


def test_small_random_kkt():
    for seed in range(3):
        s = small_kkt(seed=seed)
        L, d, e, perm = ldl_superlu(kp_of(s))
        _check(L, d, e, perm, tol=1e-9, seeds=(seed,))
