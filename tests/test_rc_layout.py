"""Host-side check of the row-class layout (cpk_device.cuh: DevRc).

The builder runs on the host at `cpk_ldl2_create`; `cpk_debug_rc` returns the same arrays
without touching a device.  This file walks them exactly like `rc_level` / `ldl_solve_rc` do
(level after level, every group of a level dealt to exactly one warp by the cost split,
nothing produced in a level read inside that level) and compares with a direct solve of
P L D L' P' y = z, resp. with a plain sparse product for a matrix in row-class form.
"""
import ctypes as ct

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from cpkrylov_b200 import _lib, synth
from cpkrylov_b200.ldl import ldl_superlu
from helpers import kp_of, load_factors, load_system, small_kkt

IDX_BITS = 28
MASK = (1 << IDX_BITS) - 1
RC_B = (4, 2, 2, 1)                   # CPK_RC_B: groups per batch for the widths 0..3 (wider: 1)
LONG_W = 255


def _rc(A=None, L=None, d=None, perm=None):
    import os
    os.environ["CPK_LDL_RC"] = "1"      # the row-class form of the sweeps is opt-in (item list by default)
    try:
        return _rc_impl(A, L, d, perm)
    finally:
        os.environ.pop("CPK_LDL_RC", None)


def _rc_impl(A, L, d, perm):
    lib = _lib.lib()
    fn = lib.cpk_debug_rc
    fn.restype = ct.c_int
    PC = ct.POINTER(_lib.CscStruct)
    fn.argtypes = [PC, PC, PC, ct.POINTER(ct.c_int64), ct.POINTER(ct.c_int64)] + [ct.c_void_p] * 10
    keep = []
    if L is not None:
        N = d.size
        Dm = sp.diags([d], [0], shape=(N, N), format="csc")
        Lc, Dc = _lib.Csc(L), _lib.Csc(Dm)
        perm = np.ascontiguousarray(perm, dtype=np.int64)
        keep += [Lc, Dc, perm]
        args = [None, Lc.ref(), Dc.ref(), perm.ctypes.data_as(ct.POINTER(ct.c_int64))]
    else:
        Ac = _lib.Csc(A)
        keep.append(Ac)
        args = [Ac.ref(), None, None, None]
    sizes = np.zeros(8, dtype=np.int64)
    ps = sizes.ctypes.data_as(ct.POINTER(ct.c_int64))
    _lib.check(fn(*args, ps, *([None] * 10)))
    have, nlev, nfwd, npieces, nrow, ncol, nlptr, nlcol = (int(v) for v in sizes)
    R = dict(have=have, nlev=nlev, nfwd=nfwd,
             pieces=np.zeros((npieces, 4), dtype=np.int32), levp=np.zeros(nlev + 1, dtype=np.int32), levb=np.zeros(max(nlev, 1), dtype=np.int32),
             rowmap=np.zeros(nrow, dtype=np.int32), d=np.ones(nrow), col=np.zeros(ncol, dtype=np.int32),
             val=np.zeros(ncol), lptr=np.zeros(nlptr, dtype=np.int32), lcol=np.zeros(nlcol, dtype=np.int32),
             lval=np.zeros(nlcol))
    ptr = lambda a: a.ctypes.data if a.size else None
    _lib.check(fn(*args, ps, ptr(R["pieces"]), ptr(R["levp"]), ptr(R["levb"]), ptr(R["rowmap"]), ptr(R["d"]), ptr(R["col"]), ptr(R["val"]),
                  ptr(R["lptr"]), ptr(R["lcol"]), ptr(R["lval"])))
    return R


def _batch(w):
    return RC_B[w] if 0 <= w <= 3 else 1


def _cost(w, ch=4):
    """rc_cost_of: round trips per batch"""
    return 1 if w <= 3 else (3 if w == LONG_W else 1 + (w + ch - 1) // ch)


def _split(R, lev, nwarps):
    """(piece, group range) per warp, the way rc_level cuts the batch sequence of a level into
    equal contiguous ranges (integer arithmetic included)"""
    p0, p1 = int(R["levp"][lev]), int(R["levp"][lev + 1])
    T = int(R["levb"][lev])
    out = []
    for gw in range(nwarps):
        q, rem = divmod(T, nwarps)
        lo = q * gw + rem * gw // nwarps
        hi = q * (gw + 1) + rem * (gw + 1) // nwarps
        for p in range(p0, p1):
            wn, cum = int(R["pieces"][p][0]), int(R["pieces"][p][1])
            if cum >= hi:
                break
            w, ng = wn & 255, wn >> 8
            B, c = _batch(w), _cost(w)
            nb = (ng + B - 1) // B
            a = (lo - cum + c - 1) // c if lo > cum else 0
            b = min((hi - cum + c - 1) // c, nb)
            if a < b:
                out.append((p, a * B, min(b * B, ng)))
    return out


def _groups_once(R, lev, nwarps):
    """every group of the level is dealt to exactly one warp; the prefix sums of the piece table
    are the batch counts the device computes"""
    cum = 0
    for p in range(int(R["levp"][lev]), int(R["levp"][lev + 1])):
        wn, c = int(R["pieces"][p][0]), int(R["pieces"][p][1])
        assert c == cum
        cum += ((wn >> 8) + _batch(wn & 255) - 1) // _batch(wn & 255) * _cost(wn & 255)
    assert cum == int(R["levb"][lev])
    seen = {}
    for (p, ga, gb) in _split(R, lev, nwarps):
        for g in range(ga, gb):
            assert (p, g) not in seen
            seen[(p, g)] = 1
    want = sum(int(R["pieces"][p][0]) >> 8 for p in range(int(R["levp"][lev]), int(R["levp"][lev + 1])))
    assert len(seen) == want


def _piece_rows(R, p):
    """yields (position, code, cols, vals) for every live row of piece p"""
    wn, _, row_off, ent_off = (int(v) for v in R["pieces"][p])
    w, ng = wn & 255, wn >> 8
    if w == LONG_W:
        for r in range(ng):
            pos = row_off + r
            b, e = int(R["lptr"][ent_off + r]), int(R["lptr"][ent_off + r + 1])
            yield pos, int(R["rowmap"][pos]), R["lcol"][b:e], R["lval"][b:e]
        return
    stride = ng * 32
    for k in range(stride):
        pos = row_off + k
        code = int(R["rowmap"][pos])
        if code < 0:
            continue
        idx = ent_off + np.arange(w) * stride + k
        yield pos, code, R["col"][idx], R["val"][idx]


def _walk_sweeps(R, z, N):
    """ldl_solve_rc: wv / yv / out indexed by the user index"""
    wy = np.full(2 * N, np.nan); out = np.full(N, np.nan)       # [w | y], user-indexed halves
    for lev in range(R["nlev"]):
        fwd = lev < R["nfwd"]
        new_w, new_y = {}, {}
        for p in range(int(R["levp"][lev]), int(R["levp"][lev + 1])):
            for pos, code, cols, vals in _piece_rows(R, p):
                i, fl = code & MASK, code >> IDX_BITS
                s = 0.0
                for c, v in zip(cols.tolist(), vals.tolist()):
                    isin, j = c >> IDX_BITS, c & MASK
                    x = z[j] if isin else wy[j]
                    assert not np.isnan(x), "gather of a value that no earlier level produced"
                    s += v * x
                if fwd:
                    acc = z[i] - s
                    if fl & 1:
                        acc = acc / R["d"][pos]
                        if fl & 2:
                            new_y[N + i] = acc
                        assert np.isnan(out[i]); out[i] = acc
                    else:
                        assert i not in new_w; new_w[i] = acc
                else:
                    base = z[i] if (fl & 1) else wy[i]
                    assert not np.isnan(base)
                    acc = base / R["d"][pos] - s
                    if fl & 2:
                        new_y[N + i] = acc
                    assert np.isnan(out[i]); out[i] = acc
        # values of a level become visible after its barrier
        for i, v in new_w.items():
            wy[i] = v
        for i, v in new_y.items():
            wy[i] = v
    return out


def _direct(L, d, perm, z):
    N = d.size
    w = spla.spsolve_triangular(sp.csr_matrix(L), z[perm], lower=True, unit_diagonal=True)
    y = spla.spsolve_triangular(sp.csr_matrix(L).T.tocsr(), w / d, lower=False, unit_diagonal=True)
    out = np.empty(N)
    out[perm] = y
    return out


def _check_sweeps(L, d, perm, seed=0):
    R = _rc(L=L, d=d, perm=perm)
    assert R["have"] == 1
    N = d.size
    for lev in range(R["nlev"]):
        for nw in (28, 518, 7):
            _groups_once(R, lev, nw)
    z = np.random.default_rng(seed).standard_normal(N)
    out = _walk_sweeps(R, z, N)
    assert not np.isnan(out).any(), "rows without an output"
    ref = _direct(L, d, perm, z)
    assert np.linalg.norm(out - ref) <= 1e-10 * np.linalg.norm(ref)
    return R


def test_rc_sweeps_cvxqp2_superlu():
    s = load_system("cvxqp2_s")
    L, d, e, perm = ldl_superlu(kp_of(s))
    R = _check_sweeps(L, d, perm)
    assert R["nlev"] >= 2


def test_rc_sweeps_cvxqp1_superlu_deep():
    """69+69 levels in the factor: too deep for the row-class tables (kRcMaxLev = 48) as they stand,
    but level merging (build_sweeps) brings the sweeps down to a dozen groups, so the row-class
    form is built after all and must give the direct solve."""
    s = load_system("cvxqp1_m")
    L, d, e, perm = load_factors("cvxqp1_m", "superlu")
    R = _rc(L=L, d=d, perm=perm)
    assert R["have"] == 1 and R["nlev"] <= 48
    _check_sweeps(L, d, perm, seed=3)


@pytest.mark.parametrize("g,k,window", [(8, 2, 0), (10, 2, 0), (8, 6, 16)])
def test_rc_sweeps_kkt_lap3d(g, k, window):
    w = synth.kkt_lap3d(g=g, k=k, window=window)
    L, d, e, perm = ldl_superlu(synth.kp_matrix(w))
    R = _rc(L=L, d=d, perm=perm)
    if R["have"]:
        _check_sweeps(L, d, perm, seed=g)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_rc_sweeps_small_random(seed):
    s = small_kkt(n=80, m=30, seed=seed)
    L, d, e, perm = ldl_superlu(kp_of(s))
    R = _rc(L=L, d=d, perm=perm)
    if R["have"]:
        _check_sweeps(L, d, perm, seed=seed)


def test_rc_plain_matrix():
    """K_P of a fixture as one level: every row once, entries in CSR order, long rows apart"""
    s = load_system("cvxqp1_m")
    K = kp_of(s).tocsr()
    # add one very long row to reach the CSR (warp per row) list
    K = K.tolil(); K[5, :40] = 1.5; K = K.tocsr()
    R = _rc(A=K)
    assert R["have"] == 1 and R["nlev"] == 1
    for nw in (28, 518):
        _groups_once(R, 0, nw)
    x = np.random.default_rng(1).standard_normal(K.shape[1])
    y = np.full(K.shape[0], np.nan)
    nlong = 0
    for p in range(int(R["levp"][0]), int(R["levp"][1])):
        nlong += (int(R["pieces"][p][0]) & 255) == LONG_W
        for pos, code, cols, vals in _piece_rows(R, p):
            assert np.isnan(y[code])
            # storage order = CSR order of the row
            assert np.array_equal(cols, K.indices[K.indptr[code]:K.indptr[code + 1]])
            y[code] = float(np.dot(vals, x[cols]))
    assert nlong == 1
    assert np.allclose(y, K @ x, rtol=1e-13, atol=1e-13)
