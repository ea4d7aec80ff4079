"""Host-side check of the compact-walk stream (cpk_device.cuh: DevCompact).

The stream builder runs on the host at `cpk_ldl2_create`; `cpk_debug_cw_stream`
returns the same bytes without touching a device.  This file decodes the blocks
with the layout documented in cpk_device.cuh, walks them exactly like
`ldl_solve_compact` does (groups in order, every item of a group against the state
left by the previous groups) and compares with a direct solve of P L D L' P' y = z.
"""
import ctypes as ct

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from cpkrylov_b200 import _lib
from cpkrylov_b200.ldl import ldl_dense_bk, ldl_superlu
from helpers import kp_of, load_factors, load_system, small_kkt

BLK = 16384
NW = 12                              # warps that walk the stream (kCwWarps) = slots per step
CW_ROWS2, CW_ROWS, CW_WARPROW, CW_DCHUNK = 1, 2, 3, 4
CW_BARRIER = 16
STAGE = 512                          # kCwStage


def _stream(L, d, e, perm):
    lib = _lib.lib()
    fn = lib.cpk_debug_cw_stream
    fn.restype = ct.c_int
    fn.argtypes = [ct.POINTER(_lib.CscStruct), ct.POINTER(_lib.CscStruct), ct.POINTER(ct.c_int64),
                   ct.c_void_p, ct.c_int64, ct.POINTER(ct.c_int64)]
    N = d.size
    e = np.concatenate([np.asarray(e, dtype=np.float64), np.zeros(N)])[:N]
    D = sp.diags([d, e[:N - 1], e[:N - 1]], [0, -1, 1], shape=(N, N), format="csc")
    D.eliminate_zeros()
    Lc, Dc = _lib.Csc(L), _lib.Csc(D)
    perm = np.ascontiguousarray(perm, dtype=np.int64)
    pp = perm.ctypes.data_as(ct.POINTER(ct.c_int64))
    nb = ct.c_int64(0)
    _lib.check(fn(Lc.ref(), Dc.ref(), pp, None, 0, ct.byref(nb)))
    buf = np.zeros(nb.value, dtype=np.uint8)
    _lib.check(fn(Lc.ref(), Dc.ref(), pp, buf.ctypes.data, nb.value, ct.byref(nb)))
    return buf


def _run_slot(blk, slot, warp, sv, yoff, written):
    """One slot of a step; `written` = sv indices produced since the last barrier: nothing
    in the same dependency level may read or rewrite them."""
    off, y, z, _ = (int(v) for v in slot)
    kind, width, stride = y & 15, (y >> 8) & 0xffff, (y >> 24) & 0xff
    if kind == 0:
        return 0

    def publish(idx, vals):
        idx = np.atleast_1d(idx)
        assert not (set(idx.tolist()) & written), "two writers in one level"
        sv[idx] = vals
        written.update(idx.tolist())

    def gather(cols):
        used = cols[cols >= 0]
        assert not (set(used.tolist()) & written), "read of a value produced in the same level"
        return np.where(cols >= 0, sv[np.maximum(cols, 0)], 0.0)

    if kind == CW_DCHUNK:
        n, has2, i0 = width, stride, z
        lo, hi = 32 * warp, min(32 * warp + 32, n)
        if lo >= hi:
            return 0
        dd = blk[off:off + 8 * n].view(np.float64)
        rows = np.arange(i0 + lo, i0 + hi)
        w = gather(rows)
        if not has2:
            publish(yoff + rows, w / dd[lo:hi])
            return 0
        o = off + 8 * n
        ee = blk[o:o + 8 * n].view(np.float64); o += 8 * n
        dp = blk[o:o + 8 * n].view(np.float64); o += 8 * n
        pr = blk[o:o + 4 * n].view(np.int32)
        out = np.empty(hi - lo)
        for r in range(lo, hi):
            if pr[r] < 0:
                out[r - lo] = w[r - lo] / dd[r]
            else:
                det = dd[r] * dp[r] - ee[r] * ee[r]
                out[r - lo] = (dp[r] * w[r - lo] - ee[r] * gather(np.array([pr[r]]))[0]) / det
        publish(yoff + rows, out)
        return 0
    S = stride
    assert off % 16 == 0 and 1 <= S <= 32
    if kind == CW_ROWS2:
        assert S % 4 == 0 and width <= 2
        rec = blk[off:off + 32 * S]
        ri = rec.view(np.int32).reshape(S, 8)[:, :4]
        rv = rec.view(np.float64).reshape(S, 4)[:, 2:]
        tgt = ri[:, 0]; col = ri[:, 1:3].T; val = rv.T
        assert np.all(ri[:, 3] == 0)
    else:
        o = off
        if kind == CW_ROWS:
            assert S % 4 == 0 and 2 < width <= 8
            tgt = blk[o:o + 4 * S].view(np.int32); o += 4 * S
        else:
            assert kind == CW_WARPROW and S == 32 and width <= 16
        val = blk[o:o + 8 * S * width].view(np.float64).reshape(width, S); o += 8 * S * width
        col = blk[o:o + 4 * S * width].view(np.int32).reshape(width, S)
    part = -(val * gather(col) * (col >= 0))
    if kind == CW_WARPROW:
        publish(z, sv[z] + part.sum())
    else:
        live = tgt >= 0
        publish(tgt[live], sv[tgt[live]] + part.sum(axis=0)[live])
    return 1


def _walk(buf, N, perm, z, yoff=None):
    """Emulation of ldl_solve_compact: per block a table of steps x 16 warp slots; the slots
    of the steps between two barriers form one dependency level and must be independent.
    yoff: where y lives in the shared vector (0: one vector updated in place -- diagonal D;
    N: separate w and y -- 2x2 pivots); default = what the builder chooses for a diagonal D."""
    assert buf.size % BLK == 0 and buf.size > 0
    yoff = 0 if yoff is None else yoff
    sv = np.zeros(N + yoff + STAGE)          # sweep values + the staging slots of merged chains (zero when a solve starts)
    sv[:N] = z[perm]
    stats = dict(blocks=buf.size // BLK, steps=0, items=0, barriers=0)
    written = set()
    for b in range(buf.size // BLK):
        blk = buf[b * BLK:(b + 1) * BLK]
        i32 = blk.view(np.int32)
        nsteps = int(i32[0])
        assert nsteps >= 1 and 16 + 16 * NW * nsteps <= BLK
        table = i32[4:4 + 4 * NW * nsteps].reshape(nsteps, NW, 4)
        for st in range(nsteps):
            flags = table[st, :, 1] & CW_BARRIER
            assert flags.min() == flags.max(), "barrier flag differs between the warps of a step"
            for w in range(NW):
                stats["items"] += _run_slot(blk, table[st, w], w, sv, yoff, written)
            stats["steps"] += 1
            if flags[0]:
                stats["barriers"] += 1
                written.clear()
    assert not written                  # the last step ends with a barrier
    y = np.zeros(N)
    y[perm] = sv[yoff:yoff + N]
    return y, stats


def _direct(L, d, e, perm, z):
    N = d.size
    e = np.concatenate([np.asarray(e, dtype=np.float64), np.zeros(N)])[:N]
    Lu = sp.csr_matrix(sp.tril(L, -1) + sp.identity(N))
    D = sp.diags([d, e[:N - 1], e[:N - 1]], [0, -1, 1], shape=(N, N), format="csc")
    w = spla.spsolve_triangular(Lu, z[perm], lower=True)
    v = spla.spsolve(D, w)
    y = spla.spsolve_triangular(Lu.T.tocsr(), v, lower=False)
    out = np.zeros(N)
    out[perm] = y
    return out


@pytest.fixture(scope="module")
def cpk_lib():
    try:
        return _lib.lib()
    except _lib.CpkLibraryMissing:
        pytest.skip("libcpk_b200.so not built")


@pytest.mark.parametrize("name,kind", [("cvxqp2_s", "superlu"), ("cvxqp2_s", "densebk"), ("cvxqp1_m", "superlu")])
def test_stream_walk_equals_direct_solve_on_fixtures(cpk_lib, name, kind):
    L, d, e, perm = load_factors(name, kind)
    N = d.size
    rng = np.random.default_rng(5)
    z = rng.standard_normal(N)
    buf = _stream(L, d, e, perm)
    y, st = _walk(buf, N, perm, z, yoff=N if np.any(e) else 0)
    ref = _direct(L, d, e, perm, z)
    assert np.linalg.norm(y - ref) <= 1e-9 * np.linalg.norm(ref), st
    # and it really inverts K_P
    s = load_system(name)
    KP = kp_of(s)
    assert np.linalg.norm(KP @ y - z) <= 1e-6 * np.linalg.norm(z)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_stream_walk_small_random(cpk_lib, seed):
    s = small_kkt(n=90, m=35, seed=seed)
    KP = kp_of(s)
    for fac in (ldl_superlu, ldl_dense_bk):
        L, d, e, perm = fac(KP)
        z = np.random.default_rng(seed).standard_normal(s["N"])
        y, st = _walk(_stream(L, d, e, perm), s["N"], np.asarray(perm), z, yoff=s["N"] if np.any(e) else 0)
        assert np.linalg.norm(KP @ y - z) <= 1e-8 * np.linalg.norm(z), (fac.__name__, st)


def test_stream_long_rows_are_split(cpk_lib):
    """A dense factor: rows longer than one item (512 entries) are split into ordered parts."""
    rng = np.random.default_rng(3)
    N = 700
    Ld = np.tril(rng.standard_normal((N, N)) * 0.02, -1)
    L = sp.csr_matrix(Ld + np.eye(N))
    d = rng.uniform(1.0, 2.0, N) * np.where(np.arange(N) % 3 == 0, -1.0, 1.0)
    e = np.zeros(N)
    perm = rng.permutation(N).astype(np.int64)
    z = rng.standard_normal(N)
    ref = _direct(L, d, e, perm, z)
    import os
    os.environ["CPK_CW_NO_CHAINS"] = "1"
    try:
        y, st = _walk(_stream(L, d, e, perm), N, perm, z)
    finally:
        os.environ.pop("CPK_CW_NO_CHAINS")
    assert np.linalg.norm(y - ref) <= 1e-9 * np.linalg.norm(ref)
    assert st["barriers"] > 2 * N    # one row per level; rows longer than 512 entries take several
    # with chain merging: runs of one-row levels collapse (as far as the staging area and the guards allow)
    y2, st2 = _walk(_stream(L, d, e, perm), N, perm, z)
    assert np.linalg.norm(y2 - ref) <= 1e-9 * np.linalg.norm(ref)
    assert st2["barriers"] < st["barriers"]


def test_chains_of_tiny_levels_are_merged(cpk_lib):
    """cvxqp1 (SuperLU factor): 47 consecutive one-row levels at the end of the forward sweep --
    merged into two levels through the staging area; same result, far fewer barriers."""
    import os
    L, d, e, perm = load_factors("cvxqp1_m", "superlu")
    N = d.size
    z = np.random.default_rng(11).standard_normal(N)
    ref = _direct(L, d, e, perm, z)
    os.environ["CPK_CW_NO_CHAINS"] = "1"
    try:
        y0, st0 = _walk(_stream(L, d, e, perm), N, perm, z)
    finally:
        os.environ.pop("CPK_CW_NO_CHAINS")
    y1, st1 = _walk(_stream(L, d, e, perm), N, perm, z)
    assert np.linalg.norm(y0 - ref) <= 1e-9 * np.linalg.norm(ref)
    assert np.linalg.norm(y1 - ref) <= 1e-9 * np.linalg.norm(ref)
    assert st1["barriers"] <= st0["barriers"] - 40, (st0, st1)
    print("compact walk of cvxqp1: without / with chain merging", st0, st1)


def test_stream_diagonal_factor(cpk_lib):
    N = 40
    L = sp.identity(N, format="csr")
    d = np.linspace(1.0, 2.0, N)
    perm = np.arange(N, dtype=np.int64)[::-1].copy()
    z = np.arange(N, dtype=np.float64)
    y, st = _walk(_stream(L, d, np.zeros(N), perm), N, perm, z)
    assert st["blocks"] == 1 and st["items"] == 0 and st["barriers"] == 1
    ref = np.zeros(N); ref[perm] = z[perm] / d
    assert np.array_equal(y, ref)
