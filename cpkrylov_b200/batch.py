"""Batches of independent KKT systems: the unit that shards across GPUs.

One GPU solves its whole share in ONE launch (one CTA per system,
``cpk_batch_reg_solve``); across GPUs the batch is block-partitioned over ranks
with no collective inside the iteration -- ``torch.distributed`` only gathers the
solutions / iteration counts and reduces the convergence flags afterwards.
"""
from __future__ import annotations

import ctypes as ct

import numpy as np

from . import _lib
from .operators import KktSystem, opLDL2
from .solvers import _fill_opts, apply_opts_to_M


def partition(count, world, rank):
    """Static block partition of `count` systems over `world` ranks: [lo, hi)."""
    base, rem = divmod(count, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class BatchSolver:
    def __init__(self, systems, factors=None, opts=None, device=0):
        self.systems = []
        for i, w in enumerate(systems):
            M = opLDL2(w["G"], w["B"], -w["C"], factors=None if factors is None else factors[i], device=device)
            apply_opts_to_M(M, opts)
            self.systems.append(KktSystem(w["H"], w["C"], M))
        self.last_launches = 0
        self.last_ms = 0.0

    def solve(self, method, rhs, opts=None):
        cnt = len(self.systems)
        if len(rhs) != cnt:
            raise ValueError("batch of %d systems got %d right-hand sides" % (cnt, len(rhs)))
        name = method if isinstance(method, str) else method.cpk_name
        # ONE option set serves the whole launch: the defaults that depend on the size (itmax = n
        # resp. n + m, cpcg.m:99, cpgmres.m:105) are taken from the LARGEST system of the batch
        n = max(S.n for S in self.systems)
        m = max(S.m for S in self.systems)
        sid, o = _fill_opts(name, opts, n, m)
        L = _lib.lib()
        cap = int(L.cpk_hist_capacity(sid, ct.byref(o)))
        handles = (ct.c_uint64 * cnt)(*[S.handle.value for S in self.systems])
        bs = [np.ascontiguousarray(np.asarray(b, dtype=np.float64).reshape(-1)) for b in rhs]
        for i, (b, S) in enumerate(zip(bs, self.systems)):
            if b.size != S.N:
                raise ValueError("right-hand side %d has %d entries, system %d has N = %d" % (i, b.size, i, S.N))
        xs = [np.empty(S.N) for S in self.systems]
        hs = [np.zeros((3, cap)) for _ in range(cnt)]
        bp = (ct.c_void_p * cnt)(*[b.ctypes.data for b in bs])
        xp = (ct.c_void_p * cnt)(*[x.ctypes.data for x in xs])
        hp = (ct.c_void_p * cnt)(*[h.ctypes.data for h in hs])
        st = (_lib.StatsStruct * cnt)()
        l0 = L.cpk_launch_count()
        _lib.check(L.cpk_batch_reg_solve(handles, cnt, sid, bp, ct.byref(o), xp, st, hp, cap))
        self.last_launches = L.cpk_launch_count() - l0
        stats = []
        for i in range(cnt):
            d = _lib.stats_to_dict(st[i])
            d["residHistory"] = hs[i][0, :d["hist_len"]].copy()
            stats.append(d)
        self.last_ms = stats[0]["t_solve_ms"]
        return xs, stats

    def close(self):
        for S in self.systems:
            S.close()
        self.systems = []
