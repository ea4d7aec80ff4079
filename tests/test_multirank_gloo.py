"""world_size-2 gloo test (CPU) of the multi-rank host logic: block partition of a
batch of independent systems over ranks, gather of per-system results, reduction
of the convergence flags -- the only collectives of the multi-GPU path."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cpkrylov_b200.batch import partition


def test_partition_covers_everything_once():
    for count in (1, 7, 256, 1000):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                lo, hi = partition(count, world, r)
                assert 0 <= lo <= hi <= count
                seen += list(range(lo, hi))
            assert seen == list(range(count))
            sizes = [partition(count, world, r)[1] - partition(count, world, r)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, count, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = partition(count, world, rank)
    # stand-in for the per-rank batch solve: system j "converges" in 10+j iterations
    niters = torch.zeros(count, dtype=torch.int64)
    solved = torch.ones(count, dtype=torch.int64)
    for j in range(lo, hi):
        niters[j] = 10 + j
        solved[j] = 0 if j == 5 else 1
    dist.all_reduce(niters, op=dist.ReduceOp.SUM)          # gather by disjoint ownership
    flag = solved[lo:hi].min() if hi > lo else torch.tensor(1)
    flag = flag.clone()
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)            # global "all solved"
    total = niters.sum().clone()
    if rank == 0:
        out.put((niters.tolist(), int(flag), int(total)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_ranks_gather_and_reduce():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    count = 9
    procs = [ctx.Process(target=_worker, args=(r, 2, port, count, q)) for r in range(2)]
    for p in procs:
        p.start()
    niters, flag, total = q.get(timeout=90)
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    assert niters == [10 + j for j in range(count)]
    assert flag == 0                                        # system 5 did not converge
    assert total == sum(10 + j for j in range(count))
