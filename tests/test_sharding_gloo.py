"""N>1 host logic on CPU: batch sharding across ranks and the max-over-ranks
timing reduction bench.py uses, exercised with world_size=2 over gloo."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bench import shard_batch


def test_shard_batch_partitions_exactly():
    for total in (1, 4, 7, 32, 33):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_batch(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
                assert a1 == b0 and a1 >= a0
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    assert [shard_batch(32, 8, r) for r in range(8)] == [(4 * r, 4 * r + 4) for r in range(8)]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        b0, b1 = shard_batch(9, world, rank)
        # every rank reports its shard; together they must tile the batch
        spans = [None] * world
        dist.all_gather_object(spans, (b0, b1))
        # whole-job time = max over ranks, whole-job units = sum over ranks
        t = torch.tensor([10.0 + 5.0 * rank], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        n = torch.tensor([float(b1 - b0)], dtype=torch.float64)
        dist.all_reduce(n, op=dist.ReduceOp.SUM)
        dist.barrier()
        if rank == 0:
            q.put((spans, float(t.item()), float(n.item())))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_reduction():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    spans, tmax, total = q.get()
    assert spans == [(0, 5), (5, 9)]
    assert tmax == 15.0 and total == 9.0
