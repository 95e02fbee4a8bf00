"""N>1 host logic on CPU: world_size-2 gloo processes shard utterances with no overlap and agree on the max-over-ranks
time reduction bench.py uses.  (The GPU data path has no collective: every rank generates its own block.)"""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import srnn_b200 as S


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_utt, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    cond = torch.rand(n_utt, 3, 5, generator=g)
    spk = torch.arange(n_utt)
    uni = torch.rand(12, n_utt, generator=g)
    c, s, u, (lo, hi) = S.shard_batch(cond, spk, uni, rank, world)
    assert c.shape[0] == hi - lo == s.numel() == u.shape[1]
    mine = torch.zeros(n_utt, dtype=torch.int64)
    mine[lo:hi] = 1
    dist.all_reduce(mine)                                   # every utterance owned exactly once
    t = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)                # bench.py: time = max over ranks
    if rank == 0:
        out.put((mine.tolist(), float(t.item()), (lo, hi)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_utterance_sharding():
    world, n_utt = 2, 7
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_utt, q)) for r in range(world)]
    for p in procs:
        p.start()
    owned, tmax, r0 = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert owned == [1] * n_utt
    assert tmax == 11.0
    assert r0 == (0, 4)


def test_shard_range_properties():
    for n in (0, 1, 7, 256, 4096):
        for w in (1, 2, 4, 8):
            edges = [S.shard_range(n, r, w) for r in range(w)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1
