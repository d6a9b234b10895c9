"""N > 1 host logic on CPU: world_size 2 over gloo (frames shard, timings reduce with max,
the optional all_gather of decoded frames reassembles the batch in rank order)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import pkgload

pkgload.load()
from ntsc_crt_b200 import sharding


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 8, 296, 4999):
        for world in (1, 2, 3, 4, 8):
            spans = [sharding.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and b - a >= d - c >= b - a - 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import pkgload as pl
    pl.load()
    from ntsc_crt_b200 import sharding as sh
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert sh.rank_info() == (rank, rank, world)
        lo, hi = sh.shard_range(10, rank, world)
        # every rank "decodes" its slice of 10 frames: frame k is filled with k
        local = torch.stack([torch.full((4, 6, 4), k, dtype=torch.uint8) for k in range(lo, hi)])
        allf = sh.allgather_frames(local)
        ms = sh.max_over_ranks([10.0 + rank, 5.0 - rank])
        q.put((rank, allf.numpy().copy(), ms))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = np.stack([np.full((4, 6, 4), k, dtype=np.uint8) for k in range(10)])
    for rank, allf, ms in results:
        assert np.array_equal(allf, want)
        assert ms == [11.0, 5.0]
