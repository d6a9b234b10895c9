"""video.VideoConverter's scheduling on the CPU: the batch interface is stood in for by the oracle
(tests/mock_batch.py), single process and world sizes 2 and 3 over gloo, with noise levels at which the speculated
sync state fails and segments -- including the ones that straddle ranks -- must be repaired.  The result has to be
the sequential loop's (extra/video_convert.c:226-277), image for image.  The GPU suite runs the same class on the
CUDA library (tests/test_gpu_video.py, tests/multi_gpu_check.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import support as S
from ntsc_crt_b200 import layout, sharding, video


def make_frames(n, w=160, h=240):
    rng = np.random.default_rng(9)
    base = S.bars_image(w, h)
    out = []
    for k in range(n):
        f = np.roll(base, 5 * k, axis=1).copy()
        if k % 4 == 2:
            f[..., :3] = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
        out.append(f)
    return np.stack(out)


def sequential(frames, outw, outh, noise):
    ora = S.OracleEngine("ntsc", outw, outh)
    ora.set(blend=0, scanlines=1, saturation=10)
    want = []
    for f in range(len(frames)):
        field, frame = video.frame_parity(f)
        ora.modulate(frames[f], format=layout.PIX_BGRA, as_color=1, field=field, frame=frame)
        ora.demodulate(noise)
        want.append(ora.out.copy())
    return np.stack(want)


@pytest.mark.parametrize("noise,segments,n", [(0, 4, 13), (200, 5, 16), (255, 16, 16)])
def test_single_process(noise, segments, n):
    from mock_batch import OracleBatch
    frames = make_frames(n)
    vc = video.VideoConverter("ntsc", 320, 240, noise=noise, scanlines=1, segments=segments, batch_factory=OracleBatch)
    got = vc.convert(torch.from_numpy(frames)).numpy()
    want = sequential(frames, 320, 240, noise)
    for f in range(n):
        assert np.array_equal(got[f], want[f]), "image %d (redone %d)" % (f, vc.recomputed)
    if noise == 0:
        assert vc.recomputed == 0
    else:
        assert vc.recomputed > 0, "these noise levels are meant to break the speculation"


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q, noise, segments, n):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import pkgload
    pkgload.load()
    from mock_batch import OracleBatch
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        frames = make_frames(n)
        lo, hi = sharding.shard_range(n, rank, world)
        vc = video.VideoConverter("ntsc", 320, 240, noise=noise, scanlines=1, segments=segments, batch_factory=OracleBatch)
        mine = vc.convert(torch.from_numpy(frames[lo:hi]), first_frame=lo)
        q.put((rank, lo, hi, mine.numpy().copy(), vc.recomputed))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,noise,segments,n", [(2, 0, 3, 12), (2, 255, 3, 16), (3, 200, 2, 18)])
def test_over_ranks(world, noise, segments, n):
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, noise, segments, n)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = sequential(make_frames(n), 320, 240, noise)
    redone = 0
    for rank, lo, hi, got, rec in results:
        redone += rec
        for f in range(lo, hi):
            assert np.array_equal(got[f - lo], want[f]), "rank %d image %d (redone %d)" % (rank, f, rec)
    if noise:
        assert redone > 0
