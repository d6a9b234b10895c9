#!/usr/bin/env python
"""Multi-GPU parity check, run under torchrun on a box with >= 2 GPUs (not collected by pytest: the round's
`-m gpu` run has one GPU; the same host logic is covered there by tests/test_gpu_lineshard.py and, over
gloo, by tests/test_sharding_gloo.py):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/multi_gpu_check.py

1. one image over the ranks by scanline block (sharding.ImageSharder over NCCL: spill rows after every
   field, all_gather of row blocks at the end) == the sequential decode of the same calls on one GPU;
2. an image sequence over the ranks (video.VideoConverter, contiguous frame ranges, NCCL exchanges of the
   halo frames / states) == the same converter on one GPU.
Rank 0 prints one JSON line and writes it to gpurun_out/multi_gpu_check.json."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import pkgload  # noqa: E402

pkgload.load()
import support as S  # noqa: E402
from ntsc_crt_b200 import capi, layout, sharding, video  # noqa: E402


def decode_image(variant, outw, outh, blend, scanlines, img, fields, noise, part_of=None):
    """fields x (modulate, demodulate) on one monitor; part_of(image, batch) -> ImageSharder or None."""
    dev = torch.device("cuda", torch.cuda.current_device())
    b = capi.Batch(variant, 1)
    out = torch.zeros(outh, outw, 4, dtype=torch.uint8, device=dev)
    b.set_monitor(0, out, fmt=layout.PIX_BGRA, noise=noise, blend=blend, scanlines=scanlines)
    b.commit_monitors()
    part = part_of(out, b) if part_of else None
    for it in range(fields):
        b.set_source(0, img, format=layout.PIX_BGRA, as_color=1, field=it & 1, frame=(it >> 1) & 1)
        if part is not None:
            part.fetch_halo_rows()
        b.modulate()
        b.demodulate()
        if part is not None:
            last = b.get_lines(0)[part.hi - 1]
            part.exchange_spill_rows(0 if last.beg < 0 else max(last.beg + 1, last.end - scanlines))
    torch.cuda.synchronize()
    full = part.gather() if part is not None else out
    b.close()
    return full


def main():
    rank, local, world = sharding.rank_info()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    report = {"world": world, "checks": []}
    ok = True

    # ---- 1. one image, scanline blocks
    img = torch.from_numpy(S.rand_image(400, 300, seed=3)).to(dev)
    for variant, outw, outh, blend, scanlines in (("ntsc", 832, 624, 1, 1), ("ntsc", 640, 480, 1, 0),
                                                  ("ntsc_conv", 832, 624, 1, 1), ("ntsc", 400, 1080, 1, 0),
                                                  ("ntsc", 320, 360, 1, 0), ("ntsc", 320, 360, 1, 1)):
        def make(image, batch):
            p = sharding.ImageSharder(image, batch.spec.lines)
            p.apply(batch)
            return p
        got = decode_image(variant, outw, outh, blend, scanlines, img, 8, 5, make)
        want = decode_image(variant, outw, outh, blend, scanlines, img, 8, 5, None)
        same = bool(torch.equal(got, want))
        flags = torch.tensor([int(same)], device=dev)
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        same_all = bool(flags.item())
        ok &= same_all
        report["checks"].append({"what": "one image by scanline blocks", "variant": variant, "out": [outw, outh],
                                 "blend": blend, "scanlines": scanlines, "fields": 8, "identical_on_all_ranks": same_all})

    # ---- 2. an image sequence, contiguous frame ranges per rank
    n, w, h = 24 * world, 320, 240
    rng = np.random.default_rng(0)
    base = S.bars_image(w, h)
    frames = np.stack([np.roll(base, 7 * k, axis=1) ^ np.pad(rng.integers(0, 16, size=(h, w, 3), dtype=np.uint8), ((0, 0), (0, 0), (0, 1)))
                       for k in range(n)])
    lo, hi = sharding.shard_range(n, rank, world)
    for noise in (0, 12):
        vc = video.VideoConverter("ntsc", 640, 480, noise=noise, scanlines=1, segments=6)
        mine = vc.convert(torch.from_numpy(frames[lo:hi]).to(dev), first_frame=lo)
        everyone = sharding.allgather_frames(mine)
        flags = torch.tensor([1], device=dev)
        solo_group = dist.new_group([0])  # collective: every rank makes the call
        if rank == 0:
            # the whole sequence on this GPU alone (no process group inside the converter: pass a 1-rank group)
            vc1 = video.VideoConverter("ntsc", 640, 480, noise=noise, scanlines=1, segments=6)
            want = vc1.convert(torch.from_numpy(frames).to(dev), first_frame=0, group=solo_group)
            flags[0] = int(bool(torch.equal(everyone, want)))
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        same_all = bool(flags.item())
        ok &= same_all
        report["checks"].append({"what": "image sequence by frame ranges", "frames": n, "noise": noise,
                                 "segments_per_rank": 6, "recomputed_segments": int(vc.recomputed),
                                 "identical_to_single_gpu": same_all})
    report["ok"] = bool(ok)
    if rank == 0:
        line = json.dumps(report)
        print(line, flush=True)
        os.makedirs(os.path.join(os.path.dirname(HERE), "gpurun_out"), exist_ok=True)
        with open(os.path.join(os.path.dirname(HERE), "gpurun_out", "multi_gpu_check.json"), "w") as f:
            f.write(line + "\n")
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
