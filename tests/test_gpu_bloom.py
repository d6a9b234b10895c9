"""the reference's CRT_DO_BLOOM 1 build (crt_core.h:70; crt_core.c:399-402, 512-531; crt_ntsc.c:148-161)
-- every line's width follows a filtered beam energy carried from line to line -- through
libcrt_b200_ntsc_bloom.so, bit for bit against the oracle and a reference compiled with the option on."""
import numpy as np
import pytest

import support as S
from ntsc_crt_b200 import layout
from test_gpu_parity import check, run_all, trio

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("outw,outh,raw", [(832, 624, 0), (640, 480, 0), (333, 250, 1)])
def test_dropin_bloom(outw, outh, raw):
    img = S.bars_image(300, 260) if not raw else S.rand_image(200, 180, seed=5)
    gpu, ora, ref = trio("ntsc_bloom", outw, outh)
    run_all((gpu, ora, ref), lambda e: e.set(blend=1, scanlines=1, brightness=4, contrast=190))
    for it in range(5):
        run_all((gpu, ora, ref), lambda e: e.modulate(img, format=layout.PIX_BGRA, as_color=1, raw=raw,
                                                      field=it & 1 if not raw else 0, frame=(it >> 1) & 1))
        check(gpu, ora, ref, "bloom mod %d" % it)
        run_all((gpu, ora, ref), lambda e: e.demodulate(0 if it < 2 else 20))
        check(gpu, ora, ref, "bloom demod %d" % it)


@pytest.mark.parametrize("fmt,outw,outh,blend,scanlines", [(layout.PIX_RGB, 401, 300, 1, 0), (layout.PIX_ARGB, 256, 100, 1, 0),
                                                          (layout.PIX_ABGR, 1921, 241, 0, 1), (layout.PIX_BGR, 97, 31, 0, 0),
                                                          (layout.PIX_RGBA, 640, 720, 1, 1)])
def test_dropin_bloom_formats_and_geometries(fmt, outw, outh, blend, scanlines):
    """3-byte pixels, fewer output rows than decoded lines (lines share rows: ordered passes), more than two rows
    per line, knobs outside the packed equaliser's range, a bright picture (wide lines) after a dark one (narrow)"""
    dark = np.zeros((240, 320, 4), dtype=np.uint8)
    bright = np.full((240, 320, 4), 255, dtype=np.uint8)
    gpu, ora, ref = trio("ntsc_bloom", outw, outh, fmt)
    run_all((gpu, ora, ref), lambda e: e.set(blend=blend, scanlines=scanlines, saturation=25, brightness=-30, contrast=260, hue=77))
    for it, img in enumerate([dark, bright, S.rand_image(500, 300, seed=fmt), bright, dark]):
        run_all((gpu, ora, ref), lambda e: e.modulate(img, format=layout.PIX_BGRA, as_color=1, field=it & 1, frame=0,
                                                      xoffset=4 * (it & 1), yoffset=it % 2))
        run_all((gpu, ora, ref), lambda e: e.demodulate(7 * it))
        check(gpu, ora, ref, "bloom fmt %d %dx%d call %d" % (fmt, outw, outh, it))
    run_all((gpu, ora, ref), lambda e: e.set(saturation=900, brightness=5000))
    run_all((gpu, ora, ref), lambda e: e.demodulate(3))
    check(gpu, ora, ref, "bloom extreme knobs")


def test_batch_bloom_matches_oracle():
    import torch
    from ntsc_crt_b200 import capi
    n = 4
    b = capi.Batch("ntsc_bloom", n)
    outs, oras, imgs = [], [], []
    for i in range(n):
        t = torch.zeros(480, 640, 4, dtype=torch.uint8, device="cuda")
        outs.append(t)
        b.set_monitor(i, t, fmt=layout.PIX_BGRA, noise=20 * i, blend=i & 1, scanlines=1 - (i >> 1), contrast=170 + 10 * i)
        o = S.OracleEngine("ntsc_bloom", 640, 480)
        o.set(blend=i & 1, scanlines=1 - (i >> 1), contrast=170 + 10 * i)
        oras.append(o)
        imgs.append(S.rand_image(256 + 32 * i, 224, seed=400 + i) if i else np.full((200, 300, 4), 250, dtype=np.uint8))
    b.commit_monitors()
    dimgs = [torch.from_numpy(im).cuda() for im in imgs]
    for it in range(4):
        for i in range(n):
            kw = dict(format=layout.PIX_BGRA, as_color=1, hue=15 * i, field=it & 1, frame=(it >> 1) & 1)
            b.set_source(i, dimgs[i], **kw)
            oras[i].modulate(imgs[i], **kw)
            oras[i].demodulate(20 * i)
        b.modulate()
        b.demodulate()
        torch.cuda.synchronize()
        st = b.get_state()
        for i in range(n):
            got = outs[i].cpu().numpy()
            assert np.array_equal(got, oras[i].out), "bloom batch monitor %d field %d: %s" % (
                i, it, S.diff_report("out", got, oras[i].out))
            assert np.array_equal(b.signal(i, "inp"), oras[i].inp), (i, it)
            assert (st[i].hsync, st[i].vsync, st[i].rn) == (oras[i].hsync, oras[i].vsync, oras[i].rn), (i, it)
    b.close()
