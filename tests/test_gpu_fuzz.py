"""Seeded random sweep of the drop-in path: output geometry, pixel formats, monitor knobs, encoder
settings, noise and field sequences drawn at random (fixed seeds, so every run is the same), CUDA library
against the oracle after every call pair.  Complements the hand-picked cases of test_gpu_parity.py /
test_gpu_conv.py with combinations nobody thought of."""
import numpy as np
import pytest

import support as S
from ntsc_crt_b200 import layout

pytestmark = pytest.mark.gpu


def draw_case(rng, variant):
    fmt = int(rng.integers(0, 6))
    outw = int(rng.choice([64, 97, 256, 320, 333, 400, 512, 640, 641, 832, 1024, 1280, 1921]))
    outh = int(rng.choice([31, 80, 224, 240, 241, 300, 448, 480, 624, 720, 1081]))
    knobs = dict(blend=int(rng.integers(0, 2)), scanlines=int(rng.integers(0, 2)),
                 hue=int(rng.integers(-400, 400)), brightness=int(rng.integers(-60, 60)),
                 contrast=int(rng.integers(60, 320)), saturation=int(rng.integers(0, 40)),
                 black_point=int(rng.integers(-10, 20)), white_point=int(rng.integers(50, 130)))
    if rng.random() < 0.15:  # far outside the packed path's exact range
        knobs.update(saturation=int(rng.integers(300, 5000)), brightness=int(rng.integers(-6000, 6000)),
                     contrast=int(rng.integers(300, 1200)))
    if variant.startswith("nes"):
        w, h = 256, 240
    else:
        w, h = int(rng.integers(40, 900)), int(rng.integers(240, 700))
    return fmt, outw, outh, knobs, w, h


@pytest.mark.parametrize("variant,seed", [("ntsc", 1), ("ntsc", 2), ("ntsc", 3), ("ntsc_conv", 4), ("ntsc_conv", 5),
                                          ("nes", 6), ("nes_p0", 7), ("ntsc", 8), ("ntsc_conv", 9), ("snes", 10),
                                          ("ntsc_conv5", 11), ("template", 12), ("pv1k", 13), ("ntsc_bloom", 14),
                                          ("pv1k", 15), ("template", 16), ("ntsc_bloom", 17)])
def test_random_configurations(variant, seed):
    rng = np.random.default_rng(1000 + seed)
    for case in range(4):
        fmt, outw, outh, knobs, w, h = draw_case(rng, variant)
        gpu = S.ProductEngine(variant, outw, outh, fmt)
        ora = S.OracleEngine(variant, outw, outh, fmt)
        for e in (gpu, ora):
            e.set(**knobs)
        nes = variant.startswith("nes")
        if nes:
            img = rng.integers(0, 512, size=(h, w), dtype=np.uint16)
        else:
            src_fmt = int(rng.integers(0, 6))
            img = S.pack_rgb(S.rand_image(w, h, bpp=3, seed=int(rng.integers(0, 1 << 30))), src_fmt)
        for call in range(3):
            noise = int(rng.choice([0, 0, 3, 12, 40, 255]))
            if nes:
                kw = dict(dot_crawl_offset=int(rng.integers(0, 3)), hue=int(rng.integers(0, 360)),
                          xoffset=int(rng.integers(0, 3)) * 4, yoffset=int(rng.integers(0, 3)))
            else:
                field = int(rng.integers(0, 2))
                kw = dict(format=src_fmt, as_color=int(rng.integers(0, 2)), field=field, frame=int(rng.integers(0, 2)),
                          raw=0, hue=int(rng.integers(0, 360)), xoffset=int(rng.integers(0, 4)) * 4,
                          yoffset=int(rng.integers(0, 3)))
                if variant in ("snes", "template", "pv1k"):
                    kw["dot_crawl_offset"] = int(rng.integers(0, 4))
            for e in (gpu, ora):
                e.modulate(img, **kw)
                e.demodulate(noise)
            S.assert_same_state(gpu.state(), ora.state(),
                                "%s seed %d case %d call %d: %dx%d fmt %d knobs %r settings %r noise %d src %dx%d"
                                % (variant, seed, case, call, outw, outh, fmt, knobs, kw, noise, w, h))
