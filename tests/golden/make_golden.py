#!/usr/bin/env python
"""Generate tests/golden/golden.json from the UNMODIFIED reference (oracle/_ref/libref_*.so).

The reference ships no golden vectors; these are produced here, where
/root/reference is mounted, by driving the compiled reference through seeded synthetic inputs, and
committed so that the oracle (CPU suite) and the CUDA library (GPU suite) can be checked against the
reference even where neither /root/reference nor oracle/_ref exists.  For every case the file holds
sha256 digests of analog / inp / out after the LAST call plus the scalar state after EVERY call.

    python tests/golden/make_golden.py          # rewrites tests/golden/golden.json
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import support as S  # noqa: E402
from ntsc_crt_b200 import layout  # noqa: E402

# (name, variant, outw, outh, fmt, knobs, image spec, calls [(settings, noise)])
CASES = []


def case(name, variant, outw, outh, fmt, knobs, image, calls):
    CASES.append(dict(name=name, variant=variant, outw=outw, outh=outh, fmt=fmt, knobs=knobs,
                      image=image, calls=calls))


def rgb_calls(n, noise, fmt=layout.PIX_BGRA, **kw):
    return [(dict(format=fmt, as_color=kw.get("as_color", 1), field=i & 1, frame=(i >> 1) & 1,
                  raw=kw.get("raw", 0), hue=kw.get("hue", 0), xoffset=kw.get("xoffset", 0),
                  yoffset=kw.get("yoffset", 0)), noise) for i in range(n)]


case("cfg1_256x240_progressive", "ntsc", 832, 624, layout.PIX_BGRA, dict(blend=1, scanlines=1),
     ("lcg", 256, 240, 4, 12345), [(dict(format=layout.PIX_BGRA, as_color=1, field=0, frame=0), 0)] * 4)
case("cfg2_832x624_interlaced", "ntsc", 832, 624, layout.PIX_BGRA, dict(blend=1, scanlines=1),
     ("rand", 832, 624, 4, 7), rgb_calls(8, 0))
case("cfg4_640x480_video_noise12", "ntsc", 640, 480, layout.PIX_BGRA, dict(blend=0, scanlines=1),
     ("bars", 640, 480, 4, 0), rgb_calls(6, 12))
case("noise255", "ntsc", 640, 480, layout.PIX_BGRA, dict(blend=0, scanlines=1),
     ("bars", 640, 480, 4, 0), rgb_calls(3, 255))
for f in range(6):
    case("pixfmt%d" % f, "ntsc", 400, 300, f, dict(blend=1, scanlines=0),
         ("randfmt", 320, 200, f, f), rgb_calls(3, 5, fmt=f))
case("knobs_raw_hue", "ntsc", 512, 448, layout.PIX_BGRA,
     dict(hue=37, brightness=9, contrast=200, saturation=14, black_point=3, white_point=90, blend=0, scanlines=1),
     # raw + odd field reads source row h, one past the image, in the reference (crt_ntsc.c:263,
     # undefined) whenever h <= 236 -- so the raw case stays on the even field
     ("bars", 300, 200, 4, 0),
     [(dict(format=layout.PIX_BGRA, as_color=1, field=0, frame=i & 1, raw=1, hue=20, xoffset=8, yoffset=2), 3)
      for i in range(4)])
case("mono", "ntsc", 512, 448, layout.PIX_BGRA, dict(blend=0, scanlines=1),
     ("bars", 300, 200, 4, 0), rgb_calls(3, 0, as_color=0))
case("generic_eq_saturation400", "ntsc", 320, 240, layout.PIX_BGRA,
     dict(saturation=400, brightness=5000, blend=0, scanlines=0),
     ("rand", 256, 240, 4, 3), rgb_calls(2, 2))
# the USE_CONVOLUTION 1 decoder (crt_core.c:85-147), from libref_ntsc_conv.so
case("conv_cfg2_832x624_interlaced", "ntsc_conv", 832, 624, layout.PIX_BGRA, dict(blend=1, scanlines=1),
     ("rand", 832, 624, 4, 7), rgb_calls(6, 0))
case("conv_640x480_noise12_rgb", "ntsc_conv", 640, 480, layout.PIX_RGB, dict(blend=0, scanlines=1),
     ("bars", 640, 480, 4, 0), rgb_calls(4, 12))
case("conv_generic_saturation4000", "ntsc_conv", 333, 250, layout.PIX_ARGB,
     dict(saturation=4000, contrast=900, brightness=5000, blend=1, scanlines=0),
     ("rand", 256, 240, 4, 3), rgb_calls(3, 9))
for v in ("ntsc_conv6", "ntsc_conv5", "ntsc_conv4"):  # the other kernels of the option (crt_core.c:86-88)
    case("%s_640x480" % v, v, 640, 480, layout.PIX_BGRA, dict(blend=1, scanlines=1, hue=-20, saturation=13),
         ("rand", 333, 250, 4, 11), rgb_calls(4, 6))
case("snes_640x480", "snes", 640, 480, layout.PIX_BGRA, dict(blend=1, scanlines=1, saturation=12),
     ("rand", 300, 230, 4, 21),
     [(dict(format=layout.PIX_BGRA, as_color=1, raw=0, field=i & 1, frame=0, hue=(i * 50) % 360,
            dot_crawl_offset=i % 3, xoffset=4 * (i & 1), yoffset=i % 3), 3 * i) for i in range(4)])
case("nesrgb_640x480", "nesrgb", 640, 480, layout.PIX_BGRA, dict(blend=1, scanlines=1, saturation=11),
     ("rand", 256, 240, 4, 23),
     [(dict(format=layout.PIX_BGRA, hue=(i * 70) % 360, dot_crawl_offset=i % 3, xoffset=4 * (i & 1), yoffset=i % 2), 2 * i)
      for i in range(4)])
# CRT_SYSTEM_TEMP (crt_template.c), from libref_template.so
case("template_640x480", "template", 640, 480, layout.PIX_BGRA, dict(blend=1, scanlines=1, hue=-15, saturation=12),
     ("rand", 300, 260, 4, 25),
     [(dict(format=layout.PIX_BGRA, as_color=1, raw=0, field=i & 1, frame=(i >> 1) & 1, hue=(i * 50) % 360,
            dot_crawl_offset=i % 4, xoffset=4 * (i & 1), yoffset=i % 3), 3 * i) for i in range(6)])
case("template_wide_mono_rgb", "template", 333, 250, layout.PIX_RGB, dict(blend=0, scanlines=0, white_point=90),
     ("rand", 1920, 400, 4, 26),
     [(dict(format=layout.PIX_BGRA, as_color=i & 1, raw=0, field=i & 1, frame=0, hue=0, dot_crawl_offset=i, xoffset=0,
            yoffset=0), 7) for i in range(3)])
# CRT_SYSTEM_PV1K (crt_pv1k.c, CRT_CC_SAMPLES 5), from libref_pv1k.so.  The 832x624 interlaced case
# has decode windows that run a few samples past inp[] into outw / outh (crt_core.h:74-92): part of the parity domain.
case("pv1k_832x624_interlaced", "pv1k", 832, 624, layout.PIX_BGRA, dict(blend=1, scanlines=1, hue=25, saturation=11),
     ("rand", 300, 260, 4, 27),
     [(dict(format=layout.PIX_BGRA, as_color=1, raw=0, field=i & 1, frame=(i >> 1) & 1, hue=(i * 50) % 360,
            dot_crawl_offset=i % 4, xoffset=4 * (i & 1), yoffset=i % 3), 0 if i < 2 else 9) for i in range(6)])
case("pv1k_333x250_rgb_generic", "pv1k", 333, 250, layout.PIX_RGB, dict(blend=0, scanlines=0, saturation=400, brightness=3000),
     ("bars", 400, 300, 4, 0),
     [(dict(format=layout.PIX_BGRA, as_color=1, raw=0, field=i & 1, frame=0, hue=100 * i, dot_crawl_offset=i, xoffset=0,
            yoffset=0), 4 * i) for i in range(3)])
# the CRT_DO_BLOOM 1 build (crt_core.h:70), from libref_ntsc_bloom.so
case("bloom_832x624_interlaced", "ntsc_bloom", 832, 624, layout.PIX_BGRA, dict(blend=1, scanlines=1, brightness=4, contrast=190),
     ("bars", 300, 260, 4, 0), rgb_calls(5, 20))
case("bloom_401x300_rgb", "ntsc_bloom", 401, 300, layout.PIX_RGB, dict(blend=1, scanlines=0, saturation=25, hue=77),
     ("rand", 500, 300, 4, 31), rgb_calls(3, 7))
for v in ("nes", "nes_p0", "nes_p1"):
    case("cfg3_%s" % v, v, 832, 624, layout.PIX_BGRA, dict(blend=0, scanlines=1),
         ("nes", 256, 240, 0, 5), [(dict(dot_crawl_offset=i % 3, hue=(i * 30) % 360), 4 * i) for i in range(5)])
case("cfg5_vhs_colour", "vhs", 832, 624, layout.PIX_BGRA, dict(blend=1, scanlines=1),
     ("bars", 832, 624, 4, 0), [(dict(format=layout.PIX_BGRA, as_color=1, field=i & 1, frame=(i >> 1) & 1,
                                      do_aberration=0), 24) for i in range(4)])
case("cfg5_vhs_mono", "vhs", 832, 624, layout.PIX_BGRA, dict(blend=1, scanlines=1),
     ("bars", 832, 624, 4, 0), [(dict(format=layout.PIX_BGRA, as_color=0, field=i & 1, frame=0,
                                      do_aberration=0), 24) for i in range(3)])


def make_image(spec):
    kind, w, h, a, seed = spec
    if kind == "lcg":
        return S.lcg_image(w, h, a, seed)
    if kind == "rand":
        return S.rand_image(w, h, a, seed)
    if kind == "bars":
        return S.bars_image(w, h)
    if kind == "randfmt":
        return S.pack_rgb(S.rand_image(w, h, bpp=3, seed=seed), a)
    if kind == "nes":
        return S.nes_image(w, h, seed=seed)
    raise ValueError(kind)


def sha(a):
    return hashlib.sha256(a.tobytes()).hexdigest()


def run_case(c, make_engine):
    """Drive one engine through a case; returns the record golden.json stores."""
    eng = make_engine(c["variant"], c["outw"], c["outh"], c["fmt"])
    eng.set(**c["knobs"])
    img = make_image(c["image"])
    scal = []
    for settings, noise in c["calls"]:
        eng.modulate(img, **settings)
        eng.demodulate(noise)
        scal.append([eng.hsync, eng.vsync, eng.rn, eng.ccf.tolist()])
    st = eng.state()
    return dict(analog=sha(st["analog"]), inp=sha(st["inp"]), out=sha(st["out"]), scalars=scal)


def main():
    out = {}
    for c in CASES:
        out[c["name"]] = run_case(c, lambda v, w, h, f: S.RefEngine(v, w, h, f, seed=1))
        print(c["name"], out[c["name"]]["out"][:16])
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
