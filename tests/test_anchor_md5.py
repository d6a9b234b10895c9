"""the five parity anchors -- md5 of the reference CLI's output file for five command lines,
recorded from the reference during the survey.

    ntsc -op 832 624 0 0 601cb.ppm out.ppm        e7bb5f48656e848f61036dc935802ea2
    ntsc -o  832 624 0 0 cbar.bmp out.bmp          aabe954f22b3655e6e876282cbd04590
    ntsc -o  832 624 24 0 cbar.bmp out.bmp         7b0463afc6c41cf120e6839755ed970a
    VHS ntsc -o  832 624 24 0 cbar.bmp out.bmp     ad2f1e4bc7d256c6855345f73ad66bc7   (glibc 2.39, default seed)
    VHS ntsc -om 832 624 24 0 cbar.bmp out.bmp     7a4f94b5e45cd8ae02b83c285159aa12

Three engines must hit them: the compiled reference CLI (oracle/_ref/cli_ref_*, CPU: pins the build recipe), the oracle
replaying the CLI loop (crt_main.c:221-255; first anchor), and -- `-m gpu` -- the UNMODIFIED crt_main.c linked against the
CUDA library (oracle/_ref/cli_b200_*).  The two input patterns are committed under tests/golden/inputs/ (extracted
from the reference's archive by tests/golden/make_anchor_inputs.py) and converted with Pillow exactly as the survey did."""
import hashlib
import os
import subprocess

import numpy as np
import pytest

import support as S
from ntsc_crt_b200 import layout

INPUTS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "inputs")

ANCHORS = [  # (system, flags, noise, input, output name, md5)
    ("ntsc", "-op", 0, "601cb.ppm", "out.ppm", "e7bb5f48656e848f61036dc935802ea2"),
    ("ntsc", "-o", 0, "cbar.bmp", "out.bmp", "aabe954f22b3655e6e876282cbd04590"),
    ("ntsc", "-o", 24, "cbar.bmp", "out.bmp", "7b0463afc6c41cf120e6839755ed970a"),
    ("vhs", "-o", 24, "cbar.bmp", "out.bmp", "ad2f1e4bc7d256c6855345f73ad66bc7"),
    ("vhs", "-om", 24, "cbar.bmp", "out.bmp", "7a4f94b5e45cd8ae02b83c285159aa12"),
]


def make_inputs(where):
    Image = pytest.importorskip("PIL.Image")
    Image.open(os.path.join(INPUTS, "601cb.png")).convert("RGB").save(os.path.join(where, "601cb.ppm"))
    Image.open(os.path.join(INPUTS, "cbar.png")).convert("RGB").save(os.path.join(where, "cbar.bmp"))


def run_cli(kind, system, flags, noise, src, out, where):
    exe = os.path.join(S.REF_DIR, "cli_%s_%s" % (kind, system))
    if not os.path.exists(exe):
        pytest.skip("%s not built" % exe)
    subprocess.run([exe, flags, "832", "624", str(noise), "0", src, out], cwd=where, check=True, stdout=subprocess.DEVNULL, timeout=600)
    return hashlib.md5(open(os.path.join(where, out), "rb").read()).hexdigest()


@pytest.mark.parametrize("system,flags,noise,src,out,md5", ANCHORS)
def test_compiled_reference_cli_hits_the_anchor(tmp_path, system, flags, noise, src, out, md5):
    make_inputs(str(tmp_path))
    assert run_cli("ref", system, flags, noise, src, out, str(tmp_path)) == md5


@pytest.mark.gpu
@pytest.mark.parametrize("system,flags,noise,src,out,md5", ANCHORS)
def test_unmodified_cli_on_the_cuda_library_hits_the_anchor(tmp_path, system, flags, noise, src, out, md5):
    make_inputs(str(tmp_path))
    assert run_cli("b200", system, flags, noise, src, out, str(tmp_path)) == md5


def test_cli_601cb_progressive_md5():
    """the oracle replaying the CLI loop (crt_main.c:221-255) and the PPM writer (ppm_rw.c:96-121)"""
    Image = pytest.importorskip("PIL.Image")
    rgb = np.array(Image.open(os.path.join(INPUTS, "601cb.png")).convert("RGB"))
    img = np.zeros(rgb.shape[:2] + (4,), dtype=np.uint8)  # ppm_read24: 0x00RRGGBB ints
    img[..., 0], img[..., 1], img[..., 2] = rgb[..., 2], rgb[..., 1], rgb[..., 0]
    eng = S.OracleEngine("ntsc", 832, 624)
    S.cli_sequence(eng, img, noise=0, progressive=True, format=layout.PIX_BGRA)
    body = eng.out[..., [2, 1, 0]].tobytes()
    blob = b"P6\n832 624\n255\n" + body
    assert hashlib.md5(blob).hexdigest() == "e7bb5f48656e848f61036dc935802ea2"
