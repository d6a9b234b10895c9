"""BASELINE.md section 4 parity anchor: md5 of the reference CLI's output file.

`ntsc -op 832 624 0 0 601cb.ppm out.ppm` -> e7bb5f48656e848f61036dc935802ea2 (recorded from the
reference CLI during the survey, re-verified when the oracle was written).  The input image
lives in the reference's own zip, so this only runs where /root/reference is mounted; the
oracle replays the CLI loop (crt_main.c:221-255) and the PPM writer (ppm_rw.c:96-121).
"""
import hashlib
import io
import os
import zipfile

import numpy as np
import pytest

import support as S
from ntsc_crt_b200 import layout

ZIP = "/root/reference/extra/test_output_images.zip"


@pytest.mark.skipif(not os.path.exists(ZIP), reason="reference assets not mounted")
def test_cli_601cb_progressive_md5():
    Image = pytest.importorskip("PIL.Image")
    z = zipfile.ZipFile(ZIP)
    rgb = np.array(Image.open(io.BytesIO(z.read("test_output_images/original/601cb.png"))).convert("RGB"))
    img = np.zeros(rgb.shape[:2] + (4,), dtype=np.uint8)  # ppm_read24: 0x00RRGGBB ints
    img[..., 0], img[..., 1], img[..., 2] = rgb[..., 2], rgb[..., 1], rgb[..., 0]
    eng = S.OracleEngine("ntsc", 832, 624)
    S.cli_sequence(eng, img, noise=0, progressive=True, format=layout.PIX_BGRA)
    body = eng.out[..., [2, 1, 0]].tobytes()
    blob = b"P6\n832 624\n255\n" + body
    assert hashlib.md5(blob).hexdigest() == "e7bb5f48656e848f61036dc935802ea2"
