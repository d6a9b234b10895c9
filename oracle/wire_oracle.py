"""TEST INFRASTRUCTURE (oracle): numpy restatements of the reference's image wire formats and of the live driver's
phosphor decay -- byte shuffles around the hot path, checked against the device kernels crtx_bmp_* / crtx_ppm_* /
crtx_fade_phosphors.  Only tests/ imports this.

    bmp_unpack / bmp_pack   bmp_rw.c:22-94 / 96-146   (24- and 32-bit bottom-up rows <-> top-down BGRA ints)
    ppm_unpack / ppm_pack   ppm_rw.c:79-89 / 113-118  (P6 RGB bytes <-> int 0x00RRGGBB)
    fade_phosphors          crt_main.c:437-452
"""
import numpy as np


def bmp_unpack(file_pixels, w, h, bits):
    """bmp_rw.c:50-94: rows are stored bottom-up, padded to 4 bytes; 24-bit pixels get alpha 255 (line 90)."""
    bpp = bits // 8
    rowbytes = (w * bpp + 3) & ~3
    rows = np.frombuffer(file_pixels, dtype=np.uint8, count=rowbytes * h).reshape(h, rowbytes)[::-1, : w * bpp]
    px = rows.reshape(h, w, bpp).astype(np.uint32)
    out = px[..., 0] | px[..., 1] << 8 | px[..., 2] << 16
    out |= (px[..., 3] << 24) if bpp == 4 else np.uint32(0xFF000000)
    return np.ascontiguousarray(out, dtype=np.uint32)


def bmp_pack(bgra, w, h):
    """bmp_rw.c:138-143: 32-bit pixels, bottom-up rows."""
    return np.ascontiguousarray(np.asarray(bgra, dtype=np.uint32).reshape(h, w)[::-1])


def ppm_unpack(file_pixels, w, h, maxc=255):
    """ppm_rw.c:79-89: out[i] = TO_8_BIT(r) << 16 | TO_8_BIT(g) << 8 | TO_8_BIT(b), TO_8_BIT(x) = (x*255 + maxc/2) / maxc."""
    rgb = np.frombuffer(file_pixels, dtype=np.uint8, count=3 * w * h).reshape(h * w, 3).astype(np.uint32)
    rgb = (rgb * 255 + maxc // 2) // maxc
    return (rgb[:, 0] << 16 | rgb[:, 1] << 8 | rgb[:, 2]).astype(np.uint32).reshape(h, w)


def ppm_pack(xrgb, w, h):
    """ppm_rw.c:113-118: three fputc per pixel, red first."""
    c = np.asarray(xrgb, dtype=np.uint32).reshape(h * w)
    return np.stack([(c >> 16) & 0xFF, (c >> 8) & 0xFF, c & 0xFF], axis=1).astype(np.uint8).reshape(-1)


def fade_phosphors(image):
    """crt_main.c:446-450."""
    c = np.asarray(image, dtype=np.uint32) & 0xFFFFFF
    return ((c >> 1) & 0x7F7F7F) + ((c >> 2) & 0x3F3F3F) + ((c >> 3) & 0x1F1F1F) + ((c >> 4) & 0x0F0F0F)
