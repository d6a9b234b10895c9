"""k_lines2's ring schedule (csrc/crt_lines2.cuh), restated in Python and checked exhaustively on the CPU: for every
output width the host-side predicate `lines2_geometry_ok` admits, every block of 8 pixels is emitted while both
samples of each of its pixels are still in the 24-slot ring (plus the guard slot that repeats slot 0), and every
pixel of the line is emitted by the time the last filter sub-chunk is done.  This is the proof behind the bound
`7 * dx <= 10 * 4096` and does not need a GPU."""
import pytest

AV_LEN = {"ntsc": 753, "nes": 682, "nes_p0": 684}  # crt_ntsc.h / crt_nes.h AV_LEN
SUB, RING, BLOCK, MAX_OUTW = 12, 24, 8, 1312


def geometry_ok(outw, av_len):
    if outw < 16 or outw > MAX_OUTW or (outw & 3):
        return False
    dx = ((av_len - 1) << 12) // outw
    return 7 * dx <= 10 * 4096


@pytest.mark.parametrize("system", sorted(AV_LEN))
def test_every_admitted_width_keeps_its_samples_in_the_ring(system):
    av_len = AV_LEN[system]
    padded = (av_len + SUB - 1) // SUB * SUB
    admitted = 0
    for outw in range(16, MAX_OUTW + 1, 4):
        if not geometry_ok(outw, av_len):
            continue
        admitted += 1
        dx = ((av_len - 1) << 12) // outw
        nblk = (outw + BLOCK - 1) // BLOCK
        samp = [(min(k, outw - 1) * dx) >> 12 for k in range(nblk * BLOCK)]
        blk = 0
        ring = [None] * (RING + 1)
        for sub in range(padded // SUB):
            for t in range(SUB):  # the filter block writes sample sub * 12 + t
                s = sub * SUB + t
                ring[s % RING] = s
                if s % RING == 0:
                    ring[RING] = s  # guard slot
            have = sub * SUB + SUB - 1
            while blk < nblk and samp[blk * BLOCK + BLOCK - 1] + 1 <= have:
                for k in range(blk * BLOCK, blk * BLOCK + BLOCK):
                    a = samp[k] % RING
                    assert ring[a] == samp[k] and ring[a + 1] == samp[k] + 1, (system, outw, k, sub)
                blk += 1
        assert blk == nblk, (system, outw)
        assert samp[-1] + 1 < av_len  # the resampler stops before AV_LEN (crt_core.c:529, 555)
    assert admitted > 150


def test_the_widths_the_drivers_use_are_admitted():
    for outw in (640, 832, 1024, 1280):
        assert geometry_ok(outw, 753), outw
    assert not geometry_ok(256, 753) and not geometry_ok(1316, 753) and not geometry_ok(1920, 753) and not geometry_ok(830, 753)
