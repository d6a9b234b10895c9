// crt_kernels.cuh -- the sm_100a kernels of the composite modulate -> noise -> demodulate path.
//
// Kernel map (reference lines each one replaces):
//   k_mod_skeleton_rgb  sync / blanking / burst skeleton of all 262 lines  crt_ntsc.c:205-252, 325-329
//   k_mod_picture_rgb   RGB->YIQ, 3 band-limit IIRs, carrier mix -> analog crt_ntsc.c:254-324
//   k_mod_nes           PPU pixels -> square-wave IRE sums                 crt_nes.c:21-61, 81-201
//   k_noise             analog + LCG noise -> inp                          crt_core.c:346-367
//   k_sync              vsync search, hsync / burst-lock chain, line table crt_core.c:379-479
//   k_lines             Y/I/Q equalisers, resample, YIQ->RGB, blend, store crt_core.c:511-664
//
// Everything is int32 fixed point with two's-complement wrap and arithmetic right shift, as the
// compiled reference behaves; wrap-sensitive operations go through the
// w*() helpers so the compiler cannot exploit signed-overflow UB.
//
// Parallel shape: the recurrences along a scanline (eqf, iirf) round at every step and cannot be
// scanned, so one LANE carries one scanline (32 lines per warp); anything that is parallel along
// the line (pixel fetch, RGB->YIQ, row stores) runs lane-per-sample on tiles transposed through
// shared memory so that every global access is coalesced.  Scanline windows are staged into
// shared memory with 1-D TMA bulk copies (cp.async.bulk + mbarrier), double buffered.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "crt_sys.cuh"
#include "crtx_batch.h"
#include "crt_records.h"

namespace crt {


__constant__ int c_quarter15[18] = CRT_QUARTER15;

// ---------------------------------------------------------------------------------------
// arithmetic helpers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ int wmul(int a, int b) { return (int) ((unsigned) a * (unsigned) b); }
__device__ __forceinline__ int wadd(int a, int b) { return (int) ((unsigned) a + (unsigned) b); }
__device__ __forceinline__ int wsub(int a, int b) { return (int) ((unsigned) a - (unsigned) b); }
__device__ __forceinline__ int posmod(int x, int n) { return ((x % n) + n) % n; }
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

__device__ __forceinline__ int quarter_d(int a)
{
    int k = (a >> 8) & 0xff, fr = a & 0xff;
    int lo = c_quarter15[k], hi = c_quarter15[k + 1];
    return lo + (((hi - lo) * fr) >> 8);
}

__device__ __forceinline__ void sincos14_d(int &s, int &c, int n) // crt_core.c:42-61
{
    n &= 16383;
    int h = n & 8191;
    if (h >= 4096) {
        c = -quarter_d(h - 4096);
        s = quarter_d(8192 - h);
    } else {
        c = quarter_d(4096 - h);
        s = quarter_d(h);
    }
    if (n >= 8192) {
        c = -c;
        s = -s;
    }
}

__device__ __forceinline__ int warp_scan_incl(int v, int lane)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// byte positions of R, G, B in a pixel of each CRT_PIX_FORMAT (crt_core.h:62-67)
__device__ __forceinline__ void fmt_positions(int f, int &r, int &g, int &b)
{
    switch (f) {
        case CRT_PIX_FORMAT_RGB:  r = 0; g = 1; b = 2; break;
        case CRT_PIX_FORMAT_BGR:  r = 2; g = 1; b = 0; break;
        case CRT_PIX_FORMAT_ARGB: r = 1; g = 2; b = 3; break;
        case CRT_PIX_FORMAT_RGBA: r = 0; g = 1; b = 2; break;
        case CRT_PIX_FORMAT_ABGR: r = 3; g = 2; b = 1; break;
        default:                  r = 2; g = 1; b = 0; break; // BGRA
    }
}

// ---------------------------------------------------------------------------------------
// TMA (1-D bulk copy) + mbarrier primitives
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// global -> shared bulk copy; dst, src and bytes must all be multiples of 16
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, unsigned bytes, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// =======================================================================================
// encoder, RGB systems
// =======================================================================================
#if (CRT_SYSTEM != CRT_SYSTEM_NES)

// level of sample t of line n in the sync / blanking / burst skeleton (crt_ntsc.c:205-252)
__device__ __forceinline__ int skeleton_level(int n, int t, int field, int flip, int aberration, const int *burst)
{
    constexpr int H = kHres;
    if (n <= 3 || (n >= 7 && n <= 9)) {
        bool sync = (t < 4 * H / 100) || (t >= 50 * H / 100 && t < 54 * H / 100);
        return sync ? kSync : kBlank;
    }
    if (n >= 4 && n <= 6) {
        int first = (field == 1 ? 4 : 46) * H / 100;
        bool sync = (t < first) || (t >= 50 * H / 100 && t < 96 * H / 100);
        return sync ? kSync : kBlank;
    }
    if (t >= kCbBeg && t < kCbBeg + kBurstLen)
        return (int) (signed char) ((kBlank + burst[(t + flip * 2) & 3] * kBurst) >> 5);
    if (t >= kSyncBeg && t < kBwBeg && n < kVres - aberration) return kSync; // crt_ntscvhs.c:234-238
    return kBlank;
}

// One CTA per monitor.  Lines above CRT_TOP are written whole, active lines only up to AV_BEG
// (the rest of an active line belongs to the picture pass or keeps its old content).
__global__ void __launch_bounds__(256) k_mod_skeleton_rgb(const SrcCfg *__restrict__ srcs,
                                                          MonState *__restrict__ states,
                                                          signed char *__restrict__ analog_base, int first)
{
    static_assert(kHres % 2 == 0 && kAvBeg % 2 == 0, "pairs");
    constexpr int kFullPairs = kHres / 2, kHeadPairs = kAvBeg / 2;
    constexpr int kTotal = kTop * kFullPairs + (kVres - kTop) * kHeadPairs;
    const SrcCfg s = srcs[blockIdx.x];
    signed char *analog = analog_base + (size_t) (first + blockIdx.x) * kSignalBytes;
    __shared__ int burst[4];

    if (bpp_of(s.format) == 0) return; // crt_ntsc.c:190-193
    if (threadIdx.x < 4) {
        int v = 0;
        if (s.as_color) { // crt_ntsc.c:174-188
            int sn, cs;
            sincos14_d(sn, cs, (s.hue + (int) threadIdx.x * 90 + 33) * 8192 / 180);
            v = sn >> 10;
        }
        burst[threadIdx.x] = v;
    }
    __syncthreads();
    const int field = s.field & 1, frame = s.frame & 1;
    const int flip = (field == frame);

    for (int idx = threadIdx.x; idx < kTotal; idx += blockDim.x) {
        int n, t;
        if (idx < kTop * kFullPairs) {
            n = idx / kFullPairs;
            t = 2 * (idx - n * kFullPairs);
        } else {
            int r = idx - kTop * kFullPairs;
            n = kTop + r / kHeadPairs;
            t = 2 * (r % kHeadPairs);
        }
        char2 v;
        v.x = (signed char) skeleton_level(n, t, field, flip, s.aberration, burst);
        v.y = (signed char) skeleton_level(n, t + 1, field, flip, s.aberration, burst);
        *reinterpret_cast<char2 *>(analog + n * kHres + t) = v;
    }
    if (threadIdx.x < 4) { // prime the burst lock (crt_ntsc.c:325-329 / crt_ntscvhs.c:332-336)
        MonState *st = &states[first + blockIdx.x];
        int p = (int) (signed char) ((kBlank + burst[(threadIdx.x + flip * 2) & 3] * kBurst) >> 5);
        st->ccf[0][threadIdx.x] = kIsVhs ? 0 : p * 128;
        if (kIsVhs && threadIdx.x == 0) st->hsync = 0; // crt_ntscvhs.c:258-259
    }
}

// Picture pass.  One CTA (8 warps) per monitor; a warp owns 32 consecutive picture lines.
// Per 64-sample chunk: (A) lane-per-sample: fetch source pixels coalesced, RGB->YIQ, pack into a
// [line][sample] tile; (B) lane-per-line: the three serial band-limit IIRs, carrier mix, clamp ->
// bytes; (C) lane-per-sample: coalesced 2-byte stores into analog[].
constexpr int kModChunk = 64;
constexpr int kModTilePitch = kModChunk + 1; // words
constexpr int kModOutPitch = kModChunk / 4 + 1; // words
constexpr int kModWarpSmem = (32 * kModTilePitch + 32 * kModOutPitch) * 4;
constexpr int kModSmem = 8 * kModWarpSmem;

__device__ __forceinline__ void load_rgb(const unsigned char *data, size_t pix, int bpp, int rp, int gp, int bp,
                                         bool aligned4, int &r, int &g, int &b)
{
    if (bpp == 4 && aligned4) {
        unsigned v = __ldg(reinterpret_cast<const unsigned *>(data) + pix);
        r = (v >> (8 * rp)) & 0xff;
        g = (v >> (8 * gp)) & 0xff;
        b = (v >> (8 * bp)) & 0xff;
    } else {
        const unsigned char *p = data + pix * bpp;
        r = __ldg(p + rp);
        g = __ldg(p + gp);
        b = __ldg(p + bp);
    }
}

__global__ void __launch_bounds__(256) k_mod_picture_rgb(const SrcCfg *__restrict__ srcs,
                                                         const MonCfg *__restrict__ cfgs,
                                                         signed char *__restrict__ analog_base, int first)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned *tile = reinterpret_cast<unsigned *>(smem_raw + warp * kModWarpSmem);
    unsigned *obuf = tile + 32 * kModTilePitch;

    const SrcCfg s = srcs[blockIdx.x];
    const MonCfg cfg = cfgs[first + blockIdx.x];
    signed char *analog = analog_base + (size_t) (first + blockIdx.x) * kSignalBytes;

    const int bpp = bpp_of(s.format);
    if (bpp == 0) return;
    int destw = kAvLen, desth = (kLines * 64500) >> 16;
    if (s.raw) { // crt_ntsc.c:163-172
        destw = min(s.w, kAvLen);
        desth = min(s.h, desth);
    }
    if (destw <= 0 || desth <= 0 || s.w <= 0 || s.h <= 0) return;
    const int field = s.field & 1, frame = s.frame & 1;
    const int flip = (field == frame);
    const int ph = flip ? -1 : 1;
    const int xo = (kAvBeg + s.xoffset + (kAvLen - destw) / 2) & ~3;
    const int yo = kTop + s.yoffset + (kLines - desth) / 2;
    const int white = kWhite * cfg.white_point / 100;
    const int ire0 = kBlack + cfg.black_point;
    int rp, gp, bp;
    fmt_positions(s.format, rp, gp, bp);
    const unsigned char *data = static_cast<const unsigned char *>(s.data);
    const bool aligned4 = ((reinterpret_cast<uintptr_t>(data) & 3) == 0);

    int mI[4], mQ[4]; // ph * ccmodI/Q (crt_ntsc.c:174-188, 314-315)
#pragma unroll
    for (int k = 0; k < 4; k++) {
        mI[k] = mQ[k] = 0;
        if (s.as_color) {
            int sn, cs, deg = s.hue + k * 90;
            sincos14_d(sn, cs, deg * 8192 / 180);
            mI[k] = ph * (sn >> 10);
            sincos14_d(sn, cs, (deg - 90) * 8192 / 180);
            mQ[k] = ph * (sn >> 10);
        }
    }

    const int y = warp * 32 + lane; // this lane's picture line
    const int y0 = warp * 32;
    if (y0 >= desth) return;
    const int nlines = min(32, desth - y0);
    // source row of this lane's line (crt_ntsc.c:258-266).  The reference lets row == h read one
    // row past the image (undefined); we clamp to the last row instead.
    int row = (int) (((long long) min(y, desth - 1) * s.h) / desth) + (field * s.h + desth) / desth / 2;
    if (row >= s.h) row = s.h - 1;
    const int rowoff = row * s.w;

    int hy = 0, hi = 0, hq = 0;
    for (int c0 = 0; c0 < destw; c0 += kModChunk) {
        const int nx = min(kModChunk, destw - c0);
        // (A) fetch + RGB->YIQ, two samples per lane per line
        const int xa = c0 + lane, xb = c0 + lane + 32;
        const int cola = (int) (((long long) min(xa, destw - 1) * s.w) / destw);
        const int colb = (int) (((long long) min(xb, destw - 1) * s.w) / destw);
        for (int l = 0; l < nlines; l++) {
            const int ro = __shfl_sync(0xffffffffu, rowoff, l);
#pragma unroll
            for (int half = 0; half < 2; half++) {
                int r, g, b;
                load_rgb(data, (size_t) ro + (half ? colb : cola), bpp, rp, gp, bp, aligned4, r, g, b);
                int fy = (19595 * r + 38470 * g + 7471 * b) >> 14; // crt_ntsc.c:308-310
                int fi = (39059 * r - 18022 * g - 21103 * b) >> 14;
                int fq = (13894 * r - 34275 * g + 20382 * b) >> 14;
                tile[l * kModTilePitch + lane + 32 * half] =
                    (unsigned) fy | (((unsigned) fi & 0x7ffu) << 10) | ((unsigned) fq << 21);
            }
        }
        __syncwarp();
        // (B) serial along the line, one line per lane
        for (int x4 = 0; x4 < nx; x4 += 4) {
            unsigned packed = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                unsigned w = tile[lane * kModTilePitch + min(x4 + k, kModChunk - 1)];
                int fy = (int) (w & 0x3ffu);
                int fi = ((int) (w << 11)) >> 21;
                int fq = ((int) w) >> 21;
                hy += wmul(fy - hy, kIirY) >> 11; // iirf, crt_ntsc.c:117-126
                hi += wmul(fi - hi, kIirI) >> 11;
                hq += wmul(fq - hq, kIirQ) >> 11;
                int ci = wmul(hi, mI[k]) >> 4; // (x + xo) & 3 == k: xo and c0 + x4 are multiples of 4
                int cq = wmul(hq, mQ[k]) >> 4;
                int ire = ire0 + (wmul(hy + ci + cq, white) >> 10);
                ire = clampi(ire, 0, 110);
                packed |= (unsigned) ire << (8 * k);
            }
            obuf[lane * kModOutPitch + (x4 >> 2)] = packed;
        }
        __syncwarp();
        // (C) coalesced stores, 2 bytes per lane per line
        for (int l = 0; l < nlines; l++) {
            signed char *dst = analog + (c0 + xo) + (y0 + l + yo) * kHres;
            unsigned w = obuf[l * kModOutPitch + (lane >> 1)];
            unsigned two = (w >> (16 * (lane & 1))) & 0xffffu;
            if (2 * lane + 1 < nx) {
                *reinterpret_cast<unsigned short *>(dst + 2 * lane) = (unsigned short) two;
            } else if (2 * lane < nx) {
                dst[2 * lane] = (signed char) (two & 0xff);
            }
        }
        __syncwarp();
    }
}

#endif // RGB systems

// =======================================================================================
// encoder, NES (fully parallel: no recurrence along the line)
// =======================================================================================
#if (CRT_SYSTEM == CRT_SYSTEM_NES)

__constant__ int c_nes_level[16] = { -12042, 0,     34406,  81427,  -17203, -8028, 19497, 57342,
                                     43581,  75693, 112965, 112965, 26951,  52181, 83721, 83721 };
__constant__ int c_nes_emph[6] = { 0300, 0100, 0500, 0400, 0600, 0200 };

__device__ __forceinline__ int nes_square(int p, int phase) // crt_nes.c:21-61
{
    int hue = p & 15;
    if (hue >= 14) return 0;
    int emph = ((p & 0700) & c_nes_emph[(phase >> 1) % 6]) > 0;
    int high = (hue == 0) ? 1 : ((hue == 13) ? 0 : (((hue + phase) % 12) < 6));
    return c_nes_level[high * 8 + emph * 4 + ((p >> 4) & 3)];
}

// grid (ceil(H / 256), 262, n): one thread per sample of the field; each byte is written at most
// once with its final value (skeleton < burst < picture precedence of crt_nes.c:81-201).
__global__ void __launch_bounds__(256) k_mod_nes(const SrcCfg *__restrict__ srcs, const MonCfg *__restrict__ cfgs,
                                                 MonState *__restrict__ states,
                                                 signed char *__restrict__ analog_base, int first)
{
    const int t = blockIdx.x * 256 + threadIdx.x, n = blockIdx.y, m = blockIdx.z;
    const SrcCfg s = srcs[m];
    const MonCfg cfg = cfgs[first + m];
    signed char *analog = analog_base + (size_t) (first + m) * kSignalBytes;
    const int xo = (kAvBeg + s.xoffset) & ~3, yo = kTop + s.yoffset;

    if (n == 0 && blockIdx.x == 0 && threadIdx.x < 12) { // prime the burst lock (crt_nes.c:196-200)
        int row = threadIdx.x >> 2, x = threadIdx.x & 3, sn, cs;
        int deg = (s.hue + x * 90 + (row + s.dot_crawl_offset) * 120 + 33) % 360;
        sincos14_d(sn, cs, deg * 8192 / 180);
        // iccf[n % 3][t & 3] holds the last burst sample written for that (row, phase)
        states[first + m].ccf[row][x] = ((int) (signed char) ((kBlank + (sn >> 10) * kBurst) >> 5)) * 128;
    }
    if (t >= kHres) return;
    const int y = n - yo;
    const bool pic_line = (y >= 0 && y < kLines);
    int v = 0;
    bool write = false;
    if (s.reinit) { // setup_field, crt_nes.c:81-104
        int sync_end = (n >= 259) ? kNesVsyncEnd : kBwBeg;
        v = (t >= kSyncBeg && t < sync_end) ? kSync : kBlank;
        write = true;
    }
    if (pic_line && t >= kCbBeg && t < kCbBeg + kBurstLen) { // crt_nes.c:174-178
        int sn, cs;
        int deg = (s.hue + (t & 3) * 90 + ((n % 3) + s.dot_crawl_offset) * 120 + 33) % 360;
        sincos14_d(sn, cs, deg * 8192 / 180);
        v = (int) (signed char) ((kBlank + (sn >> 10) * kBurst) >> 5);
        write = true;
    }
    if (pic_line && t >= xo && t < xo + kAvLen) { // crt_nes.c:180-193
        const int x = t - xo;
        int row = (y * s.h) / kLines;
        if (row >= s.h) row = s.h - 1; // reference reads one row past the image here (undefined)
        if (row < 0) row = 0;
        const unsigned short *data = static_cast<const unsigned short *>(s.data);
        int p = __ldg(data + ((x * s.w) / kAvLen) + row * s.w);
        int phase = ((y + yo + s.dot_crawl_offset) % 3) * 4 + 3 * x;
        int ire = kBlack + cfg.black_point;
        ire += nes_square(p, phase) + nes_square(p, phase + 1) + nes_square(p, phase + 2)
             + nes_square(p, phase + 3);
        v = (int) (signed char) ((wmul(ire, cfg.white_point) / 100) >> 12);
        write = true;
    }
    if (write) analog[n * kHres + t] = (signed char) v;
}

#endif // NES

// =======================================================================================
// noise pass (crt_core.c:346-367), LCG variant.  16 samples per thread, 128-bit accesses;
// the generator state of sample i is rn0 advanced i + 1 steps, reached by two table look-ups.
// =======================================================================================
constexpr int kNoiseVec = 16;
constexpr int kNoiseThreads = (kInputSize + kNoiseVec - 1) / kNoiseVec;
constexpr int kNoiseBlocks = (kNoiseThreads + 255) / 256;
constexpr int kJumpLo = 128; // lo[k] = jump(16 * k), hi[k] = jump(16 * 128 * k)
constexpr int kJumpHi = (kNoiseThreads + kJumpLo - 1) / kJumpLo + 1;

__global__ void __launch_bounds__(256) k_noise(const MonCfg *__restrict__ cfgs, const MonState *__restrict__ states,
                                               const signed char *__restrict__ analog_base,
                                               signed char *__restrict__ inp_base,
                                               const Affine *__restrict__ jump_lo,
                                               const Affine *__restrict__ jump_hi, int first)
{
    const int m = first + blockIdx.y;
    const int t = blockIdx.x * 256 + threadIdx.x;
    const int i0 = t * kNoiseVec;
    if (i0 >= kInputSize) return;
    if (cfgs[m].bpp == 0) return; // crt_core.c:312-315: nothing at all happens
    const int noise = cfgs[m].noise;
    const signed char *analog = analog_base + (size_t) m * kSignalBytes;
    signed char *inp = inp_base + (size_t) m * kSignalBytes;

    uint4 in = *reinterpret_cast<const uint4 *>(analog + i0);
    unsigned w[4] = { in.x, in.y, in.z, in.w };
    if (noise == 0) {
#pragma unroll
        for (int k = 0; k < 4; k++) w[k] = __vmaxs4(w[k], 0x81818181u); // clamp -128 -> -127
    } else {
        const Affine lo = jump_lo[t % kJumpLo], hi = jump_hi[t / kJumpLo];
        unsigned rn = ((unsigned) states[m].rn * hi.mul + hi.add) * lo.mul + lo.add;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            unsigned o = 0;
#pragma unroll
            for (int b = 0; b < 4; b++) {
                rn = rn * kLcgMul + kLcgAdd;
                int a = (int) (signed char) (w[k] >> (8 * b));
                int s = a + (wmul((int) ((rn >> 16) & 0xff) - 0x7f, noise) >> 8);
                s = clampi(s, -127, 127);
                o |= ((unsigned) s & 0xffu) << (8 * b);
            }
            w[k] = o;
        }
    }
    if (i0 + kNoiseVec <= kInputSize) {
        *reinterpret_cast<uint4 *>(inp + i0) = make_uint4(w[0], w[1], w[2], w[3]);
    } else {
        for (int b = 0; i0 + b < kInputSize; b++) inp[i0 + b] = (signed char) (w[b >> 2] >> (8 * (b & 3)));
    }
}

#if (CRT_SYSTEM == CRT_SYSTEM_NTSCVHS)
// VHS noise pass with the per-sample noise term drawn elsewhere (crt_core.c:343-366): the drop-in
// path draws from the process's libc rand() on the host, exactly as the reference does, and hands
// the resulting terms over; this kernel only adds and clamps.
__global__ void __launch_bounds__(256) k_noise_terms(const MonCfg *__restrict__ cfgs,
                                                     const signed char *__restrict__ analog_base,
                                                     signed char *__restrict__ inp_base,
                                                     const short *__restrict__ terms, int first)
{
    const int m = first + blockIdx.y;
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= kInputSize || cfgs[m].bpp == 0) return;
    const signed char *analog = analog_base + (size_t) m * kSignalBytes;
    signed char *inp = inp_base + (size_t) m * kSignalBytes;
    int s = analog[i] + terms[(size_t) blockIdx.y * kInputSize + i];
    inp[i] = (signed char) clampi(s, -127, 127);
}
#endif

// =======================================================================================
// sync pre-pass: one warp per monitor (crt_core.c:379-479).
// vsync search in parallel over a line (segment sums + warp scan), then the serial line-to-line
// chain: hsync search (16 samples: one load per lane + scan + ballot) and the colour-burst lock
// (4 lanes, one per carrier phase, 10 truncating steps each).  Emits one LineRec per decoded line.
// =======================================================================================
__global__ void __launch_bounds__(32) k_sync(const MonCfg *__restrict__ cfgs, MonState *__restrict__ states,
                                             LineRec *__restrict__ lines_base,
                                             const signed char *__restrict__ inp_base, int first,
                                             int force_generic)
{
    const int m = first + blockIdx.x, lane = threadIdx.x;
    const MonCfg cfg = cfgs[m];
    if (cfg.bpp == 0) return;
    MonState *st = &states[m];
    const signed char *inp = inp_base + (size_t) m * kSignalBytes;
    LineRec *lines = lines_base + (size_t) m * kLines;

    int huesn, huecs;
    {
        int sn, cs;
        sincos14_d(sn, cs, ((cfg.hue % 360) + 33) * 8192 / 180); // crt_core.c:318-320
        huesn = sn >> 11;
        huecs = cs >> 11;
    }
    int vs = st->vsync, hs = st->hsync;

    // ---- vsync (crt_core.c:379-396)
    constexpr int kSeg = (kHres + 31) / 32;
    int line = 0, j = kHres;
    bool found = false;
    for (int i = -kVsyncWindow; i < kVsyncWindow && !found; i++) {
        line = posmod(vs + i, kVres);
        const signed char *sig = inp + line * kHres;
        const int b0 = lane * kSeg, b1 = min(kHres, b0 + kSeg);
        int sum = 0;
        for (int t = b0; t < b1; t++) sum += __ldg(sig + t);
        int acc = warp_scan_incl(sum, lane) - sum, idx = -1;
        for (int t = b0; t < b1; t++) {
            acc += __ldg(sig + t);
            if (idx < 0 && acc <= kVsyncLevel) idx = t;
        }
        unsigned hit = __ballot_sync(0xffffffffu, idx >= 0);
        if (hit) {
            j = __shfl_sync(0xffffffffu, idx, __ffs(hit) - 1);
            found = true;
        }
    }
    vs = line;
    int field = (j > kHres / 2);
    const int ratio = (((cfg.outh << 16) / kLines) + 32768) >> 16; // crt_core.c:404-407
    field *= ratio / 2;

    // ---- line chain.  Lane p < 4 carries ccf[row][p] for the three possible rows.
    int c0 = st->ccf[0][lane & 3], c1 = st->ccf[1 % kVper][lane & 3], c2 = st->ccf[2 % kVper][lane & 3];
    // The fast equaliser path of k_lines is exact while |wave| <= 65536 and |bright| <= 4096
    // (see eq_step in crt_lines.cuh); anything else sends the whole monitor down the generic one.
    int generic = force_generic || abs(cfg.brightness - (kBlack + cfg.black_point)) > 4096;
    for (int k = 0; k < kLines; k++) {
        LineRec rec;
        rec.pad0 = rec.pad1 = 0;
        int beg = (int) ((unsigned) k * ((unsigned) cfg.outh + cfg.v_fac) / (unsigned) kLines + (unsigned) field);
        int end = (int) ((unsigned) (k + 1) * ((unsigned) cfg.outh + cfg.v_fac) / (unsigned) kLines + (unsigned) field);
        if (beg >= cfg.outh) { // crt_core.c:431: no state is touched
            if (lane == 0) {
                rec.pos = 0; rec.wave0 = rec.wave1 = 0; rec.beg = -1; rec.end = -1; rec.hsync = hs;
                lines[k] = rec;
            }
            continue;
        }
        if (end > cfg.outh) end = cfg.outh;

        const int ln = posmod(kTop + k + vs, kVres) * kHres;
        { // hsync (crt_core.c:437-450)
            int v = 0;
            if (lane < 2 * kHsyncWindow) v = __ldg(inp + ln + hs + kSyncBeg - kHsyncWindow + lane);
            int acc = warp_scan_incl(v, lane);
            unsigned hit = __ballot_sync(0xffffffffu, lane < 2 * kHsyncWindow && acc <= kHsyncLevel);
            int i = (hit ? __ffs(hit) - 1 : 2 * kHsyncWindow) - kHsyncWindow;
            hs = posmod(i + hs, kHres);
        }
        const int xpos = posmod(kAvBeg + hs - 3, kHres);
        const int ypos = posmod(kTop + k + vs + 3, kVres);
        const int row = ypos % kVper;

        // burst lock (crt_core.c:456-467): ccr[i & 3] = ccr[i & 3] * 127 / 128 + sig[i]
        int x = (row == 0) ? c0 : ((row == 1) ? c1 : c2);
        if (lane < 4) {
            const signed char *bs = inp + ln + (hs & ~3) + kCbBeg;
            const int t0 = (lane - kCbBeg) & 3;
            int smp[kBurstLen / 4];
#pragma unroll
            for (int q = 0; q < kBurstLen / 4; q++) smp[q] = __ldg(bs + t0 + 4 * q);
#pragma unroll
            for (int q = 0; q < kBurstLen / 4; q++) x = wadd(wmul(x, 127) / 128, smp[q]);
        }
        if (row == 0) c0 = x; else if (row == 1) c1 = x; else c2 = x;

        const int pa = hs & 3; // crt_core.c:469-479
        const int a0 = __shfl_sync(0xffffffffu, x, pa), a1 = __shfl_sync(0xffffffffu, x, (pa + 1) & 3);
        const int a2 = __shfl_sync(0xffffffffu, x, (pa + 2) & 3), a3 = __shfl_sync(0xffffffffu, x, (pa + 3) & 3);
        const int dci = wsub(a1, a3), dcq = wsub(a2, a0);
        if (lane == 0) {
            rec.pos = xpos + ypos * kHres;
            rec.wave0 = wmul(wsub(wmul(dci, huecs), wmul(dcq, huesn)) >> 4, cfg.saturation);
            rec.wave1 = wmul(wadd(wmul(dcq, huecs), wmul(dci, huesn)) >> 4, cfg.saturation);
            rec.beg = beg;
            rec.end = end;
            rec.hsync = hs;
            lines[k] = rec;
            if (abs(rec.wave0) > 65536 || abs(rec.wave1) > 65536) generic = 1;
        }
    }
    if (lane < 4) {
        st->ccf[0][lane] = c0;
        if (kVper > 1) st->ccf[1 % kVper][lane] = c1;
        if (kVper > 2) st->ccf[2 % kVper][lane] = c2;
    }
    if (lane == 0) {
        st->vsync = vs;
        st->hsync = hs;
        st->field = field;
        st->generic = generic;
        if (!kIsVhs) st->rn = (int) ((unsigned) st->rn * kLcgField.mul + kLcgField.add); // crt_core.c:367
    }
}

} // namespace crt

#include "crt_lines.cuh"
