// crt_lines_fir.cuh -- k_lines_fir, the line pass of crt_demodulate for the reference's
// USE_CONVOLUTION 1 build (crt_core.c:85-147 with crt_core.c:511-664): eqf() is the 7-tap kernel
// [1 4 7 8 7 4 1] >> 5 over a history that is zero at the start of every line.
//
// Shape.  Unlike the three-band equaliser (crt_lines.cuh) this filter has no recurrence, so the line
// itself is data parallel: ONE WARP DECODES ONE SCANLINE.
//   in : the line's window of inp[] arrives in shared memory with one 1-D TMA bulk copy (784 bytes,
//        the 16-byte aligned superset of 768 samples at any byte phase);
//   F  : every lane filters 24 consecutive samples (plus a 6-sample run-in that rebuilds the filter
//        history), the kernel factored as [1 1]^3 * [1 1 1 1] -- five additions per channel and
//        sample, exact because nothing is rounded before the final shift -- and writes packed Y/I/Q
//        into the warp's shared-memory row;
//   P  : lane = output pixel, 32 consecutive pixels per step: resample, YIQ->RGB, contrast, clamp,
//        blend with the previous image (one coalesced 128-byte load) and one coalesced 128-byte store
//        per output row the line covers (crt_core.c:662-664).
#pragma once

#include "crt_lines.cuh"

namespace crt {

constexpr int kFirWarps = 8;                       // lines in flight per CTA
constexpr int kFirChunk = 24;                      // samples per lane: a multiple of 8 (slot padding) and 4 (carrier)
constexpr int kFirSamples = 32 * kFirChunk;        // 768 >= AV_LEN of every system
constexpr int kFirHalo = 6;                        // taps - 1
constexpr int kFirStage = ((kFirSamples + 15 + 15) / 16) * 16; // staged bytes per line
constexpr int kFirSeg = 1024;                      // output pixels staged per bulk load / store
constexpr int kFirGroups = (kLines + kFirWarps - 1) / kFirWarps;
static_assert(kFirSamples >= kAvLen, "one warp covers a whole line");
static_assert(kFirChunk % 8 == 0 && kFirChunk % 4 == 0, "slot padding and carrier phase are per-lane constants");

// Y/I/Q row of one line.  Sample e lives in slot 1 + e + (e >> 3): one pad slot after every 8 samples
// makes the per-lane chunk pitch 27 entries (54 words), which spreads "all lanes, same t" stores over
// all banks; the pad slot after samples 8j..8j+7 holds a COPY of sample 8j+8, so the resampler always
// finds sample s + 1 in the slot after sample s.
template <bool FAST> struct FirRow {
    static constexpr int kEntryBytes = FAST ? 8 : 16;
    static constexpr int kSlots = 1 + kFirSamples + kFirSamples / 8;
    static constexpr int kBytes = ((kSlots * kEntryBytes + 15) / 16) * 16;
};
template <bool FAST> constexpr int fir_warp_smem() { return kFirStage + FirRow<FAST>::kBytes + kFirSeg * 4; }
template <bool FAST> constexpr int fir_smem() { return kFirWarps * fir_warp_smem<FAST>() + kFirWarps * 2 * 8; }

// shared -> global bulk copy (the mirror of tma_load_1d) and its bookkeeping
__device__ __forceinline__ void tma_store_1d(void *dst, const void *src, unsigned bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// one channel of the factored kernel: [1 1] three times, then a 4-wide box as two pair sums
struct FirChan {
    int a, b, c, d, p1, p2;
};
__device__ __forceinline__ void fir_reset(FirChan &f) { f.a = f.b = f.c = f.d = f.p1 = f.p2 = 0; }
__device__ __forceinline__ int fir_push(FirChan &f, int x)
{
    const int s1 = wadd(x, f.a);
    f.a = x;
    const int s2 = wadd(s1, f.b);
    f.b = s1;
    const int s3 = wadd(s2, f.c);
    f.c = s2;
    const int p = wadd(s3, f.d); // s3[i] + s3[i-1]
    f.d = s3;
    const int out = wadd(p, f.p2); // + s3[i-2] + s3[i-3]
    f.p2 = f.p1;
    f.p1 = p;
    return out;
}

// Grid: (line groups, monitors).  blockIdx.x strides over the kFirGroups groups of kFirWarps lines, so the
// host can launch the common instantiation with one group per CTA and the rarely-needed one (see FAST)
// with a few CTAs per monitor that cost almost nothing when they find nothing to do.
template <bool FAST, int MODE, int FMT>
__global__ void __launch_bounds__(kFirWarps * 32)
k_lines_fir(const MonCfg *__restrict__ cfgs, const MonState *__restrict__ states, const LineRec *__restrict__ lines_base,
            const signed char *__restrict__ inp_base, int first, const LinesGeom geo)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int kEntry = FirRow<FAST>::kEntryBytes;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m = first + blockIdx.y;
    // everything a line needs from global memory, requested together
    const int is_generic = states[m].generic;
    const MonCfg *cfg = &cfgs[m];
    const int contrast = cfg->contrast;
    const int bright = cfg->brightness - (kBlack + cfg->black_point); // crt_core.c:304
    const int scanlines = cfg->scanlines;
    unsigned char *out = cfg->out;
    if ((is_generic != 0) == FAST) return; // the other instantiation handles this monitor
    if (geo.bpp == 0 || geo.outw <= 0) return;

    unsigned char *stage = smem_raw + warp * fir_warp_smem<FAST>();
    unsigned char *yiq = stage + kFirStage;
    unsigned *orow = reinterpret_cast<unsigned *>(yiq + FirRow<FAST>::kBytes);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + kFirWarps * fir_warp_smem<FAST>()) + 2 * warp;
    if (lane == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
    }
    __syncwarp();
    unsigned ph_sig = 0, ph_old = 0;

    constexpr int bpp = (MODE == 2) ? 3 : 4;
    const int pitch = geo.outw * bpp;
    // rows go through shared memory and bulk copies when they are 16-byte granular; otherwise (3-byte
    // pixels, odd widths, unaligned images) every lane reads and writes its own pixels
    const bool bulk = (MODE != 2) && geo.use_tma && ((geo.outw & 3) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    const signed char *inp = inp_base + (size_t) m * kSignalBytes;
    const unsigned dx = (unsigned) (((kAvLen - 1) << 12) / geo.outw); // crt_core.c:527
    constexpr unsigned sel_store = (FMT == CRT_PIX_FORMAT_RGBA) ? 0x4012u : (FMT == CRT_PIX_FORMAT_ARGB) ? 0x0124u
                                 : (FMT == CRT_PIX_FORMAT_ABGR) ? 0x2104u : 0x4210u;
    constexpr unsigned alpha_ff = (FMT == CRT_PIX_FORMAT_ARGB || FMT == CRT_PIX_FORMAT_ABGR) ? 0x000000ffu : 0xff000000u;
    constexpr unsigned blend_mask = 0x7f7f7f7fu & ~(alpha_ff >> 1) & ~alpha_ff;
    int rp = 0, gp = 0, bp = 0;
    if (MODE == 2) fmt_positions(geo.out_format, rp, gp, bp);

#pragma unroll 1
    for (int g = blockIdx.x; g < kFirGroups; g += gridDim.x) {
        const int kline = g * kFirWarps + warp;
        if (kline >= kLines) break;
        const LineRec rec = lines_base[(size_t) m * kLines + kline];
        const bool active = rec.beg >= 0 && (geo.pass == -1 || (geo.pass == -2 ? rec.pad1 != 0 : rec.pad0 == geo.pass));
        if (!active) continue; // warp-uniform

        const int a = rec.pos & 15;
        const signed char *src = inp + (rec.pos & ~15);
        const int nrows = max(1, rec.end - scanlines - rec.beg); // crt_core.c:662-664
        unsigned char *row0 = out + (size_t) rec.beg * pitch;
        const int seg0 = min(geo.outw, kFirSeg);
        if (geo.use_tma) {
            if (lane == 0) {
                mbar_expect_tx(&bars[0], kFirStage);
                tma_load_1d(stage, src, kFirStage, &bars[0]);
                if (bulk && MODE == 1) { // the previous image's pixels of the first segment, for the blend
                    mbar_expect_tx(&bars[1], seg0 * 4);
                    tma_load_1d(orow, row0, seg0 * 4, &bars[1]);
                }
            }
        } else {
            for (int q = lane; q < kFirStage / 16; q += 32)
                reinterpret_cast<uint4 *>(stage)[q] = __ldg(reinterpret_cast<const uint4 *>(src) + q);
        }
        const int nw0 = wsub(0, rec.wave0), nw1 = wsub(0, rec.wave1);
        // wave[(i + 0) & 3] feeds I, wave[(i + 3) & 3] feeds Q (crt_core.c:538-543)
        const int wi[4] = { rec.wave0, rec.wave1, nw0, nw1 };
        const int wq[4] = { nw1, rec.wave0, rec.wave1, nw0 };

        if (geo.use_tma) {
            mbar_wait(&bars[0], ph_sig);
            ph_sig ^= 1;
        } else {
            __syncwarp();
        }

        // ---- (F) samples [24 * lane - 6, 24 * lane + 24); the first six only rebuild the history
        const int e0 = lane * kFirChunk;
        if (e0 - kFirHalo < kAvLen) {
            const signed char *sg = reinterpret_cast<const signed char *>(stage) + a + e0;
            const bool head = (lane == 0); // samples before the line start are zeros, not signal (crt_core.c:534-536)
            FirChan fy, fi, fq;
            fir_reset(fy);
            fir_reset(fi);
            fir_reset(fq);
#pragma unroll
            for (int j = 0; j < kFirHalo; j++) {
                const int i4 = (j + 2) & 3; // (e0 - 6 + j) & 3, e0 a multiple of 4
                const int s = head ? 0 : (int) sg[head ? 0 : j - kFirHalo];
                (void) fir_push(fy, head ? 0 : wadd(s, bright));
                (void) fir_push(fi, wmul(s, wi[i4]) >> 9);
                (void) fir_push(fq, wmul(s, wq[i4]) >> 9);
            }
            unsigned char *dst = yiq + (size_t) (1 + lane * (kFirChunk + kFirChunk / 8)) * kEntry;
#pragma unroll
            for (int t = 0; t < kFirChunk; t++) {
                const int s = sg[t];
                const int y = fir_push(fy, wadd(s, bright)) >> 5;
                const int ci = fir_push(fi, wmul(s, wi[t & 3]) >> 9) >> 8; // (v >> 5) >> 3
                const int cq = fir_push(fq, wmul(s, wq[t & 3]) >> 9) >> 8;
                const int slot = t + (t >> 3);
                // FAST stores {Y, I | Q << 16} with Y not yet scaled by 16 (see crt_lines.cuh: the pixel pass
                // folds the scale into its weights); the generic path keeps {Y * 16, I, Q} verbatim.
                if (FAST) {
                    const uint2 v = make_uint2((unsigned) y, __byte_perm((unsigned) ci, (unsigned) cq, 0x5410));
                    *reinterpret_cast<uint2 *>(dst + slot * kEntry) = v;
                    if ((t & 7) == 0) *reinterpret_cast<uint2 *>(dst + (slot - 1) * kEntry) = v;
                } else {
                    const uint4 v = make_uint4((unsigned) wmul(y, 16), (unsigned) ci, (unsigned) cq, 0u);
                    *reinterpret_cast<uint4 *>(dst + slot * kEntry) = v;
                    if ((t & 7) == 0) *reinterpret_cast<uint4 *>(dst + (slot - 1) * kEntry) = v;
                }
            }
        }
        __syncwarp();

        // ---- (P) pixels (crt_core.c:555-659), 32 consecutive ones per step
        const unsigned char *slot1 = yiq + kEntry; // slot of sample 0
        auto pixel = [&](int px) -> unsigned { // 0x00RRGGBB of output pixel px
            const unsigned pos = (unsigned) px * dx;
            const unsigned s = pos >> 12;
            const unsigned char *sp = slot1 + (s + (s >> 3)) * kEntry;
            if (FAST) {
                const uint2 va = *reinterpret_cast<const uint2 *>(sp);
                const uint2 vb = *reinterpret_cast<const uint2 *>(sp + kEntry);
                const int R4 = (int) ((pos & 0xfffu) << 2), L4 = 0x3ffc - R4; // 4 * R, 4 * L
                const int ai = (int) (short) (unsigned short) va.y, aq = ((int) va.y) >> 16;
                const int bi = (int) (short) (unsigned short) vb.y, bq = ((int) vb.y) >> 16;
                const int y = wadd(wmul((int) va.x, L4), wmul((int) vb.x, R4));
                // (v * 4L) >> 16 == (v * L) >> 14: the two dropped bits are zeros
                return yiq_to_rgb(y, wadd(wmul(ai, L4) >> 16, wmul(bi, R4) >> 16),
                                  wadd(wmul(aq, L4) >> 16, wmul(bq, R4) >> 16), contrast);
            } else {
                const uint4 va = *reinterpret_cast<const uint4 *>(sp);
                const uint4 vb = *reinterpret_cast<const uint4 *>(sp + kEntry);
                const int R = (int) (pos & 0xfffu), L = 0xfff - R;
                return yiq_pixel((int) va.x, (int) va.y, (int) va.z, (int) vb.x, (int) vb.y, (int) vb.z, R, L, contrast);
            }
        };

        if (bulk) {
            // The segment's previous pixels are already in `orow` (blend) -- each lane turns its own words
            // into the new pixels in place, then one lane sends the finished segment to every row of the line.
#pragma unroll 1
            for (int k0 = 0; k0 < geo.outw; k0 += kFirSeg) {
                const int cnt = min(kFirSeg, geo.outw - k0);
                if (MODE == 1) {
                    if (k0 > 0) {
                        if (lane == 0) {
                            mbar_expect_tx(&bars[1], cnt * 4);
                            tma_load_1d(orow, row0 + (size_t) k0 * 4, cnt * 4, &bars[1]);
                        }
                    }
                    mbar_wait(&bars[1], ph_old);
                    ph_old ^= 1;
                }
#pragma unroll 2
                for (int j = lane; j < cnt; j += 32) {
                    const unsigned rgb = pixel(k0 + j);
                    unsigned v = (FMT == CRT_PIX_FORMAT_BGRA) ? (rgb | alpha_ff) : __byte_perm(rgb, 0xffu, sel_store);
                    if (MODE == 1) v = (((v >> 1) & blend_mask) | alpha_ff) + ((orow[j] >> 1) & blend_mask); // crt_core.c:608
                    orow[j] = v;
                }
                fence_async_smem();
                __syncwarp();
                if (lane == 0) {
                    for (int r = 0; r < nrows; r++) tma_store_1d(row0 + (size_t) r * pitch + (size_t) k0 * 4, orow, cnt * 4);
                    tma_store_commit();
                    tma_store_wait_read(); // `orow` is rewritten by the next segment / line
                }
                __syncwarp();
            }
        } else {
#pragma unroll 1
            for (int px = lane; px < geo.outw; px += 32) {
                unsigned rgb = pixel(px);
                if (MODE != 2) {
                    unsigned char *p = row0 + (size_t) px * 4;
                    unsigned v = (FMT == CRT_PIX_FORMAT_BGRA) ? (rgb | alpha_ff) : __byte_perm(rgb, 0xffu, sel_store);
                    if (MODE == 1)
                        v = (((v >> 1) & blend_mask) | alpha_ff) + ((__ldcg(reinterpret_cast<const unsigned *>(p)) >> 1) & blend_mask);
                    for (int r = 0; r < nrows; r++) __stcg(reinterpret_cast<unsigned *>(p + (size_t) r * pitch), v);
                } else {
                    unsigned char *p = row0 + (size_t) px * 3;
                    if (geo.blend) {
                        const unsigned o = (unsigned) p[rp] << 16 | (unsigned) p[gp] << 8 | (unsigned) p[bp];
                        rgb = ((rgb >> 1) & 0x7f7f7fu) + ((o >> 1) & 0x7f7f7fu);
                    }
                    for (int r = 0; r < nrows; r++) {
                        unsigned char *d = p + (size_t) r * pitch;
                        d[rp] = (unsigned char) (rgb >> 16);
                        d[gp] = (unsigned char) (rgb >> 8);
                        d[bp] = (unsigned char) rgb;
                    }
                }
            }
            __syncwarp(); // every lane is done with `yiq` before the next line's filter pass rewrites it
        }
    }
}

} // namespace crt
