// crt_sync.cuh -- k_sync, the sync pre-pass of crt_demodulate (crt_core.c:379-479).
//
// What is serial in the reference and stays serial here (the values are exact integers with
// truncation at every step, so there is no closed form):
//   * the hsync chain: line k searches a 2W-sample window positioned by line k-1's result;
//   * the colour-burst lock: ccr[p] = ccr[p] * 127 / 128 + sample, 10 steps per phase per line, carried
//     from line to line (and, with CRT_CC_VPER = 3, along three interleaved rows).
// What this kernel does about it: one CTA of 4 warps per monitor;
//   1. all warps copy the head of every signal line (the only bytes either chain can touch while
//      sync holds) into shared memory with coalesced loads -- the chains then run at shared-memory
//      latency instead of L2 latency;
//   2. the 2W vsync candidates are integrated in parallel, 4 lines at a time, one per warp;
//   3. warp 0 runs the hsync chain (window prefix sums by byte dot-products, no shuffles) and
//      publishes hsync[k]; warp 1 trails it with the burst-lock chain (4 lanes per colour row) and
//      emits the per-line records k_lines consumes.  The two chains overlap.
// Lines whose windows leave the staged heads (sync lost, |hsync| large) fall back to global loads.
#pragma once

#include "crt_kernels.cuh"

namespace crt {

constexpr int kHeadBefore = 16;                 // staged bytes before each line start
constexpr int kHeadAfter = ((kCbBeg + kBurstLen + 40 + 15) / 16) * 16; // ... and after it
constexpr int kHeadWords = (kHeadBefore + kHeadAfter) / 4 + 1;        // +1: start is aligned down to 4
constexpr int kSyncSmem = kVres * kHeadWords * 4;
constexpr int kSyncThreads = 128;

struct SyncLine { // what depends only on k, vsync and the detected field (not on the chains)
    short jl;   // signal line the decoded line reads: posmod(top + k + vsync, vres)
    short row;  // colour row: ypos % CC_VPER
    int ypos;   // posmod(top + k + vsync + 3, vres)
    int beg, end; // output rows, beg = -1 when the line is skipped (crt_core.c:428-432)
};

struct SyncShared {
    SyncLine ln[kLines];
    int hs[kLines];     // hsync after each decoded line's search
    int ccr[kLines][4]; // burst-lock accumulator of the line's colour row after its 10 steps
    volatile int ready; // lines published by the hsync warp
    int vs_found[2 * kVsyncWindow]; // per vsync candidate: crossing index or -1
    int generic;
};

__global__ void __launch_bounds__(kSyncThreads) k_sync(const MonCfg *__restrict__ cfgs, MonState *__restrict__ states,
                                                       LineRec *__restrict__ lines_base,
                                                       const signed char *__restrict__ inp_base, int first,
                                                       int force_generic)
{
    extern __shared__ __align__(16) unsigned heads[]; // [kVres][kHeadWords]
    __shared__ SyncShared sh;
    const int m = first + blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const MonCfg cfg = cfgs[m];
    if (cfg.bpp == 0) return; // crt_core.c:312-315
    MonState *st = &states[m];
    const signed char *inp = inp_base + (size_t) m * kSignalBytes;
    LineRec *lines = lines_base + (size_t) m * kLines;

    // ---- 1. stage line heads: heads[j][w] = the aligned word at ((j * H - 16) & ~3) + 4w
    for (int idx = tid; idx < kVres * kHeadWords; idx += kSyncThreads) {
        const int j = idx / kHeadWords, w = idx - j * kHeadWords;
        const int p = ((j * kHres - kHeadBefore) & ~3) + 4 * w;
        heads[idx] = (p >= 0) ? __ldg(reinterpret_cast<const unsigned *>(inp + p)) : 0u;
    }
    if (tid == 0) {
        sh.ready = 0;
        sh.generic = force_generic || abs(cfg.brightness - (kBlack + cfg.black_point)) > 4096;
    }

    // ---- 2. vsync (crt_core.c:379-396): candidate c = line posmod(vsync + c - W); first crossing wins
    const int vs_in = st->vsync;
    constexpr int kSeg = (kHres + 31) / 32;
    for (int c = warp; c < 2 * kVsyncWindow; c += kSyncThreads / 32) {
        const signed char *sig = inp + posmod(vs_in + c - kVsyncWindow, kVres) * kHres;
        const int b0 = lane * kSeg, b1 = min(kHres, b0 + kSeg);
        int sum = 0;
        for (int t = b0; t < b1; t++) sum += __ldg(sig + t);
        int acc = warp_scan_incl(sum, lane) - sum, idx = -1;
        for (int t = b0; t < b1; t++) {
            acc += __ldg(sig + t);
            if (idx < 0 && acc <= kVsyncLevel) idx = t;
        }
        const unsigned hit = __ballot_sync(0xffffffffu, idx >= 0);
        const int j = hit ? __shfl_sync(0xffffffffu, idx, __ffs(hit) - 1) : -1;
        if (lane == 0) sh.vs_found[c] = j;
    }
    __syncthreads();
    int vs = posmod(vs_in + kVsyncWindow - 1, kVres), jcross = kHres; // "gave up" defaults
    for (int c = 0; c < 2 * kVsyncWindow; c++) {
        if (sh.vs_found[c] >= 0) {
            vs = posmod(vs_in + c - kVsyncWindow, kVres);
            jcross = sh.vs_found[c];
            break;
        }
    }
    int field = (jcross > kHres / 2);
    const int ratio = (((cfg.outh << 16) / kLines) + 32768) >> 16; // crt_core.c:404-407
    field *= ratio / 2;
    for (int k = tid; k < kLines; k += kSyncThreads) { // chain-independent per-line geometry
        SyncLine g;
        int beg = (int) ((unsigned) k * ((unsigned) cfg.outh + cfg.v_fac) / (unsigned) kLines + (unsigned) field);
        int end = (int) ((unsigned) (k + 1) * ((unsigned) cfg.outh + cfg.v_fac) / (unsigned) kLines + (unsigned) field);
        if (beg >= cfg.outh) beg = end = -1; // crt_core.c:431
        else if (end > cfg.outh) end = cfg.outh;
        g.jl = (short) posmod(kTop + k + vs, kVres);
        g.ypos = posmod(kTop + k + vs + 3, kVres);
        g.row = (short) (g.ypos % kVper);
        g.beg = beg;
        g.end = end;
        sh.ln[k] = g;
    }
    __syncthreads();

    if (warp == 0) {
        // ---- 3a. hsync chain (crt_core.c:437-450), lanes 0..2W-1 hold the window prefix sums
        int hs = st->hsync;
        // byte masks: lane j sums window bytes 0..j -> 0x01 in every byte position <= j
        unsigned mask[4];
#pragma unroll
        for (int w = 0; w < 4; w++) {
            unsigned v = 0;
#pragma unroll
            for (int b = 0; b < 4; b++)
                if (4 * w + b <= lane && 4 * w + b < 2 * kHsyncWindow) v |= 1u << (8 * b);
            mask[w] = v;
        }
        for (int k = 0; k < kLines; k++) {
            if (sh.ln[k].beg >= 0) { // skipped lines leave hsync alone (crt_core.c:431)
                const int jl = sh.ln[k].jl;
                const int j = (hs > kHres / 2) ? jl + 1 : jl;             // line whose head holds the window
                const int p0 = jl * kHres + hs + kSyncBeg - kHsyncWindow; // first window sample
                const int off = p0 - ((j * kHres - kHeadBefore) & ~3);
                int prefix;
                if (j < kVres && off >= 0 && off + 20 <= kHeadWords * 4) {
                    const unsigned *wp = heads + j * kHeadWords + (off >> 2);
                    const int shft = 8 * (off & 3);
                    const unsigned w0 = wp[0], w1 = wp[1], w2 = wp[2], w3 = wp[3], w4 = wp[4];
                    prefix = __dp4a((int) __funnelshift_r(w0, w1, shft), (int) mask[0], 0);
                    prefix = __dp4a((int) __funnelshift_r(w1, w2, shft), (int) mask[1], prefix);
                    prefix = __dp4a((int) __funnelshift_r(w2, w3, shft), (int) mask[2], prefix);
                    prefix = __dp4a((int) __funnelshift_r(w3, w4, shft), (int) mask[3], prefix);
                } else { // window outside the staged heads: plain loads + scan
                    int v = 0;
                    if (lane < 2 * kHsyncWindow) v = __ldg(inp + p0 + lane);
                    prefix = warp_scan_incl(v, lane);
                }
                const unsigned hit = __ballot_sync(0xffffffffu, lane < 2 * kHsyncWindow && prefix <= kHsyncLevel);
                hs += (hit ? __ffs(hit) - 1 : 2 * kHsyncWindow) - kHsyncWindow;
                if (hs < 0) hs += kHres; // POSMOD(i + hsync, HRES), |i| <= W
                else if (hs >= kHres) hs -= kHres;
            }
            if (lane == 0) {
                sh.hs[k] = hs;
                __threadfence_block();
                sh.ready = k + 1;
            }
        }
        if (lane == 0) {
            st->vsync = vs;
            st->hsync = hs;
            st->field = field;
            if (!kIsVhs) st->rn = (int) ((unsigned) st->rn * kLcgField.mul + kLcgField.add); // crt_core.c:367
        }
    } else if (warp == 1) {
        // ---- 3b. burst lock (crt_core.c:456-467).  Lane = 4 * row + phase; trails the hsync warp.
        const int row_of_lane = lane >> 2, phase = lane & 3;
        const bool chain_lane = lane < 4 * kVper;
        int x = chain_lane ? st->ccf[row_of_lane][phase] : 0;
        const int t0 = (phase - kCbBeg) & 3; // burst samples of this phase: t0, t0 + 4, ...
        for (int k = 0; k < kLines; k++) {
            const SyncLine g = sh.ln[k];
            if (g.beg < 0) continue;
            while (sh.ready <= k) { }
            __threadfence_block();
            const int hs = sh.hs[k];
            if (chain_lane && row_of_lane == g.row) {
                const int pb = g.jl * kHres + (hs & ~3) + kCbBeg + t0; // this phase's first burst sample
                const int j = (hs > kHres / 2) ? g.jl + 1 : g.jl;
                const int off = pb - ((j * kHres - kHeadBefore) & ~3);
                int smp[kBurstLen / 4];
                if (j < kVres && off >= 0 && off + kBurstLen <= kHeadWords * 4) {
                    const signed char *hb = reinterpret_cast<const signed char *>(heads + j * kHeadWords) + off;
#pragma unroll
                    for (int q = 0; q < kBurstLen / 4; q++) smp[q] = hb[4 * q];
                } else {
#pragma unroll
                    for (int q = 0; q < kBurstLen / 4; q++) smp[q] = __ldg(inp + pb + 4 * q);
                }
                // ccr = ccr * 127 / 128 + sample, C division truncating towards zero.  While no product can
                // wrap, trunc(127 x / 128) == x - ((x + (x > 0 ? 127 : 0)) >> 7), a shorter dependent chain.
                if (abs(x) < (1 << 23)) {
#pragma unroll
                    for (int q = 0; q < kBurstLen / 4; q++) x = x - ((x + (x > 0 ? 127 : 0)) >> 7) + smp[q];
                } else {
#pragma unroll
                    for (int q = 0; q < kBurstLen / 4; q++) x = wadd(wmul(x, 127) / 128, smp[q]);
                }
                sh.ccr[k][phase] = x;
            }
        }
        if (chain_lane) st->ccf[row_of_lane][phase] = x;
    }
    __syncthreads();

    // ---- 4. per-line records for k_lines (crt_core.c:452-454, 469-479), all threads
    int huesn, huecs;
    {
        int sn, cs;
        sincos14_d(sn, cs, ((cfg.hue % 360) + 33) * 8192 / 180); // crt_core.c:318-320
        huesn = sn >> 11;
        huecs = cs >> 11;
    }
    for (int k = tid; k < kLines; k += kSyncThreads) {
        const SyncLine g = sh.ln[k];
        const int hs = sh.hs[k];
        LineRec rec;
        rec.pad0 = rec.pad1 = 0;
        rec.hsync = hs;
        rec.beg = g.beg;
        rec.end = g.end;
        rec.pos = 0;
        rec.wave0 = rec.wave1 = 0;
        if (g.beg >= 0) {
            const int pa = hs & 3;
            const int dci = wsub(sh.ccr[k][(pa + 1) & 3], sh.ccr[k][(pa + 3) & 3]);
            const int dcq = wsub(sh.ccr[k][(pa + 2) & 3], sh.ccr[k][pa]);
            rec.pos = posmod(kAvBeg + hs - 3, kHres) + g.ypos * kHres;
            rec.wave0 = wmul(wsub(wmul(dci, huecs), wmul(dcq, huesn)) >> 4, cfg.saturation);
            rec.wave1 = wmul(wadd(wmul(dcq, huecs), wmul(dci, huesn)) >> 4, cfg.saturation);
            // the fast equaliser path of k_lines is exact while |wave| <= 65536 (crt_lines.cuh)
            if (abs(rec.wave0) > 65536 || abs(rec.wave1) > 65536) sh.generic = 1;
        }
        lines[k] = rec;
    }
    __syncthreads();
    if (tid == 0) st->generic = sh.generic;
}

} // namespace crt
