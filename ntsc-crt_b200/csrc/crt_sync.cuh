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
//   3. the hsync chain is solved by verified speculation (parallel sweeps to a fixed point, see 3a);
//      one warp then runs the burst-lock chain, 4 lanes per colour row, next line prefetched;
//   4. all threads emit the per-line records k_lines consumes.
// Lines whose windows leave the staged heads (sync lost, |hsync| large) fall back to global loads.
#pragma once

#include "crt_kernels.cuh"

namespace crt {

constexpr int kHeadBefore = 16;                 // staged bytes before each line start
constexpr int kHeadAfter = ((kCbBeg + kBurstLen + 40 + 15) / 16) * 16; // ... and after it
constexpr int kHeadWords = (kHeadBefore + kHeadAfter) / 4 + 1;        // +1: start is aligned down to 4
constexpr int kCandWords = (kHres + 3) / 4 + 1;                     // a whole line from a 4-aligned start
// One head more than there are signal lines: with hsync in the second half of a line (the steady state of the
// NTSC timing, ~898) the search window of the LAST signal line sits on the start of the line after it.  Those
// bytes are the padding after inp[] (the reference reads past its array there); staging them keeps that one
// line off the 16-dependent-loads fall-back, which used to hold every sweep's barrier for ~5 us.
constexpr int kHeadLines = kVres + 1;
constexpr int kSyncSmem = (kHeadLines * kHeadWords + 2 * kVsyncWindow * kCandWords) * 4;
constexpr int kSyncThreads = 256;

// CRT_CC_SAMPLES == 5 (crt_core.c:497-508): carrier tables rotated by the hue knob, one I and one Q value per phase
__device__ __forceinline__ void pv1k_waves(int dci, int dcq, int hue, int saturation, int (&wi)[5], int (&wq)[5])
{
    int ang = hue % 360;
#pragma unroll
    for (int i = 0; i < 5; i++) {
        int sn, cs;
        sincos14_d(sn, cs, ang * 8192 / 180);
        wi[i] = wmul(wadd(wmul(dci, cs), wmul(dcq, sn)) >> 15, saturation);
        sincos14_d(sn, cs, (ang + 90) * 8192 / 180);
        wq[i] = wmul(wadd(wmul(dci, cs), wmul(dcq, sn)) >> 15, saturation);
        ang += 360 / 5;
    }
}

struct SyncLine { // what depends only on k, vsync and the detected field (not on the chains)
    short jl;   // signal line the decoded line reads: posmod(top + k + vsync, vres)
    short row;  // colour row: ypos % CC_VPER
    int ypos;   // posmod(top + k + vsync + 3, vres)
    int beg, end; // output rows, beg = -1 when the line is skipped (crt_core.c:428-432)
};

struct SyncShared {
    SyncLine ln[kLines];
    int hs[kLines];     // hsync after each decoded line's search
    int ccr[kLines][kCc]; // burst-lock accumulator of the line's colour row after its 10 steps
    signed char burst[kLines][kBurstLen]; // the 40 (PV-1000: 50) burst samples each decoded line locks onto
    short rowlist[kVper > 3 ? kVper : 3][kLines]; // decoded lines of each colour row, in order
    int rowcount[kVper > 3 ? kVper : 3];
    int vs_found[2 * kVsyncWindow]; // per vsync candidate: crossing index or -1
    int generic;
};

// hsync search of one decoded line given the hsync it enters with (crt_core.c:437-447):
// integrate 2W samples starting W before the expected sync edge, stop at the threshold.
template <typename FetchByte>
__device__ __forceinline__ int hsync_step(const unsigned *heads, FetchByte fetch_byte, int jl, int hs)
{
    const int p0 = jl * kHres + hs + kSyncBeg - kHsyncWindow; // first window sample
    const int j = (hs > kHres / 2) ? jl + 1 : jl;             // line whose staged head holds the window
    const int off = p0 - ((j * kHres - kHeadBefore) & ~3);
    int i = 2 * kHsyncWindow, acc = 0;
    if (j < kHeadLines && off >= 0 && off + 2 * kHsyncWindow <= kHeadWords * 4) {
        const signed char *hb = reinterpret_cast<const signed char *>(heads + j * kHeadWords) + off;
#pragma unroll
        for (int t = 0; t < 2 * kHsyncWindow; t++) {
            acc += hb[t];
            if (acc <= kHsyncLevel && i == 2 * kHsyncWindow) i = t;
        }
    } else { // window outside the staged heads (sync lost): plain loads
        for (int t = 0; t < 2 * kHsyncWindow; t++) {
            acc += fetch_byte(p0 + t);
            if (acc <= kHsyncLevel && i == 2 * kHsyncWindow) i = t;
        }
    }
    hs += i - kHsyncWindow;
    if (hs < 0) hs += kHres; // POSMOD(i + hsync, HRES), |i| <= W
    else if (hs >= kHres) hs -= kHres;
    return hs;
}

// In the reference inp[] is followed, inside struct CRT, by outw, outh, out_format and four bytes of padding
// (crt_core.h:74-92; crt_init zeroes the struct).  A sync search or decode window that runs a few samples past the
// end of inp[] -- the PV-1000 does that in ordinary operation: its picture ends one sample before the end of the
// line and hsync settles above 4 -- reads those bytes.  They are deterministic, so they are reproduced: 16 bytes
// behind each signal buffer (the kernels stage from analog[] when the noise pass is fused, from inp[] otherwise).
// What follows them in the reference is the `out` pointer; windows that reach it are outside the parity domain.
__global__ void __launch_bounds__(64) k_struct_tail(const MonCfg *__restrict__ cfgs, signed char *__restrict__ analog_base,
                                                    signed char *__restrict__ inp_base, int first, int count)
{
    const int i = blockIdx.x * 64 + threadIdx.x;
    if (i >= count) return;
    const MonCfg c = cfgs[first + i];
    const int v[4] = { c.outw, c.outh, c.out_format, 0 };
    signed char *a = analog_base + (size_t) (first + i) * kSignalBytes + kInputSize;
    signed char *b = inp_base + (size_t) (first + i) * kSignalBytes + kInputSize;
    for (int k = 0; k < 16; k++) { // (byte stores: CRT_INPUT_SIZE is not a multiple of 4 in every system)
        const signed char byte = (signed char) ((unsigned) v[k >> 2] >> (8 * (k & 3)));
        a[k] = byte;
        b[k] = byte;
    }
}

// Four samples of inp[] starting at the 4-aligned index p, computed from the word `w` of analog[] exactly
// as the noise pass does (crt_core.c:346-367): sample i uses the LCG state advanced i + 1 steps from the
// call's seed.  Kept apart from the load so that callers can issue a batch of loads before touching any.
__device__ __forceinline__ unsigned noisy_apply(unsigned w, int p, int noise, unsigned rn0,
                                                const Affine *__restrict__ jump_lo, const Affine *__restrict__ jump_hi)
{
    const unsigned raw = w;
    if (noise == 0) {
        w = __vmaxs4(w, 0x81818181u);
    } else {
        const int t = p / kNoiseVec;
        const Affine lo = jump_lo[t % kJumpLo], hi = jump_hi[t / kJumpLo];
        unsigned rn = (rn0 * hi.mul + hi.add) * lo.mul + lo.add; // state before sample 16 t
        for (int k = 0; k < (p % kNoiseVec); k++) rn = rn * kLcgMul + kLcgAdd;
        unsigned o = 0;
#pragma unroll
        for (int b = 0; b < 4; b++) {
            rn = rn * kLcgMul + kLcgAdd;
            int v = (int) (signed char) (w >> (8 * b)) + (wmul((int) ((rn >> 16) & 0xff) - 0x7f, noise) >> 8);
            o |= ((unsigned) clampi(v, -127, 127) & 0xffu) << (8 * b);
        }
        w = o;
    }
    // beyond inp[] lie the bytes k_struct_tail put there (what follows inp[] inside the reference's struct CRT),
    // identical after analog[] and after inp[], and never touched by the noise pass: keep them raw
    if (p + 4 > kInputSize) {
        const unsigned keep = (p >= kInputSize) ? 0u : (0xffffffffu >> (8 * (p + 4 - kInputSize))); // bytes still inside inp[]
        w = (w & keep) | (raw & ~keep);
    }
    return w;
}

__device__ __forceinline__ unsigned noisy_word(const signed char *__restrict__ analog, int p, int noise, unsigned rn0,
                                               const Affine *__restrict__ jump_lo, const Affine *__restrict__ jump_hi)
{
    return noisy_apply(__ldg(reinterpret_cast<const unsigned *>(analog + p)), p, noise, rn0, jump_lo, jump_hi);
}

// FUSED (LCG systems): the kernel reads analog[], applies the noise itself to what it stages, and its
// otherwise idle warps write the whole inp[] while warp 0 runs the burst-lock chain -- the separate noise
// pass disappears.  !FUSED (VHS, whose noise comes from rand()): inp[] was written by k_noise_terms.
template <bool FUSED>
__global__ void __launch_bounds__(kSyncThreads) k_sync(const MonCfg *__restrict__ cfgs, MonState *__restrict__ states,
                                                       LineRec *__restrict__ lines_base,
                                                       const signed char *__restrict__ analog_base,
                                                       signed char *__restrict__ inp_base,
                                                       const Affine *__restrict__ jump_lo,
                                                       const Affine *__restrict__ jump_hi, int first,
                                                       int force_generic)
{
    grid_dep_launch();
    grid_dep_wait(); // (programmatic launch behind the encoder: analog[] must be complete)
    extern __shared__ __align__(16) unsigned heads[]; // [kHeadLines][kHeadWords]
    __shared__ SyncShared sh;
    const int m = first + blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const MonCfg cfg = cfgs[m];
    if (cfg.bpp == 0) return; // crt_core.c:312-315
    MonState *st = &states[m];
    const signed char *analog = analog_base + (size_t) m * kSignalBytes;
    signed char *inp_w = inp_base + (size_t) m * kSignalBytes;
    const signed char *inp = inp_w; // !FUSED reads; fall-back loads when sync is lost (see below)
    LineRec *lines = lines_base + (size_t) m * kLines;
    const unsigned rn0 = (unsigned) st->rn;
    const int noise = cfg.noise;
    auto fetch = [&](int p) { // one staged word
        if (p < 0) return 0u;
        if (FUSED) return noisy_word(analog, p, noise, rn0, jump_lo, jump_hi);
        return __ldg(reinterpret_cast<const unsigned *>(inp + p));
    };
    auto fetch_byte = [&](int p) { // a single sample outside the staged regions (sync lost)
        const unsigned w = fetch(p & ~3);
        return (int) (signed char) (w >> (8 * (p & 3)));
    };

    // ---- 1. stage line heads (heads[j][w] = the aligned word at ((j * H - 16) & ~3) + 4w) and the 2W
    // vsync candidate lines in full (cand[c][w] = aligned words covering line posmod(vsync + c - W)).
    // All loads of a batch are issued before any is stored, so the copy runs at memory-level parallelism
    // instead of one L2 round trip per word.
    unsigned *cand = heads + kHeadLines * kHeadWords; // [2W][kCandWords]
    const int vs_in = st->vsync;
    {
        constexpr int kBatch = 8;
        constexpr int kHeadTotal = kHeadLines * kHeadWords, kCandTotal = 2 * kVsyncWindow * kCandWords;
        const signed char *from = FUSED ? analog : inp;
        for (int base = 0; base < kHeadTotal + kCandTotal; base += kBatch * kSyncThreads) {
            unsigned v[kBatch];
            int pos[kBatch];
#pragma unroll
            for (int b = 0; b < kBatch; b++) { // where each word comes from (-1: nothing to load)
                const int idx = base + b * kSyncThreads + tid;
                pos[b] = -1;
                if (idx < kHeadTotal) {
                    const int j = idx / kHeadWords, w = idx - j * kHeadWords;
                    pos[b] = ((j * kHres - kHeadBefore) & ~3) + 4 * w;
                } else if (idx < kHeadTotal + kCandTotal) {
                    const int q = idx - kHeadTotal, c = q / kCandWords, w = q - c * kCandWords;
                    pos[b] = (posmod(vs_in + c - kVsyncWindow, kVres) * kHres & ~3) + 4 * w;
                }
            }
#pragma unroll
            for (int b = 0; b < kBatch; b++) // the raw loads, back to back
                v[b] = (pos[b] >= 0) ? __ldg(reinterpret_cast<const unsigned *>(from + pos[b])) : 0u;
#pragma unroll
            for (int b = 0; b < kBatch; b++) {
                const int idx = base + b * kSyncThreads + tid;
                if (FUSED && pos[b] >= 0) v[b] = noisy_apply(v[b], pos[b], noise, rn0, jump_lo, jump_hi);
                if (idx < kHeadTotal + kCandTotal) heads[idx] = v[b];
            }
        }
    }
    // |bright| bound of the fast equaliser path (crt_lines.cuh); halved for the PV-1000, whose luma cascade
    // (hf = 80024) is only proven wrap-free up to there
    if (tid == 0) sh.generic = force_generic || abs(cfg.brightness - (kBlack + cfg.black_point)) > (kCc == 5 ? 2048 : 4096);
    __syncthreads();

    // ---- 2. vsync (crt_core.c:379-396): candidate c = line posmod(vsync + c - W); first crossing wins
    constexpr int kSeg = (kHres + 31) / 32;
    for (int c = warp; c < 2 * kVsyncWindow; c += kSyncThreads / 32) {
        const int lstart = posmod(vs_in + c - kVsyncWindow, kVres) * kHres;
        const signed char *sig = reinterpret_cast<const signed char *>(cand + c * kCandWords) + (lstart & 3);
        const int b0 = lane * kSeg, b1 = min(kHres, b0 + kSeg);
        int sum = 0;
        for (int t = b0; t < b1; t++) sum += sig[t];
        int acc = warp_scan_incl(sum, lane) - sum, idx = -1;
        for (int t = b0; t < b1; t++) {
            acc += sig[t];
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
    const int hs_in = st->hsync;
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
        sh.hs[k] = hs_in; // initial guess of the chain
    }
    __syncthreads();
    // Decoded lines of each colour row, in order (skipped lines do not touch ccf).  Warp r compacts row r
    // with ballots, 32 lines per step -- a 240-iteration loop on one thread per row here made every other
    // thread wait ~13 % of the kernel at the next barrier.
    if (warp < kVper) {
        int n = 0;
        for (int k0 = 0; k0 < kLines; k0 += 32) {
            const int k = k0 + lane;
            const bool mine = (k < kLines) && sh.ln[k].beg >= 0 && sh.ln[k].row == warp;
            const unsigned m = __ballot_sync(0xffffffffu, mine);
            if (mine) sh.rowlist[warp][n + __popc(m & ((1u << lane) - 1u))] = (short) k;
            n += __popc(m);
        }
        if (lane == 0) sh.rowcount[warp] = n;
    }

    // ---- 3a. hsync chain hs[k] = f_k(hs[k-1]) (crt_core.c:437-450) by verified speculation: every line
    // is evaluated in parallel from the current guess of its predecessor; a sweep that changes nothing
    // proves hs[] is the chain's unique solution.  The sync edge recaptures the search from any start
    // within the window, so this takes a handful of sweeps instead of 240 dependent steps.
    for (int sweep = 0; sweep <= kLines; sweep++) {
        int nh[(kLines + kSyncThreads - 1) / kSyncThreads];
        int changed = 0;
#pragma unroll
        for (int q = 0; q < (kLines + kSyncThreads - 1) / kSyncThreads; q++) {
            const int k = tid + q * kSyncThreads;
            nh[q] = 0;
            if (k < kLines) {
                const int prev = (k == 0) ? hs_in : sh.hs[k - 1];
                nh[q] = (sh.ln[k].beg >= 0) ? hsync_step(heads, fetch_byte, sh.ln[k].jl, prev) : prev; // crt_core.c:431
                changed |= (nh[q] != sh.hs[k]);
            }
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < (kLines + kSyncThreads - 1) / kSyncThreads; q++) {
            const int k = tid + q * kSyncThreads;
            if (k < kLines) sh.hs[k] = nh[q];
        }
        if (!__syncthreads_or(changed)) break;
    }

    // ---- 3b. burst lock (crt_core.c:456-467): ccr = ccr * 127 / 128 + sample, 10 samples per phase per line.
    // First every thread gathers the 40 burst samples of its lines (their position depends on hsync, now
    // known) into shared memory, so the serial chain below is nothing but the recurrence.
    for (int idx = tid; idx < kLines * kBurstLen; idx += kSyncThreads) {
        const int k = idx / kBurstLen, t = idx - k * kBurstLen;
        int v = 0;
        if (sh.ln[k].beg >= 0) {
            const int hs = sh.hs[k], jl = sh.ln[k].jl;
            const int p = jl * kHres + (hs - hs % kCc) + kCbBeg + t; // crt_core.c:458-462 (hs >= 0: "& ~3" when kCc == 4)
            const int j = (hs > kHres / 2) ? jl + 1 : jl;
            const int off = p - ((j * kHres - kHeadBefore) & ~3);
            if (j < kHeadLines && off >= 0 && off < kHeadWords * 4)
                v = (int) (signed char) (heads[j * kHeadWords + (off >> 2)] >> (8 * (off & 3)));
            else
                v = fetch_byte(p);
        }
        sh.burst[k][t] = (signed char) v;
    }
    __syncthreads();
    if (warp == 0) {
        // Lane = kCc * row + phase walks its own colour row's lines (at most 5 x 5 = 25 lanes).
        static_assert(kCc * kVper <= 32, "one lane per (colour row, phase)");
        const int row = lane / kCc, phase = lane % kCc;
        const bool chain_lane = lane < kCc * kVper;
        const int t0 = posmod(phase - kCbBeg, kCc); // burst samples of this phase: t0, t0 + kCc, ...
        int x = chain_lane ? st->ccf[row][phase] : 0;
        const int count = chain_lane ? sh.rowcount[row] : 0;
        for (int n = 0; n < count; n++) {
            const int k = sh.rowlist[row][n];
            const signed char *bs = &sh.burst[k][t0];
            // C's x * 127 / 128 truncates towards zero.  While the product cannot wrap it equals
            // x - ((x + (x >= 0 ? 127 : 0)) >> 7): ceil(x / 128) for x >= 0, floor for x < 0.
            if (abs(x) < (1 << 23)) {
                // The sign of x almost never changes within a line, so the rounding bias is taken from the
                // line's first x and the 10 dependent steps are add, shift, add; the sign bits of the
                // intermediate values are collected off the critical path and the rare line on which one
                // differs is redone with the per-step bias.
                const int bias = (x >= 0) ? 127 : 0;
                int y = x, flips = 0;
#pragma unroll
                for (int q = 0; q < kBurstLen / kCc; q++) {
                    flips |= y ^ x;
                    y = y - ((y + bias) >> 7) + bs[kCc * q];
                }
                if (flips < 0) {
#pragma unroll
                    for (int q = 0; q < kBurstLen / kCc; q++) x = x - ((x + ((x >= 0) ? 127 : 0)) >> 7) + bs[kCc * q];
                } else {
                    x = y;
                }
            } else {
#pragma unroll
                for (int q = 0; q < kBurstLen / kCc; q++) x = wadd(wmul(x, 127) / 128, bs[kCc * q]);
            }
            sh.ccr[k][phase] = x;
        }
        if (chain_lane) st->ccf[row][phase] = x;
        if (lane == 0) {
            st->vsync = vs;
            st->hsync = sh.hs[kLines - 1];
            st->field = field;
            if (!kIsVhs) st->rn = (int) (rn0 * kLcgField.mul + kLcgField.add); // crt_core.c:367
        }
    } else if (FUSED) {
        // ---- 3c. the noise pass proper (crt_core.c:346-367), by the 7 warps that would otherwise wait:
        // analog -> inp, 16 samples per thread per step, 128-bit accesses
        constexpr int kNB = 8; // loads in flight per thread: the 7 warps must cover DRAM latency by themselves (4 in round 1)
        for (int tb = tid - 32; tb < kNoiseThreads; tb += kNB * (kSyncThreads - 32)) {
            uint4 in[kNB];
#pragma unroll
            for (int u = 0; u < kNB; u++) {
                const int t = tb + u * (kSyncThreads - 32);
                in[u] = make_uint4(0u, 0u, 0u, 0u);
                if (t < kNoiseThreads) in[u] = *reinterpret_cast<const uint4 *>(analog + t * kNoiseVec);
            }
#pragma unroll
            for (int u = 0; u < kNB; u++) {
                const int t = tb + u * (kSyncThreads - 32);
                if (t >= kNoiseThreads) continue;
                const int i0 = t * kNoiseVec;
                unsigned w[4] = { in[u].x, in[u].y, in[u].z, in[u].w };
                if (noise == 0) {
#pragma unroll
                    for (int k = 0; k < 4; k++) w[k] = __vmaxs4(w[k], 0x81818181u);
                } else {
                    const Affine lo = jump_lo[t % kJumpLo], hi = jump_hi[t / kJumpLo];
                    unsigned rn = (rn0 * hi.mul + hi.add) * lo.mul + lo.add;
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        unsigned o = 0;
#pragma unroll
                        for (int b = 0; b < 4; b++) {
                            rn = rn * kLcgMul + kLcgAdd;
                            int v = (int) (signed char) (w[k] >> (8 * b)) + (wmul((int) ((rn >> 16) & 0xff) - 0x7f, noise) >> 8);
                            o |= ((unsigned) clampi(v, -127, 127) & 0xffu) << (8 * b);
                        }
                        w[k] = o;
                    }
                }
                if (i0 + kNoiseVec <= kInputSize) {
                    *reinterpret_cast<uint4 *>(inp_w + i0) = make_uint4(w[0], w[1], w[2], w[3]);
                } else {
                    for (int b = 0; i0 + b < kInputSize; b++) inp_w[i0 + b] = (signed char) (w[b >> 2] >> (8 * (b & 3)));
                }
            }
        }
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
        // When the output has fewer rows than there are decoded lines, consecutive lines land on the same
        // row and the reference applies them in order (each one blending onto, or overwriting, its
        // predecessor).  pad0 = position of this line within its run, pad1 = 1 for the run's last line;
        // the host then launches the line kernel once per position (crtx.cu).
        int rank = 0;
        for (int j = k - 1; j >= 0 && g.beg >= 0 && sh.ln[j].beg == g.beg; j--) rank++;
        rec.pad0 = rank;
        rec.pad1 = (k == kLines - 1 || sh.ln[k + 1].beg != g.beg) ? 1 : 0;
        rec.hsync = hs;
        rec.beg = g.beg;
        rec.end = g.end;
        rec.pos = 0;
        rec.wave0 = rec.wave1 = 0;
        if (g.beg >= 0) {
            rec.pos = posmod(kAvBeg + hs - 3, kHres) + g.ypos * kHres;
            long long wmax;
            if (kCc == 4) {
                const int pa = hs & 3;
                const int dci = wsub(sh.ccr[k][(pa + 1) & 3], sh.ccr[k][(pa + 3) & 3]);
                const int dcq = wsub(sh.ccr[k][(pa + 2) & 3], sh.ccr[k][pa]);
                rec.wave0 = wmul(wsub(wmul(dci, huecs), wmul(dcq, huesn)) >> 4, cfg.saturation);
                rec.wave1 = wmul(wadd(wmul(dcq, huecs), wmul(dci, huesn)) >> 4, cfg.saturation);
                wmax = max(llabs((long long) rec.wave0), llabs((long long) rec.wave1));
            } else { // crt_core.c:480-509: the record carries dci / dcq, the line kernel rebuilds the ten carrier values
                const int pa = hs % kCc, peak = pa + kCc / 4;
                const int *ccr = sh.ccr[k];
                const int dci = wsub(ccr[peak % kCc], wadd(ccr[(peak + kCc / 2) % kCc], ccr[(peak + kCc / 2 + 1) % kCc]) / 2);
                const int dcq = wsub(ccr[(pa + kCc / 2) % kCc], ccr[pa % kCc]);
                rec.wave0 = dci;
                rec.wave1 = dcq;
                int wi[5], wq[5];
                pv1k_waves(dci, dcq, cfg.hue, cfg.saturation, wi, wq);
                wmax = 0;
#pragma unroll
                for (int i = 0; i < 5; i++) wmax = max(wmax, max(llabs((long long) wi[i]), llabs((long long) wq[i])));
            }
            // The fast equaliser path of k_lines is exact while every chroma input (s * wave) >> 9 stays
            // within +-16383 (crt_lines.cuh).  |s| is bounded by the largest sample of the one or two
            // signal lines the decode window covers -- measured by the noise warps when the noise pass is
            // fused, 127 (the clamp of crt_core.c:363-364) otherwise.
            // |s| <= 127 always (the clamp of crt_core.c:363-364).  With the stock saturation that bound already passes and
            // nothing more is needed; only a line whose carrier is large enough to fail it has its two signal lines scanned for
            // the real maximum -- by this one thread, from the inp[] this CTA wrote before the barrier above.  (Round 1 tracked
            // the maximum of every 16-byte chunk inside the noise loop: 13 % of the kernel's instructions for a rare case.)
            int smax = 127;
            if (FUSED && ((smax * wmax) >> 9) + 1 > 16383) {
                const int lo = (g.ypos * kHres) & ~15, hi = min((g.ypos + 2) * kHres, kInputSize);
                int mx = 0;
                for (int i = lo; i < hi; i++) mx = max(mx, abs((int) inp_w[i]));
                smax = mx;
            }
            if (((smax * wmax) >> 9) + 1 > 16383) sh.generic = 1;
        }
        lines[k] = rec;
    }
    __syncthreads();
    if (tid == 0) st->generic = sh.generic;
}

} // namespace crt
