/* oracle/crt_oracle.h -- TEST INFRASTRUCTURE ONLY (the parity checker).
 *
 * A CPU restatement, in plain C, of the reference's composite modulate -> noise ->
 * demodulate hot path.  It is NOT part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it, and only as
 * the checker.  The product library (ntsc-crt_b200/lib/libcrt_b200_*.so) never links or
 * loads it and has no CPU fallback.
 *
 * Pinning: the reference ships no tests or golden vectors of its own (SURVEY.md
 * section 4), so this restatement is pinned against the reference ITSELF, compiled
 * unmodified into oracle/_ref/libref_*.so (oracle/Makefile) -- see
 * tests/test_oracle_vs_ref.py -- and against the md5 anchors recorded in BASELINE.md.
 *
 * Unlike the reference (compile-time polymorphic via CRT_SYSTEM, crt_core.h:39-59)
 * the oracle is runtime-polymorphic: one library, a system descriptor per variant.
 * It is also decomposed the way the CUDA pipeline is (noise pass / sync pre-pass /
 * per-line filter + resample) so every intermediate table the kernels exchange can
 * be compared, not just the final image.
 */
#ifndef CRT_ORACLE_H
#define CRT_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define OCRT_SYS_NTSC 0 /* crt_core.h:30 */
#define OCRT_SYS_NES  1 /* crt_core.h:31 */
#define OCRT_SYS_PV1K 2 /* crt_core.h:32 */
#define OCRT_SYS_SNES 3 /* crt_core.h:33 */
#define OCRT_SYS_TEMP 4 /* crt_core.h:34 */
#define OCRT_SYS_VHS  5 /* crt_core.h:35 */
#define OCRT_SYS_NESRGB 6 /* crt_core.h:36 */

#define OCRT_MAX_VPER 5 /* CRT_CC_VPER of the PV-1000 */
#define OCRT_MAX_CC   5 /* CRT_CC_SAMPLES of the PV-1000 (4 everywhere else) */
#define OCRT_PAD      2048 /* slack after analog/inp for the reference's over-reads */

/* Everything the reference derives from macros (crt_ntsc.h:25-109, crt_nes.h:30-130,
 * crt_ntscvhs.h:25-131) plus the coefficients its init code computes once
 * (crt_core.c:263-289, crt_ntsc.c:95-106). */
typedef struct ocrt_sys {
    int system, chroma_pattern;
    int hres, vres, input_size;
    int top, bot, lines;
    int cc_vper;
    int hsync_window, vsync_window, hsync_thresh, vsync_thresh;
    int sync_beg, bw_beg, cb_beg, av_beg, av_len, burst_len;
    int white_level, burst_level, black_level, blank_level, sync_level;
    int vhs_noise;              /* crt_ntscvhs.h:29 */
    int nes_vsync_end;          /* PPUpx2pos(327), crt_nes.c:91 */
    /* decoder equaliser: lf, hf, g[3] for Y, I, Q (crt_core.c:278-280) */
    int eq[3][5];
    /* encoder band-limit coefficients c for Y, I, Q (crt_ntsc.c:142-146) */
    int iir_c[3];
    /* non-zero: the USE_CONVOLUTION build of the decoder (crt_core.c:85, 96-147): eqf() is a FIR kernel of
     * this many taps (7: [1 4 7 8 7 4 1] >> 5, the stock one; 6, 5, 4: crt_core.c:86-88) instead of the
     * three-band equaliser */
    int conv;
    int cc_samples; /* CRT_CC_SAMPLES: samples per chroma period, 4 (5: PV-1000) */
    int bloom;      /* CRT_DO_BLOOM (crt_core.h:70): beam-energy dependent line width */
} ocrt_sys;

/* Mirrors the caller-visible part of struct CRT (crt_core.h:74-92). */
typedef struct ocrt_monitor {
    signed char *analog; /* input_size + OCRT_PAD, owned */
    signed char *inp;    /* input_size + OCRT_PAD, owned */
    int outw, outh, out_format;
    unsigned char *out;  /* caller-owned */
    int hue, brightness, contrast, saturation;
    int black_point, white_point;
    int scanlines, blend;
    unsigned v_fac;
    int ccf[OCRT_MAX_VPER][OCRT_MAX_CC];
    int hsync, vsync, rn;
    int last_noise; /* `noise` of the crt_demodulate call in progress (the bloom energy scale uses it, crt_core.c:400) */
} ocrt_monitor;

/* struct NTSC_SETTINGS of the RGB systems (crt_ntsc.h:111-124, crt_ntscvhs.h:133-147) */
typedef struct ocrt_rgb_source {
    const unsigned char *data;
    int format, w, h, raw, as_color, field, frame, hue, xoffset, yoffset;
    int do_aberration;    /* VHS only */
    int dot_crawl_offset; /* SNES, template and PV-1000 systems (crt_snes.h:121, crt_template.h, crt_pv1k.h) */
} ocrt_rgb_source;

/* struct NTSC_SETTINGS of the NES system (crt_nes.h:132-143) */
typedef struct ocrt_nes_source {
    const unsigned short *data;
    int w, h;
    int dot_crawl_offset, hue, xoffset, yoffset;
    int field_initialized;
} ocrt_nes_source;

/* struct NTSC_SETTINGS of the NES-RGB system (crt_nesrgb.h) */
typedef struct ocrt_nesrgb_source {
    const unsigned char *data;
    int format, w, h;
    int dot_crawl_offset, hue, xoffset, yoffset;
    int field_initialized;
} ocrt_nesrgb_source;

/* glibc TYPE_3 rand() replica (the VHS variant draws from libc rand(),
 * crt_core.c:344-351, crt_ntscvhs.c:206; glibc 2.39 stdlib/random_r.c) */
typedef struct ocrt_rand {
    unsigned r[31];
    int f, b;
} ocrt_rand;

/* what the sync pre-pass decides for one decoded scanline (crt_core.c:409-479) */
typedef struct ocrt_line {
    int skip;       /* crt_core.c:431 */
    int beg, end;   /* output rows, crt_core.c:428-432 */
    int hsync;      /* after this line's search, crt_core.c:446 */
    int pos;        /* xpos + ypos * hres, crt_core.c:452-454 */
    int wave[4];    /* crt_core.c:476-479 (CRT_CC_SAMPLES 4) */
    int wave_i[OCRT_MAX_CC], wave_q[OCRT_MAX_CC]; /* crt_core.c:480-509 (CRT_CC_SAMPLES 5) */
} ocrt_line;

const ocrt_sys *ocrt_system(int system, int chroma_pattern);
/* the same system with the reference's USE_CONVOLUTION 1 decoder (crt_core.c:85) */
const ocrt_sys *ocrt_system_conv(int system, int chroma_pattern);
const ocrt_sys *ocrt_system_conv_taps(int system, int chroma_pattern, int taps); /* 4 .. 7 */
/* the same system built with CRT_DO_BLOOM 1 (crt_core.h:70; RGB systems) */
const ocrt_sys *ocrt_system_bloom(int system, int chroma_pattern);

void ocrt_sincos14(int *s, int *c, int n);
int  ocrt_bpp(int format);

void ocrt_rand_seed(ocrt_rand *g, unsigned seed);
int  ocrt_rand_next(ocrt_rand *g);

int  ocrt_monitor_create(const ocrt_sys *sys, ocrt_monitor *m, int w, int h, int f, unsigned char *out);
void ocrt_monitor_destroy(ocrt_monitor *m);
void ocrt_monitor_reset(ocrt_monitor *m);

/* crt_modulate */
void ocrt_encode_rgb(const ocrt_sys *sys, ocrt_monitor *m, ocrt_rgb_source *src, ocrt_rand *g);
void ocrt_encode_nes(const ocrt_sys *sys, ocrt_monitor *m, ocrt_nes_source *src);
void ocrt_encode_snes(const ocrt_sys *sys, ocrt_monitor *m, ocrt_rgb_source *src);
void ocrt_encode_pv1k(const ocrt_sys *sys, ocrt_monitor *m, ocrt_rgb_source *src);
void ocrt_encode_template(const ocrt_sys *sys, ocrt_monitor *m, ocrt_rgb_source *src);
void ocrt_encode_nesrgb(const ocrt_sys *sys, ocrt_monitor *m, ocrt_nesrgb_source *src);

/* crt_demodulate, and its three stages on their own */
void ocrt_decode(const ocrt_sys *sys, ocrt_monitor *m, int noise, ocrt_rand *g);
void ocrt_noise_pass(const ocrt_sys *sys, ocrt_monitor *m, int noise, ocrt_rand *g);
int  ocrt_sync_pass(const ocrt_sys *sys, ocrt_monitor *m, ocrt_line *table /* [lines] */);
void ocrt_line_pass(const ocrt_sys *sys, ocrt_monitor *m, const ocrt_line *table, int first, int count);

/* LCG jump-ahead of the noise generator: (mul, add) such that n steps of
 * rn = 214019 * rn + 140327895 (crt_core.c:359) equal rn * mul + add (mod 2^32). */
void ocrt_lcg_jump(unsigned n, unsigned *mul, unsigned *add);

#ifdef __cplusplus
}
#endif
#endif
