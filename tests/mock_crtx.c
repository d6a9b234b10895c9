/* tests/mock_crtx.c -- TEST INFRASTRUCTURE: the handful of crtx_* entry points tools/crtx_video.c uses,
 * implemented on the CPU with the oracle, so that the driver's host logic (segments, halo, speculation,
 * verification, repair, file formats) runs in the CPU suite.  "Device" memory is plain heap memory here.
 * Never part of the product: tests/test_video_driver_logic.py builds it into a scratch directory under the
 * library name the driver links against. */
#include <stdlib.h>
#include <string.h>

#include "crtx_batch.h"
#include "crt_oracle.h"

struct crtx_ctx {
    int n;
    const ocrt_sys *sys;
    ocrt_monitor *mon;
    int *noise;
    long launches;
};

static const char *g_err = "";
const char *crtx_last_error(void) { return g_err; }
int crtx_input_size(void) { return ocrt_system(OCRT_SYS_NTSC, 1)->input_size; }
long crtx_launch_count(crtx_ctx *ctx) { return ctx->launches; }
long crtx_lines2_count(crtx_ctx *ctx) { (void) ctx; return 0; }

int crtx_create(crtx_ctx **out, int n)
{
    crtx_ctx *c = (crtx_ctx *) calloc(1, sizeof(*c));
    int i;
    c->n = n;
    c->sys = ocrt_system(OCRT_SYS_NTSC, 1);
    c->mon = (ocrt_monitor *) calloc((size_t) n, sizeof(ocrt_monitor));
    c->noise = (int *) calloc((size_t) n, sizeof(int));
    for (i = 0; i < n; i++) ocrt_monitor_create(c->sys, &c->mon[i], 16, 16, 5, NULL); /* zeroed signals, rn 194 */
    *out = c;
    return 0;
}

void crtx_destroy(crtx_ctx *c)
{
    int i;
    if (!c) return;
    for (i = 0; i < c->n; i++) ocrt_monitor_destroy(&c->mon[i]);
    free(c->mon);
    free(c->noise);
    free(c);
}

int crtx_set_monitors(crtx_ctx *c, int first, int count, const crtx_monitor *m)
{
    int i;
    for (i = 0; i < count; i++) {
        ocrt_monitor *o = &c->mon[first + i];
        o->out = (unsigned char *) m[i].out;
        o->outw = m[i].outw; o->outh = m[i].outh; o->out_format = m[i].out_format;
        o->hue = m[i].hue; o->brightness = m[i].brightness; o->contrast = m[i].contrast; o->saturation = m[i].saturation;
        o->black_point = m[i].black_point; o->white_point = m[i].white_point;
        o->scanlines = m[i].scanlines; o->blend = m[i].blend; o->v_fac = m[i].v_fac;
        c->noise[first + i] = m[i].noise;
    }
    return 0;
}

int crtx_set_state(crtx_ctx *c, int first, int count, const crtx_state *s, void *stream)
{
    int i;
    (void) stream;
    for (i = 0; i < count; i++) {
        ocrt_monitor *o = &c->mon[first + i];
        {
            int r, x;
            for (r = 0; r < 3; r++)
                for (x = 0; x < 4; x++) o->ccf[r][x] = s[i].ccf[r][x];
        }
        o->hsync = s[i].hsync; o->vsync = s[i].vsync; o->rn = s[i].rn;
    }
    return 0;
}

int crtx_get_state(crtx_ctx *c, int first, int count, crtx_state *s, void *stream)
{
    int i;
    (void) stream;
    for (i = 0; i < count; i++) {
        const ocrt_monitor *o = &c->mon[first + i];
        {
            int r, x;
            for (r = 0; r < 3; r++)
                for (x = 0; x < 4; x++) s[i].ccf[r][x] = o->ccf[r][x];
        }
        s[i].hsync = o->hsync; s[i].vsync = o->vsync; s[i].rn = o->rn;
    }
    return 0;
}

int crtx_modulate(crtx_ctx *c, int first, int count, const crtx_source *src, void *stream)
{
    int i;
    (void) stream;
    for (i = 0; i < count; i++) {
        ocrt_rgb_source r;
        memset(&r, 0, sizeof(r));
        r.data = (const unsigned char *) src[i].data;
        r.format = src[i].format; r.w = src[i].w; r.h = src[i].h; r.raw = src[i].raw; r.as_color = src[i].as_color;
        r.field = src[i].field; r.frame = src[i].frame; r.hue = src[i].hue; r.xoffset = src[i].xoffset; r.yoffset = src[i].yoffset;
        ocrt_encode_rgb(c->sys, &c->mon[first + i], &r, NULL);
    }
    c->launches += 3;
    return 0;
}

int crtx_demodulate(crtx_ctx *c, int first, int count, void *stream)
{
    int i;
    (void) stream;
    for (i = 0; i < count; i++) ocrt_decode(c->sys, &c->mon[first + i], c->noise[first + i], NULL);
    c->launches += 3;
    return 0;
}

void *crtx_device_alloc(size_t bytes) { return calloc(bytes ? bytes : 1, 1); }
void crtx_device_free(void *p) { free(p); }
void *crtx_host_alloc(size_t bytes) { return malloc(bytes ? bytes : 1); }
void crtx_host_free(void *p) { free(p); }
int crtx_memcpy(void *dst, const void *src, size_t bytes, int kind, void *stream)
{
    (void) kind; (void) stream;
    memcpy(dst, src, bytes);
    return 0;
}
int crtx_memcmp_device(const void *a, const void *b, size_t bytes, int *differ, void *stream)
{
    (void) stream;
    *differ = memcmp(a, b, bytes) != 0;
    return 0;
}
int crtx_sync(void *stream) { (void) stream; return 0; }

int crtx_bmp_unpack(void *bgra, const void *file_pixels, int w, int h, int bits, void *stream)
{
    const int bytespp = bits / 8, rowbytes = (w * bytespp + 3) & ~3;
    int x, y;
    (void) stream;
    for (y = 0; y < h; y++) {
        const unsigned char *s = (const unsigned char *) file_pixels + (size_t) (h - 1 - y) * rowbytes;
        unsigned char *d = (unsigned char *) bgra + (size_t) y * w * 4;
        for (x = 0; x < w; x++) {
            d[4 * x + 0] = s[bytespp * x + 0];
            d[4 * x + 1] = s[bytespp * x + 1];
            d[4 * x + 2] = s[bytespp * x + 2];
            d[4 * x + 3] = bytespp == 4 ? s[4 * x + 3] : 255;
        }
    }
    return 0;
}

int crtx_bmp_pack(void *file_pixels, const void *bgra, int w, int h, void *stream)
{
    int y;
    (void) stream;
    for (y = 0; y < h; y++)
        memcpy((unsigned char *) file_pixels + (size_t) (h - 1 - y) * w * 4, (const unsigned char *) bgra + (size_t) y * w * 4, (size_t) w * 4);
    return 0;
}
