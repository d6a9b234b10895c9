/* tools/crtx_still.c -- batch counterpart of the reference's command-line driver (crt_main.c:150-283), C89,
 * over include/crtx_batch.h alone (no CUDA header).
 *
 *   crtx_still -m|o|f|p|r|a|h outwidth outheight noise artifact_hue in1 out1 [in2 out2 ...]
 *
 * Same flags, arguments, file formats (.ppm, anything else is read / written as BMP) and pixels as crt_main.c, but any
 * number of images at once: every (in, out) pair is one monitor of a crtx context, and the driver's "accumulate 4
 * frames" loop (crt_main.c:241-255: 4 field pairs interlaced, 4 fields progressive; blend 1, scanlines 1) advances
 * all of them together, one crtx_modulate + crtx_demodulate per step for the whole batch.  The files' pixel data
 * cross PCIe as stored and are (un)packed on the device (crtx_ppm_unpack / crtx_bmp_unpack, crtx_ppm_pack /
 * crtx_bmp_pack).  `o` is accepted and ignored (this program never prompts); `a` writes the analog signal
 * as a grey image like crt_main.c:257-268.  The output of every pair is byte-identical to what the reference
 * driver writes for it (tests/test_gpu_still_cli.py).
 */
#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "crtx_batch.h"

#define PIX_BGRA 5 /* CRT_PIX_FORMAT_BGRA: the loaders' int pixels 0x00RRGGBB in memory (crt_main.c:225-226) */

static void
die(const char *what)
{
    fprintf(stderr, "crtx_still: %s (%s)\n", what, crtx_last_error());
    exit(EXIT_FAILURE);
}

#define TRY(call) do { if ((call) != 0) die(#call); } while (0)

static int
has_suffix(const char *s, const char *suf) /* crt_main.c:29-32 */
{
    size_t n = strlen(s), m = strlen(suf);
    return n >= m && strcmp(s + n - m, suf) == 0;
}

static int
to_int(const char *s, int *err) /* crt_main.c:47-65 */
{
    char *tail;
    long val;
    errno = 0;
    val = strtol(s, &tail, 10);
    if (errno != 0 || *tail != '\0') {
        fprintf(stderr, "crtx_still: bad integer '%s'\n", s);
        *err = 1;
    }
    return (int) val;
}

static unsigned
le32(const unsigned char *p)
{
    return (unsigned) p[0] | ((unsigned) p[1] << 8) | ((unsigned) p[2] << 16) | ((unsigned) p[3] << 24);
}

/* one input image: the file's pixel data as stored, and how to unpack it */
struct image {
    unsigned char *bytes; /* page-locked */
    size_t nbytes;
    int w, h;
    int is_ppm, maxc, bits;
};

/* P6 header the way ppm_rw.c:36-70 reads it: three fgets lines ("P6", "w h", "max"), '#' lines skipped */
static int
load_ppm(const char *file, struct image *im)
{
    FILE *f = fopen(file, "rb");
    char buf[64];
    int header = 0;
    if (f == NULL) return 0;
    im->maxc = 0xff;
    while (header < 3) {
        if (!fgets(buf, sizeof(buf), f)) goto bad;
        if (buf[0] == '#') continue;
        if (header == 0 && (buf[0] != 'P' || buf[1] != '6')) goto bad;
        if (header == 1 && sscanf(buf, "%d %d", &im->w, &im->h) != 2) goto bad;
        if (header == 2) {
            im->maxc = atoi(buf);
            if (im->maxc > 0xff || im->maxc < 1) goto bad;
        }
        header++;
    }
    if (im->w <= 0 || im->h <= 0) goto bad;
    im->nbytes = (size_t) im->w * (size_t) im->h * 3;
    im->bytes = (unsigned char *) crtx_host_alloc(im->nbytes);
    if (im->bytes == NULL || fread(im->bytes, 1, im->nbytes, f) != im->nbytes) goto bad; /* ppm_rw.c:84-87: early eof */
    fclose(f);
    im->is_ppm = 1;
    return 1;
bad:
    fclose(f);
    return 0;
}

/* the BMP subset bmp_rw.c:22-94 reads: 54-byte header, 24 or 32 bits, bottom-up rows padded to 4 bytes */
static int
load_bmp(const char *file, struct image *im)
{
    FILE *f = fopen(file, "rb");
    unsigned char header[54];
    unsigned bytespp, rowbytes;
    if (f == NULL) return 0;
    if (fread(header, 1, 54, f) != 54) goto bad;
    im->w = (int) le32(header + 18);
    im->h = (int) le32(header + 22);
    bytespp = (le32(header + 28) & 0xff) / 8;
    if ((bytespp != 3 && bytespp != 4) || im->w <= 0 || im->h <= 0) goto bad;
    rowbytes = ((unsigned) im->w * bytespp + 3u) & ~3u;
    im->nbytes = (size_t) rowbytes * (size_t) im->h;
    im->bytes = (unsigned char *) crtx_host_alloc(im->nbytes);
    if (im->bytes == NULL || fread(im->bytes, 1, im->nbytes, f) != im->nbytes) goto bad;
    fclose(f);
    im->is_ppm = 0;
    im->bits = (int) bytespp * 8;
    return 1;
bad:
    fclose(f);
    return 0;
}

static int
save_ppm(const char *file, const unsigned char *rgb, int w, int h) /* ppm_rw.c:96-122 */
{
    FILE *f = fopen(file, "wb");
    if (f == NULL) return 0;
    fprintf(f, "P6\n%d %d\n255\n", w, h);
    fwrite(rgb, 3, (size_t) w * (size_t) h, f);
    fclose(f);
    return 1;
}

static int
save_bmp(const char *file, const unsigned char *file_pixels, int w, int h) /* the header bmp_rw.c:96-146 writes */
{
    FILE *f;
    unsigned char head[54];
    unsigned filesize = 14 + 40 + (unsigned) w * (unsigned) h * 4;
    int k;
    memset(head, 0, sizeof(head));
    head[0] = 'B';
    head[1] = 'M';
    for (k = 0; k < 4; k++) {
        head[2 + k] = (unsigned char) (filesize >> (8 * k));
        head[18 + k] = (unsigned char) ((unsigned) w >> (8 * k));
        head[22 + k] = (unsigned char) ((unsigned) h >> (8 * k));
    }
    head[10] = 14 + 40;
    head[14] = 40;
    head[26] = 1;
    head[28] = 32;
    f = fopen(file, "wb");
    if (f == NULL) return 0;
    fwrite(head, 1, 54, f);
    fwrite(file_pixels, 4, (size_t) w * (size_t) h, f);
    fclose(f);
    return 1;
}

static void
usage(const char *p)
{
    printf("usage: %s -m|o|f|p|r|a|h outwidth outheight noise artifact_hue in1 out1 [in2 out2 ...]\n", p);
    printf("\tthe flags, arguments and files of the reference's ntsc command (crt_main.c), any number of images at once\n");
}

int
main(int argc, char **argv)
{
    int docolor = 1, field = 0, progressive = 0, raw = 0, save_analog = 0;
    int outw, outh, noise, hue, err = 0, n, i, step, nsteps;
    const char *flags;
    struct image *img;
    void **d_src, **d_out;
    crtx_ctx *ctx;
    crtx_monitor *mon;
    crtx_source *src;
    size_t out_file_bytes;

    if (argc < 8 || ((argc - 6) & 1)) {
        usage(argv[0]);
        return EXIT_FAILURE;
    }
    flags = argv[1];
    if (*flags == '-') flags++;
    for (; *flags != '\0'; flags++) { /* crt_main.c:88-112 */
        switch (*flags) {
            case 'm': docolor = 0; break;
            case 'o': break;
            case 'f': field = 1; break;
            case 'p': progressive = 1; break;
            case 'r': raw = 1; break;
            case 'a': save_analog = 1; break;
            case 'h': usage(argv[0]); return EXIT_FAILURE;
            default: fprintf(stderr, "Unrecognized flag '%c'\n", *flags); return EXIT_FAILURE;
        }
    }
    outw = to_int(argv[2], &err);
    outh = to_int(argv[3], &err);
    noise = to_int(argv[4], &err);
    hue = to_int(argv[5], &err);
    if (err || outw <= 0 || outh <= 0) return EXIT_FAILURE;
    if (noise < 0) noise = 0; /* crt_main.c:183 */
    hue %= 360;               /* crt_main.c:189 */
    n = (argc - 6) / 2;

    img = (struct image *) calloc((size_t) n, sizeof(*img));
    d_src = (void **) calloc((size_t) n, sizeof(*d_src));
    d_out = (void **) calloc((size_t) n, sizeof(*d_out));
    mon = (crtx_monitor *) calloc((size_t) n, sizeof(*mon));
    src = (crtx_source *) calloc((size_t) n, sizeof(*src));
    if (!img || !d_src || !d_out || !mon || !src) die("out of memory");
    TRY(crtx_create(&ctx, n));

    /* ---- load: file bytes -> device -> int pixels (crt_main.c:204-214) */
    for (i = 0; i < n; i++) {
        const char *file = argv[6 + 2 * i];
        void *d_file;
        if (!(has_suffix(file, ".ppm") ? load_ppm(file, &img[i]) : load_bmp(file, &img[i]))) {
            fprintf(stderr, "crtx_still: unable to read image %s\n", file);
            return EXIT_FAILURE;
        }
        d_file = crtx_device_alloc(img[i].nbytes + 16);
        d_src[i] = crtx_device_alloc((size_t) img[i].w * (size_t) img[i].h * 4 + 16);
        d_out[i] = crtx_device_alloc((size_t) outw * (size_t) outh * 4); /* zero-filled, like calloc at crt_main.c:195 */
        if (!d_file || !d_src[i] || !d_out[i]) die("device memory");
        TRY(crtx_memcpy(d_file, img[i].bytes, img[i].nbytes, 0, NULL));
        if (img[i].is_ppm) TRY(crtx_ppm_unpack(d_src[i], d_file, img[i].w, img[i].h, img[i].maxc, NULL));
        else TRY(crtx_bmp_unpack(d_src[i], d_file, img[i].w, img[i].h, img[i].bits, NULL));
        TRY(crtx_sync(NULL));
        crtx_device_free(d_file);
        crtx_host_free(img[i].bytes);
        img[i].bytes = NULL;

        mon[i].out = d_out[i]; /* crt_init + crt_reset defaults (crt_core.c:250-289), then crt_main.c:235-236 */
        mon[i].outw = outw;
        mon[i].outh = outh;
        mon[i].out_format = PIX_BGRA;
        mon[i].hue = 0;
        mon[i].brightness = 0;
        mon[i].contrast = 180;
        mon[i].saturation = 10;
        mon[i].black_point = 0;
        mon[i].white_point = 100;
        mon[i].scanlines = 1;
        mon[i].blend = 1;
        mon[i].v_fac = 0;
        mon[i].noise = noise;

        src[i].data = d_src[i]; /* crt_main.c:223-233 */
        src[i].format = PIX_BGRA;
        src[i].w = img[i].w;
        src[i].h = img[i].h;
        src[i].raw = raw;
        src[i].as_color = docolor;
        src[i].field = field & 1;
        src[i].frame = 0;
        src[i].hue = hue;
        src[i].reinit = 1; /* NES-family libraries: the first call of a stream writes the sync template */
        printf("loaded %d %d\n", img[i].w, img[i].h);
    }
    TRY(crtx_set_monitors(ctx, 0, n, mon));

    /* ---- accumulate 4 frames (crt_main.c:241-255), every image one field per step */
    nsteps = progressive ? 4 : 8;
    for (step = 0; step < nsteps; step++) {
        TRY(crtx_modulate(ctx, 0, n, src, NULL));
        TRY(crtx_demodulate(ctx, 0, n, NULL));
        for (i = 0; i < n; i++) {
            src[i].reinit = 0;
            if (!progressive) { /* crt_main.c:245-252: an iteration is two calls with one field flip between them ... */
                if ((step & 1) == 0) src[i].field ^= 1;
                /* ... and the frame flips after the second call of iterations 0 and 2 */
                else if (((step >> 1) & 1) == 0) src[i].frame ^= 1;
            }
        }
    }

    /* ---- store (crt_main.c:257-281) */
    out_file_bytes = (size_t) outw * (size_t) outh * 4;
    for (i = 0; i < n; i++) {
        const char *file = argv[7 + 2 * i];
        int w = outw, h = outh;
        void *d_img = d_out[i], *d_file;
        unsigned char *host;
        size_t nbytes;
        if (save_analog) { /* analog[i] + 128 as a grey level, CRT_HRES x CRT_VRES */
            int hres = crtx_hres(), total = crtx_input_size(), k;
            signed char *sig = (signed char *) malloc((size_t) total);
            unsigned *grey = (unsigned *) crtx_host_alloc((size_t) total * 4);
            if (!sig || !grey) die("out of memory");
            TRY(crtx_read_signal(ctx, i, 0, sig, NULL));
            for (k = 0; k < total; k++) {
                unsigned norm = (unsigned) (sig[k] + 128);
                grey[k] = norm << 16 | norm << 8 | norm;
            }
            w = hres;
            h = total / hres;
            d_img = crtx_device_alloc((size_t) total * 4);
            if (!d_img) die("device memory");
            TRY(crtx_memcpy(d_img, grey, (size_t) total * 4, 0, NULL));
            TRY(crtx_sync(NULL));
            crtx_host_free(grey);
            free(sig);
        }
        nbytes = has_suffix(file, ".ppm") ? (size_t) w * (size_t) h * 3 : (size_t) w * (size_t) h * 4;
        d_file = crtx_device_alloc(nbytes + 16);
        host = (unsigned char *) crtx_host_alloc(nbytes);
        if (!d_file || !host) die("memory");
        if (has_suffix(file, ".ppm")) TRY(crtx_ppm_pack(d_file, d_img, w, h, NULL));
        else TRY(crtx_bmp_pack(d_file, d_img, w, h, NULL));
        TRY(crtx_memcpy(host, d_file, nbytes, 1, NULL));
        TRY(crtx_sync(NULL));
        if (!(has_suffix(file, ".ppm") ? save_ppm(file, host, w, h) : save_bmp(file, host, w, h))) {
            fprintf(stderr, "crtx_still: unable to write image %s\n", file);
            return EXIT_FAILURE;
        }
        crtx_host_free(host);
        crtx_device_free(d_file);
        if (save_analog) crtx_device_free(d_img);
        crtx_device_free(d_src[i]);
        crtx_device_free(d_out[i]);
    }
    (void) out_file_bytes;
    crtx_destroy(ctx);
    free(img);
    free(d_src);
    free(d_out);
    free(mon);
    free(src);
    printf("done\n");
    return EXIT_SUCCESS;
}
