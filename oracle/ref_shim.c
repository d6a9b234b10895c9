/* oracle/ref_shim.c -- TEST INFRASTRUCTURE, not product code.
 *
 * Compiled together with the UNMODIFIED reference sources (where they lie under
 * /root/reference) into oracle/_ref/libref_<variant>.so.  It only adds a few
 * introspection entry points so the Python test-suite can lay a ctypes view over
 * the reference's compile-time-polymorphic `struct CRT` / `struct NTSC_SETTINGS`
 * (crt_core.h:74-92, crt_ntsc.h:111-124, crt_nes.h:132-143, crt_ntscvhs.h:133-147)
 * without hard-coding offsets.
 */
#include <stddef.h>
#include <stdlib.h>
#include "crt_core.h"

#define OFF(f) ((int) offsetof(struct CRT, f))

int ref_system(void) { return CRT_SYSTEM; }
#ifdef CRT_CHROMA_PATTERN
int ref_chroma_pattern(void) { return CRT_CHROMA_PATTERN; }
#else /* crt_snes.h has no such switch */
int ref_chroma_pattern(void) { return -1; }
#endif
int ref_sizeof_crt(void) { return (int) sizeof(struct CRT); }
int ref_sizeof_settings(void) { return (int) sizeof(struct NTSC_SETTINGS); }

/* geometry, in the order tests/refbind.py expects */
void
ref_geometry(int *g)
{
    g[0] = CRT_HRES;        g[1] = CRT_VRES;       g[2] = CRT_INPUT_SIZE;
    g[3] = CRT_TOP;         g[4] = CRT_BOT;        g[5] = CRT_CC_VPER;
    g[6] = CRT_CC_SAMPLES;  g[7] = SYNC_BEG;       g[8] = BW_BEG;
    g[9] = CB_BEG;          g[10] = AV_BEG;        g[11] = AV_LEN;
    g[12] = CRT_HSYNC_WINDOW; g[13] = CRT_VSYNC_WINDOW;
    g[14] = WHITE_LEVEL;    g[15] = BURST_LEVEL;   g[16] = BLACK_LEVEL;
    g[17] = BLANK_LEVEL;    g[18] = SYNC_LEVEL;    g[19] = CB_CYCLES;
}

void
ref_crt_offsets(int *o)
{
    o[0] = OFF(analog);     o[1] = OFF(inp);       o[2] = OFF(outw);
    o[3] = OFF(outh);       o[4] = OFF(out_format); o[5] = OFF(out);
    o[6] = OFF(hue);        o[7] = OFF(brightness); o[8] = OFF(contrast);
    o[9] = OFF(saturation); o[10] = OFF(black_point); o[11] = OFF(white_point);
    o[12] = OFF(scanlines); o[13] = OFF(blend);    o[14] = OFF(v_fac);
    o[15] = OFF(ccf);       o[16] = OFF(hsync);    o[17] = OFF(vsync);
    o[18] = OFF(rn);
}

/* libc PRNG control for the VHS variant (crt_core.c:344-351, crt_ntscvhs.c:206) */
void ref_srand(unsigned seed) { srand(seed); }
int ref_rand(void) { return rand(); }
