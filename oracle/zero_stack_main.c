/* TEST INFRASTRUCTURE (oracle/): a main() in front of the reference's UNMODIFIED extra/video_convert.c.
 *
 * video_convert.c:153 declares `struct NTSC_SETTINGS ntsc;` on main's stack and never clears it (the header asks for a
 * memset, crt_ntsc.h:122): xoffset, yoffset and iirs_initialized are whatever the stack held.  To compare the reference
 * build of that driver with the build linked against our library, both must see the same "whatever": the driver source is
 * compiled as it lies under /root/reference with -Dmain=video_convert_main on the command line (no edit), and this main
 * clears the stack region the driver's frame is about to occupy before calling it.  Used by both builds
 * (oracle/Makefile: video_ref_ntsc_z, video_b200_ntsc_z), never by the product. */
#include <string.h>

int video_convert_main(int argc, char **argv);

#if defined(__GNUC__)
#define NOINLINE __attribute__((noinline))
#else
#define NOINLINE
#endif

static NOINLINE void clear_stack_below(void)
{
    volatile char region[3 * 1024 * 1024]; /* struct CRT alone is 466 KB; the default stack limit is 8 MB */
    memset((void *) region, 0, sizeof(region));
#if defined(__GNUC__)
    __asm__ volatile("" : : "r"(region) : "memory");
#endif
}

int main(int argc, char **argv)
{
    clear_stack_below();
    return video_convert_main(argc, argv);
}
