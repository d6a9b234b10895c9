// TEST INFRASTRUCTURE: the byte-parallel helpers of crt_sync.cuh (clamp127_4, abs127_4, max127_4) against their definitions,
// byte by byte: every pair of byte values in every pair of positions, and a million random words.  Compiled by
// tests/test_simt_kernels.py::test_packed_byte_helpers against the interpreter's headers (the functions are plain integer C++).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#include "crt_kernels.cuh" // (pulls in crt_sync.cuh, in the order the product uses)

static unsigned ref_clamp(unsigned x)
{
    unsigned r = 0;
    for (int i = 0; i < 4; i++) {
        int b = (signed char) (x >> (8 * i));
        if (b < -127) b = -127;
        r |= ((unsigned) b & 0xffu) << (8 * i);
    }
    return r;
}
static unsigned ref_abs(unsigned x)
{
    unsigned r = 0;
    for (int i = 0; i < 4; i++) {
        int b = (signed char) (x >> (8 * i));
        r |= ((unsigned) (b < 0 ? -b : b) & 0xffu) << (8 * i);
    }
    return r;
}
static unsigned ref_max(unsigned a, unsigned b)
{
    unsigned r = 0;
    for (int i = 0; i < 4; i++) {
        const unsigned x = (a >> (8 * i)) & 0xffu, y = (b >> (8 * i)) & 0xffu;
        r |= (x > y ? x : y) << (8 * i);
    }
    return r;
}

int main()
{
    long bad = 0;
    unsigned lcg = 12345u;
    for (int p = 0; p < 4; p++)
        for (int q = 0; q < 4; q++)
            for (unsigned u = 0; u < 256; u++)
                for (unsigned v = 0; v < 256; v++) {
                    lcg = lcg * 1664525u + 1013904223u;
                    unsigned w = lcg;
                    w = (w & ~(0xffu << (8 * p))) | (u << (8 * p));
                    w = (w & ~(0xffu << (8 * q))) | (v << (8 * q));
                    if (crt::clamp127_4(w) != ref_clamp(w)) bad++;
                    const unsigned c = ref_clamp(w); // (the other two take samples in -127 .. 127 / values in 0 .. 127)
                    if (crt::abs127_4(c) != ref_abs(c)) bad++;
                    const unsigned a = ref_abs(c), b = ref_abs(ref_clamp(lcg * 2654435761u));
                    if (crt::max127_4(a, b) != ref_max(a, b)) bad++;
                }
    for (int i = 0; i < 1000000; i++) {
        lcg = lcg * 1664525u + 1013904223u;
        const unsigned w = lcg ^ (lcg >> 13);
        if (crt::clamp127_4(w) != ref_clamp(w)) bad++;
        const unsigned c = ref_clamp(w);
        if (crt::abs127_4(c) != ref_abs(c)) bad++;
    }
    if (bad) {
        printf("helpers: %ld mismatches\n", bad);
        return 1;
    }
    printf("helpers ok\n");
    return 0;
}
