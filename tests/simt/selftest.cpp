// tests/simt/selftest.cpp -- TEST INFRASTRUCTURE: the interpreter checked against the documented semantics of the CUDA
// features it stands in for (CUDA C++ Programming Guide / Math API: warp shuffles, ballots, __byte_perm, the SIMD
// video intrinsics, __vimin_s32_relu; PTX ISA: mbarrier phase / tx-count rules), with known answers.  A wrong
// emulation here could make a kernel test pass for the wrong reason.
#include <cuda_runtime.h>

#include "crt_ptx.cuh"

static int g_fail = 0;
#define CHECK(cond)                                                        \
    do {                                                                   \
        if (!(cond)) {                                                     \
            printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond);       \
            g_fail = 1;                                                    \
        }                                                                  \
    } while (0)

static unsigned g_out[8][64];
namespace crt { extern unsigned char smem_raw[]; } // the dynamic shared-memory window (build.py rewrites "extern __shared__")

__global__ void k_collectives(int dummy)
{
    (void) dummy;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int v = 100 * warp + lane;
    g_out[0][threadIdx.x] = (unsigned) __shfl_sync(0xffffffffu, v, 5);          // every lane reads lane 5
    g_out[1][threadIdx.x] = (unsigned) __shfl_up_sync(0xffffffffu, v, 3);       // lanes 0..2 keep their own value
    g_out[2][threadIdx.x] = (unsigned) __shfl_xor_sync(0xffffffffu, v, 16);
    g_out[3][threadIdx.x] = __ballot_sync(0xffffffffu, (lane % 3) == 0);
    g_out[4][threadIdx.x] = (unsigned) __shfl_sync(0xffffffffu, v, lane + 1, 8); // width 8: wraps inside the octet
    __shared__ int total;
    if (threadIdx.x == 0) total = 0;
    __syncthreads();
    atomicAdd(&total, 1);
    const int any = __syncthreads_or(threadIdx.x == 37);
    g_out[5][threadIdx.x] = (unsigned) total;
    g_out[6][threadIdx.x] = (unsigned) any;
}

__global__ void k_mbarrier(unsigned char *src)
{
    unsigned char *smem_raw = crt::smem_raw;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw + 256);
    const int lane = threadIdx.x;
    if (lane == 0) {
        crt::mbar_init(bar, 1);
        crt::mbar_fence_init();
    }
    __syncwarp();
    for (int round = 0; round < 3; round++) { // phase parity alternates 0, 1, 0
        if (lane == 0) crt::mbar_expect_tx(bar, 32u * 16u);
        __syncwarp();
        crt::tma_load_1d(smem_raw + 512 + 16 * lane, src + 512 * round + 16 * lane, 16, bar);
        // the copy is performed at the latest legal moment: not before somebody waits
        if (round == 0 && lane == 31) g_out[7][0] = smem_raw[512]; // still the 0xa5 fill
        crt::mbar_wait(bar, (unsigned) (round & 1));
        g_out[7][1 + round] = smem_raw[512 + 16 * ((lane + 1) & 31)]; // another lane's bytes have landed
        __syncwarp();
    }
}

int main()
{
    // ---- integer intrinsics, known answers
    CHECK(__byte_perm(0x33221100u, 0x77665544u, 0x5410u) == 0x55441100u);
    CHECK(__byte_perm(0x33221100u, 0x77665544u, 0x7632u) == 0x77663322u);
    // (selector nibbles above 7 are not used by any kernel and not asserted here)
    CHECK(__vabsss4(0x80ff7f01u) == 0x7f017f01u);                 // |-128| saturates to 127
    CHECK(__vmaxs4(0x80017f00u, 0x81818181u) == 0x81017f00u);     // signed per byte
    CHECK(__vmaxu4(0x80017f00u, 0x81818181u) == 0x81818181u);     // unsigned per byte
    CHECK(__vimin_s32_relu(-5, 10) == 0 && __vimin_s32_relu(20, 10) == 10 && __vimin_s32_relu(7, 10) == 7);
    CHECK(__popc(0xf0f00001u) == 9 && __ffs(0) == 0 && __ffs(0x8) == 4);
    CHECK(min(-1, 2u) == 2u);                                     // CUDA's mixed overload compares as unsigned

    // ---- warp collectives and block barriers
    ::simt::launch(dim3(1), dim3(64), 0, [&]() { k_collectives(0); });
    for (int t = 0; t < 64; t++) {
        const int lane = t & 31, warp = t >> 5;
        CHECK(g_out[0][t] == (unsigned) (100 * warp + 5));
        CHECK(g_out[1][t] == (unsigned) (100 * warp + (lane < 3 ? lane : lane - 3)));
        CHECK(g_out[2][t] == (unsigned) (100 * warp + (lane ^ 16)));
        CHECK(g_out[3][t] == 0x49249249u);
        CHECK(g_out[4][t] == (unsigned) (100 * warp + ((lane & ~7) | ((lane + 1) & 7))));
        CHECK(g_out[5][t] == 64u && g_out[6][t] == 1u);
    }

    // ---- mbarrier + bulk copy: three phases on one barrier, deferred copies
    static unsigned char src[2048] __attribute__((aligned(16)));
    for (int i = 0; i < 2048; i++) src[i] = (unsigned char) (i / 512 + 1);
    ::simt::launch(dim3(1), dim3(32), 1024, [&]() { k_mbarrier(src); });
    CHECK(g_out[7][0] == 0xa5u);
    CHECK(g_out[7][1] == 1u && g_out[7][2] == 2u && g_out[7][3] == 3u);

    printf(g_fail ? "selftest FAILED\n" : "selftest ok\n");
    return g_fail;
}
