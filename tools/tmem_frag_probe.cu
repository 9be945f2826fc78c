// tmem_frag_probe.cu -- developer probe: which (TMEM lane, column) lands in which register of tcgen05.ld.16x256b (x2), and which
// lanes a warp reaches with a +16 lane offset.  Written with the 32x32b shape (thread = lane, register = column).
#include <cstdio>
#include "tc5.cuh"
using namespace b200sr;
__device__ float g_out[4][2][32][8];
__global__ void __launch_bounds__(128, 1) probe() {
    __shared__ __align__(16) uint32_t s[8];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) tc5::tmem_alloc(smem_u32(s), 32);
    tc5::fence_before_sync(); __syncthreads(); tc5::fence_after_sync();
    const uint32_t tmem = s[0];
    uint32_t v[16];
    for (int c = 0; c < 16; ++c) v[c] = __float_as_uint((float)(tid * 100 + c));
    tc5::tmem_st16(tmem + ((uint32_t)(warp * 32) << 16), v);
    tc5::tmem_wait_st();
    tc5::fence_before_sync(); __syncthreads(); tc5::fence_after_sync();
    for (int h = 0; h < 2; ++h) {
        uint32_t r[8];
        const uint32_t a = tmem + ((uint32_t)(warp * 32 + 16 * h) << 16);
        asm volatile(
#ifdef SHAPE128
                     "tcgen05.ld.sync.aligned.16x128b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
#else
                     "tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
#endif

                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(a));
        tc5::tmem_wait_ld();
        for (int i = 0; i < 8; ++i) g_out[warp][h][lane][i] = __uint_as_float(r[i]);
    }
    tc5::fence_before_sync(); __syncthreads();
    if (warp == 0) tc5::tmem_free(tmem, 32);
}
int main() {
    probe<<<1, 128>>>();
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    static float out[4][2][32][8];
    cudaMemcpyFromSymbol(out, g_out, sizeof out);
    for (int w = 0; w < 4; w += 3)
        for (int h = 0; h < 2; ++h)
            for (int l = 0; l < 32; l += (l < 8 ? 1 : 8)) {
                printf("warp %d half %d thread %2d:", w, h, l);
                for (int i = 0; i < 8; ++i) printf("  (lane %3d, col %2d)", (int)out[w][h][l][i] / 100, (int)out[w][h][l][i] % 100);
                printf("\n");
            }
    // check the expected mapping everywhere
    int bad = 0;
    for (int w = 0; w < 4; ++w) for (int h = 0; h < 2; ++h) for (int l = 0; l < 32; ++l) for (int i = 0; i < 8; ++i) {
        const int g = l / 4, j = l % 4;
#ifdef SHAPE128
        const int row = w * 32 + 16 * h + g + (i & 1) * 8, col = 4 * (i >> 1) + j;
#else
        const int row = w * 32 + 16 * h + g + ((i >> 1) & 1) * 8, col = (i >> 2) * 8 + 2 * j + (i & 1);
#endif
        if ((int)out[w][h][l][i] != row * 100 + col) ++bad;
    }
    printf("expected mapping (reg i of thread (g,j): lane base+g+8*((i>>1)&1), col 8*(i>>2)+2j+(i&1)): %s (%d mismatches)\n", bad ? "NO" : "YES", bad);
    return 0;
}
