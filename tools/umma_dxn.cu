// umma_dxn.cu -- tensor-queue cost of the block's 3x3 in the shipped mapping (12 SS MMAs of N = 32 per 128 outputs, 4 M-tiles per tile)
// against the "horizontal taps in N" mapping proposed in DESIGN.md 9 (5 SS MMAs of N = 80 per 128 t2 pixels, 5 M-tiles per tile: K = 3 rows x
// 3 chunks = 9 chunks, one dummy half), issued alone by one thread (developer probe; operands are zeros, SWIZZLE_NONE descriptors).
#include <cstdio>
#include "tc5.cuh"
using namespace b200sr;
__device__ unsigned long long g_out[16];

__global__ void __launch_bounds__(128, 1) bench(int reps) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *ctrl = smem, *a = smem + 256, *b = a + 96 * 1024;
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t bar = smem_u32(ctrl);
    if (tid == 0) { tc5::mbar_init(bar, 1); tc5::mbar_init_fence(); }
    __syncwarp();
    if (warp == 0) tc5::tmem_alloc(smem_u32(ctrl + 16), 512);
    for (int i = tid; i < (96 + 48) * 1024 / 16; i += 128) *reinterpret_cast<uint4 *>(a + i * 16) = make_uint4(0, 0, 0, 0);
    tc5::fence_proxy_async(); tc5::fence_before_sync(); __syncthreads(); tc5::fence_after_sync();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t *>(ctrl + 16);
    const uint32_t a_u = smem_u32(a), b_u = smem_u32(b);
    uint32_t phase = 0;
    const int NS[4] = {32, 80, 96, 144};
    for (int mode = 0; mode < 4; ++mode) {
        const int N = NS[mode], nm = mode == 0 ? 12 : 5, nt = mode == 0 ? 4 : 5;
        const uint32_t idesc = tc5::idesc_bf16_f32(128, N);
        const uint64_t bw = tc5::smem_desc(b_u, 128, 16 * 128), a0 = tc5::smem_desc(a_u, 384, 128);   // chunk pairs 384 B apart (the t2 group)
        for (int rep = 0; rep < 2; ++rep) {
            long long t0 = 0, t1 = 0;
            if (warp == 0) {
                tc5::fence_after_sync();
                t0 = clock64();
                if (tc5::elect_one()) {
                    for (int r = 0; r < reps; ++r)
                        for (int m = 0; m < nt; ++m)
#pragma unroll 4
                            for (int i = 0; i < nm; ++i)
                                tc5::mma_ss(tmem + (m & 1) * 160, a0 + (uint64_t)((m * 6144 + i * 1536) >> 4), bw + (uint64_t)(16 * i), idesc, i > 0);
                    tc5::commit(bar);
                }
                t1 = clock64();
                __syncwarp();
            }
            if (warp != 0 || tid == 0) tc5::mbar_wait(bar, phase);
            phase ^= 1;
            if (tid == 0 && rep == 1) { g_out[2 * mode] = (unsigned long long)(t1 - t0); g_out[2 * mode + 1] = (unsigned long long)(clock64() - t0); }
            tc5::fence_before_sync();
            __syncthreads();
        }
    }
    if (warp == 0) tc5::tmem_free(tmem, 512);
}

int main() {
    const int reps = 8;
    size_t smem = 256 + (96 + 48) * 1024;
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    bench<<<1, 128, smem>>>(reps);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    unsigned long long out[16]; cudaMemcpyFromSymbol(out, g_out, sizeof out);
    const char *names[] = {"shipped: 4 M-tiles x 12 SS MMAs, N = 32", "dx in N: 5 M-tiles x 5 SS MMAs, N = 80", "dx in N: 5 x 5, N = 96 (32-wide blocks)",
                           "tail-like: 5 x 5, N = 144"};
    const int cnt[] = {48, 25, 25, 25};
    for (int i = 0; i < 4; ++i)
        printf("%-46s per tile: issue %6.0f  complete %6.0f clk   (%.1f clk/instr)\n", names[i], (double)out[2 * i] / reps, (double)out[2 * i + 1] / reps,
               (double)out[2 * i + 1] / reps / cnt[i]);
    return 0;
}
