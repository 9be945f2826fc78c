// tmem_ld_bench.cu -- tcgen05.ld throughput by shape and warp count (developer probe): bytes per clk per SM.
#include <cstdio>
#include "tc5.cuh"
using namespace b200sr;
__device__ unsigned long long g_out[64];
__device__ unsigned g_sink;

#define LD_ASM(shape, num, N) \
    asm volatile("tcgen05.ld.sync.aligned." shape "." num ".b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];" \
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), \
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(addr))

// every variant returns 32 registers per thread = 32 lanes x 32 columns x 4 B = 4 KB per warp-instruction
__device__ __forceinline__ void ld_32x32b_x32(uint32_t addr, uint32_t (&r)[32]) { LD_ASM("32x32b", "x32", 32); }
__device__ __forceinline__ void ld_16x256b_x8(uint32_t addr, uint32_t (&r)[32]) { LD_ASM("16x256b", "x8", 32); }   // 16 lanes x 64 columns
__device__ __forceinline__ void ld_16x128b_x16(uint32_t addr, uint32_t (&r)[32]) { LD_ASM("16x128b", "x16", 32); } // 16 lanes x 64 columns
__device__ __forceinline__ void ld_16x64b_x32(uint32_t addr, uint32_t (&r)[32]) { LD_ASM("16x64b", "x32", 32); }   // 16 lanes x 64 columns

__global__ void __launch_bounds__(1024, 1) bench(int reps) {
    __shared__ uint32_t tptr;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tc5::tmem_alloc(smem_u32(&tptr), 512);
    tc5::fence_before_sync(); __syncthreads(); tc5::fence_after_sync();
    const uint32_t tmem = tptr;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    unsigned acc = 0;
    int slot = 0;
    for (int shape = 0; shape < 4; ++shape)
        for (int nw = 4; nw <= 32; nw *= 2)
            for (int batch = 1; batch <= 2; ++batch) {   // loads in flight per wait
                __syncthreads();
                const long long t0 = clock64();
                if (warp < nw) {
                    for (int r = 0; r < reps; ++r) {
                        uint32_t a[32], b[32];
                        const uint32_t col = (uint32_t)((r * 64 + (warp >> 2) * 32) & 255);
                        const uint32_t addr = tmem + lane_base + col;
                        if (shape == 0) { ld_32x32b_x32(addr, a); if (batch == 2) ld_32x32b_x32(addr + 32, b); }
                        else if (shape == 1) { ld_16x256b_x8(addr, a); if (batch == 2) ld_16x256b_x8(addr + (16u << 16), b); }
                        else if (shape == 2) { ld_16x128b_x16(addr, a); if (batch == 2) ld_16x128b_x16(addr + (16u << 16), b); }
                        else { ld_16x64b_x32(addr, a); if (batch == 2) ld_16x64b_x32(addr + (16u << 16), b); }
                        tc5::tmem_wait_ld();
                        acc += a[0] ^ a[13] ^ a[31];
                        if (batch == 2) acc += b[0] ^ b[17] ^ b[31];
                    }
                }
                const long long t1 = clock64();
                __syncthreads();
                if (tid == 0) g_out[slot] = (unsigned long long)(t1 - t0);
                ++slot;
            }
    if (acc == 0x12345678u) g_sink = acc;
    tc5::fence_before_sync(); __syncthreads();
    if (warp == 0) tc5::tmem_free(tmem, 512);
}
int main() {
    const int reps = 256;
    bench<<<1, 1024>>>(reps);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    unsigned long long out[64]; cudaMemcpyFromSymbol(out, g_out, sizeof out);
    const char *names[] = {"32x32b.x32", "16x256b.x8", "16x128b.x16", "16x64b.x32"};
    int slot = 0;
    for (int shape = 0; shape < 4; ++shape)
        for (int nw = 4; nw <= 32; nw *= 2)
            for (int batch = 1; batch <= 2; ++batch, ++slot) {
                const double bytes = (double)reps * nw * batch * 4096.0;
                printf("%-12s %2d warps, %d x 4 KB per wait: %7.1f B/clk/SM   (%.0f clk per load+wait)\n", names[shape], nw, batch, bytes / (double)out[slot], (double)out[slot] / reps);
            }
    return 0;
}
