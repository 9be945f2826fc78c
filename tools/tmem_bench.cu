// tmem_bench.cu -- tcgen05.ld / tcgen05.st throughput per SM (developer probe).
#include <cstdio>
#include "tc5.cuh"
using namespace b200sr;
__device__ unsigned long long g_out[32];
__device__ unsigned g_sink;

template <int MODE>
__device__ unsigned run(uint32_t tmem, int reps) {
    const int warp = threadIdx.x >> 5;
    const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    unsigned acc = 0;
    for (int r = 0; r < reps; ++r) {
        if (MODE == 0) {         // ld x32, wait each
            uint32_t v[32]; tc5::tmem_ld32(base + (r & 7) * 32, v); tc5::tmem_wait_ld(); acc += v[0] ^ v[31];
        } else if (MODE == 1) {  // 4 x ld x32 in flight
            uint32_t a[32], b[32], c[32], d[32];
            tc5::tmem_ld32(base, a); tc5::tmem_ld32(base + 32, b); tc5::tmem_ld32(base + 64, c); tc5::tmem_ld32(base + 96, d);
            tc5::tmem_wait_ld(); acc += a[0] ^ b[1] ^ c[2] ^ d[3];
        } else if (MODE == 2) {  // st x16
            uint32_t p[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) p[j] = r + j;
            tc5::tmem_st16(base + (r & 7) * 16, p); tc5::tmem_wait_st();
        } else {                 // ld x32 .pack::16b
            uint32_t v[16];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.pack::16b.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(base + (r & 7) * 32));
            tc5::tmem_wait_ld(); acc += v[0] ^ v[15];
        }
    }
    return acc;
}

__global__ void __launch_bounds__(512, 1) bench(int reps) {
    __shared__ uint32_t tptr;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tc5::tmem_alloc(smem_u32(&tptr), 512);
    tc5::fence_before_sync(); __syncthreads(); tc5::fence_after_sync();
    const uint32_t tmem = tptr;
    int slot = 0;
    for (int nw = 4; nw <= 16; nw *= 2)
        for (int mode = 0; mode < 4; ++mode) {
            __syncthreads();
            const long long t0 = clock64();
            unsigned a = 0;
            if (warp < nw) a = mode == 0 ? run<0>(tmem, reps) : mode == 1 ? run<1>(tmem, reps) : mode == 2 ? run<2>(tmem, reps) : run<3>(tmem, reps);
            __syncthreads();
            if (tid == 0) g_out[slot] = (unsigned long long)(clock64() - t0);
            if (a == 0x12345) g_sink = a;
            ++slot;
        }
    __syncthreads();
    if (warp == 0) tc5::tmem_free(tmem, 512);
}

int main() {
    const int reps = 256;
    bench<<<1, 512>>>(reps);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    unsigned long long out[32]; cudaMemcpyFromSymbol(out, g_out, sizeof out);
    const char *names[] = {"ld x32 (wait each)", "4 x ld x32 in flight", "st x16", "ld x16 pack::16b (32 cols)"};
    const double bytes_per_warp_rep[] = {32 * 32 * 4.0, 4 * 32 * 32 * 4.0, 16 * 32 * 4.0, 32 * 32 * 4.0};
    int slot = 0;
    for (int nw = 4; nw <= 16; nw *= 2)
        for (int mode = 0; mode < 4; ++mode, ++slot)
            printf("%2d warps  %-28s %8llu clk  -> %6.1f B/clk/SM (TMEM-side bytes)\n", nw, names[mode], out[slot], bytes_per_warp_rep[mode] * reps * nw / out[slot]);
    return 0;
}
