// rs_e1_bench.cu -- TMEM round trip of E1 (tcgen05.ld fp32 -> relu/pack bf16x2 -> tcgen05.st, in place) alone: clk per 128 x 144 tile
// for different splits over warpgroups (developer probe for wdsr_rs.cuh).
#include <cstdio>
#include "tc5.cuh"
using namespace b200sr;
__device__ unsigned long long g_out[64];

template <int NCOL>   // columns this thread converts (multiple of 16, <= 80)
__device__ __forceinline__ void e1_body(uint32_t d1, uint32_t dst, int mode) {
    uint32_t v[NCOL];
#pragma unroll
    for (int c = 0; c < NCOL; c += 16) tc5::tmem_ld16(d1 + c, *reinterpret_cast<uint32_t(*)[16]>(&v[c]));
    tc5::tmem_wait_ld();
    if (mode == 1) { if (v[0] == 0x12345u) g_out[63] = 1; return; }   // ld only
#pragma unroll
    for (int j = 0; j < NCOL / 2; ++j) v[j] = tc5::relu_pack_bf16x2(v[2 * j], v[2 * j + 1]);
#pragma unroll
    for (int c = 0; c < NCOL / 2; c += 8) tc5::tmem_st8(dst + c, *reinterpret_cast<uint32_t(*)[8]>(&v[c]));
    tc5::tmem_wait_st();
}

__global__ void __launch_bounds__(768, 1) bench(int reps) {
    __shared__ uint32_t tptr;
    const int tid = threadIdx.x, warp = tid >> 5, wg = warp >> 2;
    if (warp == 0) tc5::tmem_alloc(smem_u32(&tptr), 512);
    tc5::fence_before_sync(); __syncthreads(); tc5::fence_after_sync();
    const uint32_t tmem = tptr;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    for (int mode = 0; mode < 6; ++mode) {
        __syncthreads();
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            const uint32_t d1 = tmem + lane_base + (r & 1) * 144;
            if (mode == 0 || mode == 1) {            // kernel split: WG1 64 cols, WG2 80 cols   (mode 1: loads only)
                if (wg == 1) e1_body<64>(d1, d1, mode);
                else if (wg == 2) e1_body<80>(d1 + 64, d1 + 104, mode);
            } else if (mode == 2) {                  // four warpgroups: 32 + 32 + 48 + 32
                if (wg == 1) e1_body<32>(d1, d1, 0);
                else if (wg == 2) e1_body<32>(d1 + 32, d1 + 16, 0);   // (destination overlaps wg1's source in the real kernel; timing only)
                else if (wg == 3) e1_body<48>(d1 + 64, d1 + 104, 0);
                else if (wg == 4) e1_body<32>(d1 + 112, d1 + 128, 0);
            } else if (mode == 3) {                  // only one warpgroup active, 64 cols
                if (wg == 1) e1_body<64>(d1, d1, 0);
            } else if (mode == 4) {                  // only one warpgroup, 16 cols
                if (wg == 1) e1_body<16>(d1, d1, 0);
            } else {                                 // five warpgroups: 32,32,32,32,16
                if (wg >= 1 && wg <= 4) e1_body<32>(d1 + 32 * (wg - 1), d1 + 16 * (wg - 1), 0);
                else if (wg == 5) e1_body<16>(d1 + 128, d1 + 136, 0);
            }
        }
        const long long t1 = clock64();
        __syncthreads();
        if (tid == 32 * 4) g_out[mode] = (unsigned long long)(t1 - t0);
        if (tid == 32 * 8) g_out[8 + mode] = (unsigned long long)(t1 - t0);
    }
    tc5::fence_before_sync(); __syncthreads();
    if (warp == 0) tc5::tmem_free(tmem, 512);
}
int main() {
    const int reps = 64;
    bench<<<1, 768>>>(reps);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    unsigned long long out[64]; cudaMemcpyFromSymbol(out, g_out, sizeof out);
    const char *names[] = {"E1 as in the kernel (WG1 64, WG2 80 cols)", "  loads only", "four warpgroups (32/32/48/32)", "one warpgroup, 64 cols", "one warpgroup, 16 cols", "five warpgroups (4 x 32 + 16)"};
    for (int i = 0; i < 6; ++i) printf("%-46s %7.0f clk per tile (WG1)  %7.0f (WG2)\n", names[i], (double)out[i] / reps, (double)out[8 + i] / reps);
    return 0;
}
