// rs_prod_bench.cu -- issue cost of the X-ring producer alternatives of wdsr_rs.cuh, one warp per CTA (developer probe).
#include <cstdio>
#include "tc5.cuh"
using namespace b200sr;
__device__ unsigned long long g_out[64];
constexpr int NX = 16, XPLANE = 2048, XSLOT = 3 * XPLANE;

__global__ void __launch_bounds__(128, 1) bench(const uint8_t *in, int H, int W, int rows, int mode) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *xs = smem + 512;
    const uint32_t bars = smem_u32(smem), xs_u = smem_u32(xs);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { for (int b = 0; b < NX; ++b) tc5::mbar_init(bars + 8 * b, mode == 0 ? 32 : 1); tc5::mbar_init_fence(); }
    __syncthreads();
    const long long plane = (long long)H * W * 16;
    const int n = blockIdx.x % 8;
    if (warp == 0) {
        const uint8_t *src4[4];
        for (int k = 0; k < 4; ++k) { const int p = lane + 32 * k; src4[k] = p < W ? in + ((long long)n * 3 * H * W + p) * 16 : nullptr; }
        long long t_issue = 0;
        const long long t0 = clock64();
        for (int s = 0; s < rows; ++s) {
            const int slot = s % NX;
            if (s >= NX) tc5::mbar_wait(bars + 8 * slot, ((s / NX) - 1) & 1);   // (consumer = nobody: wait for the row itself to have landed)
            const long long a = clock64();
            const int y = s % H;
            if (mode == 0) {
                uint8_t *dst = xs + slot * XSLOT + lane * 16;
                const long long row = (long long)y * W * 16;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (src4[k]) {
#pragma unroll
                        for (int q = 0; q < 3; ++q) cp_async16(dst + q * XPLANE + k * 512, src4[k] + row + q * plane, 16);
                    }
                tc5::cp_async_arrive_noinc(bars + 8 * slot);
            } else if (mode == 1) {   // bulk: one copy per plane of W pixels, issued by lane 0
                if (lane == 0) {
                    tc5::mbar_arrive_expect_tx(bars + 8 * slot, 48u * W);
                    for (int q = 0; q < 3; ++q) tc5::bulk_load_g2s(xs_u + slot * XSLOT + q * XPLANE, in + ((long long)n * 3 * H * W + (long long)(q * H + y) * W) * 16, 16u * W, bars + 8 * slot);
                }
                __syncwarp();
            } else {                  // bulk: 3 runs x 3 planes issued by lanes 0..2
                if (lane == 0) tc5::mbar_arrive_expect_tx(bars + 8 * slot, 48u * (W / 3) * 3);
                __syncwarp();
                if (lane < 3)
                    for (int q = 0; q < 3; ++q) tc5::bulk_load_g2s(xs_u + slot * XSLOT + q * XPLANE + lane * (W / 3) * 16, in + ((long long)n * 3 * H * W + (long long)(q * H + y) * W + lane * (W / 3)) * 16, 16u * (W / 3), bars + 8 * slot);
                __syncwarp();
            }
            t_issue += clock64() - a;
        }
        const long long t1 = clock64();
        if (blockIdx.x == 0 && lane == 0) { g_out[2 * mode] = (unsigned long long)t_issue; g_out[2 * mode + 1] = (unsigned long long)(t1 - t0); }
    }
}
int main() {
    const int H = 96, W = 96, rows = 256;
    uint8_t *din; cudaMalloc(&din, (size_t)8 * 3 * H * W * 16); cudaMemset(din, 0, (size_t)8 * 3 * H * W * 16);
    size_t smem = 512 + NX * XSLOT;
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int grid : {1, 148}) {
        for (int mode = 0; mode < 3; ++mode) {
            bench<<<grid, 128, smem>>>(din, H, W, rows, mode);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
        }
        unsigned long long out[64]; cudaMemcpyFromSymbol(out, g_out, sizeof out);
        const char *names[] = {"cp.async 12 / lane + arrive.noinc", "cp.async.bulk x3 (one per plane)", "cp.async.bulk 3 runs x 3 planes"};
        for (int m = 0; m < 3; ++m) printf("grid %3d  %-36s issue %6.0f clk/row   loop %6.0f clk/row (16 rows in flight)\n", grid, names[m], (double)out[2 * m] / rows, (double)out[2 * m + 1] / rows);
    }
    return 0;
}
