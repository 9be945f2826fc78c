// f16acc_test.cu -- does tcgen05.mma kind::f16 accept bf16 (or mixed) inputs with an F16 accumulator, and how is the f16 D laid out in TMEM?
// Can the f16 D be fed back directly as the (TMEM) A operand of the next MMA?  (developer probe)
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include "tc5.cuh"
using namespace b200sr;
__device__ uint32_t g_raw[5][128][32];
__device__ float g_d2[128][32];

// element (row r, k) of a K-major SWIZZLE_NONE operand with LBO = 128 B (k-chunk stride), SBO = 256 B (8-row group stride), K = 16
__device__ __host__ inline int op_off(int r, int k) { return (r / 8) * 256 + (k / 8) * 128 + (r % 8) * 16 + (k % 8) * 2; }

__global__ void __launch_bounds__(128, 1) test(int afmt, int bfmt, int cfmt, int slot) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *ctrl = smem, *A = smem + 1024, *B = smem + 1024 + 8192, *B2 = B + 8192;
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t bar = smem_u32(ctrl);
    if (tid == 0) { tc5::mbar_init(bar, 1); tc5::mbar_init_fence(); }
    __syncwarp();
    if (warp == 0) tc5::tmem_alloc(smem_u32(ctrl + 64), 128);
    for (int i = tid; i < 3 * 8192 / 4; i += 128) reinterpret_cast<uint32_t *>(A)[i] = 0;
    __syncthreads();
    // A[r][0] = 1 + r/128, A[r][1] = 0.5 ; B[n][0] = n + 1, B[n][1] = -2 (rows n < 32)   =>  D[r][n] = (1 + r/128)(n+1) - 1
    {
        const int r = tid;
        const float a0 = 1.f + r / 128.f, a1 = 0.5f;
        if (afmt == 1) { *reinterpret_cast<__nv_bfloat16 *>(A + op_off(r, 0)) = __float2bfloat16(a0); *reinterpret_cast<__nv_bfloat16 *>(A + op_off(r, 1)) = __float2bfloat16(a1); }
        else { *reinterpret_cast<__half *>(A + op_off(r, 0)) = __float2half(a0); *reinterpret_cast<__half *>(A + op_off(r, 1)) = __float2half(a1); }
        if (r < 32) {
            const float b0 = r + 1.f, b1 = -2.f;
            if (bfmt == 1) { *reinterpret_cast<__nv_bfloat16 *>(B + op_off(r, 0)) = __float2bfloat16(b0); *reinterpret_cast<__nv_bfloat16 *>(B + op_off(r, 1)) = __float2bfloat16(b1); }
            else { *reinterpret_cast<__half *>(B + op_off(r, 0)) = __float2half(b0); *reinterpret_cast<__half *>(B + op_off(r, 1)) = __float2half(b1); }
            // second GEMM: D2[r][n] = sum_k A2[r][k] * B2[n][k], B2[n][k] = (k == n) -> D2 = first 16 columns of A2 (identity pick), f16
            for (int k = 0; k < 16; ++k) *reinterpret_cast<__half *>(B2 + op_off(r, k)) = __float2half(k == r ? 1.f : 0.f);
        }
    }
    tc5::fence_proxy_async(); tc5::fence_before_sync(); __syncthreads(); tc5::fence_after_sync();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t *>(ctrl + 64);
    const uint32_t idesc = ((uint32_t)cfmt << 4) | ((uint32_t)afmt << 7) | ((uint32_t)bfmt << 10) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (warp == 0 && tc5::elect_one()) {
        tc5::mma_ss(tmem, tc5::smem_desc(smem_u32(A), 128, 256), tc5::smem_desc(smem_u32(B), 128, 256), idesc, false);
        tc5::commit(bar);
    }
    tc5::mbar_wait(bar, 0);
    tc5::fence_after_sync();
    uint32_t v[32];
    tc5::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16), v);
    tc5::tmem_wait_ld();
    for (int j = 0; j < 32; ++j) g_raw[slot][tid][j] = v[j];
    if (cfmt == 0) {
        // feed D (f16, as it lies in TMEM columns 0..) back as the A operand (K = 16 -> 8 columns) of a second MMA, f16 x f16 -> f32
        tc5::fence_before_sync(); __syncthreads(); tc5::fence_after_sync();
        const uint32_t idesc2 = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        if (warp == 0 && tc5::elect_one()) {
            tc5::mma_ts(tmem + 64, tmem, tc5::smem_desc(smem_u32(B2), 128, 256), idesc2, false);
            tc5::commit(bar);
        }
        tc5::mbar_wait(bar, 1);
        tc5::fence_after_sync();
        tc5::tmem_ld32(tmem + 64 + ((uint32_t)(warp * 32) << 16), v);
        tc5::tmem_wait_ld();
        for (int j = 0; j < 32; ++j) g_d2[tid][j] = __uint_as_float(v[j]);
    }
    tc5::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc5::tmem_free(tmem, 128);
}

static float h2f(uint16_t h) { __half x; memcpy(&x, &h, 2); return __half2float(x); }
int main(int argc, char **argv) {
    const int only = argc > 1 ? atoi(argv[1]) : -1;
    cudaFuncSetAttribute(test, cudaFuncAttributeMaxDynamicSharedMemorySize, 1024 + 3 * 8192);
    const int cfgs[5][3] = {{1, 1, 1}, {1, 1, 0}, {1, 0, 0}, {0, 0, 0}, {1, 0, 1}};   // (afmt, bfmt, cfmt): 1 = bf16 / f32 accumulate, 0 = f16
    static uint32_t raw[5][128][32]; static float d2[128][32];
    for (int c = 0; c < 5; ++c) {
        if (only >= 0 && c != only) continue;
        test<<<1, 128, 1024 + 3 * 8192>>>(cfgs[c][0], cfgs[c][1], cfgs[c][2], c);
        cudaError_t e = cudaDeviceSynchronize();
        printf("A=%s B=%s D=%s : %s\n", cfgs[c][0] ? "bf16" : "f16", cfgs[c][1] ? "bf16" : "f16", cfgs[c][2] ? "f32" : "f16", cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        cudaMemcpyFromSymbol(raw, g_raw, sizeof raw);
        for (int r : {0, 1, 64, 127}) {
            printf("  lane %3d expect D[n] = %.4f*(n+1)-1 :", r, 1.f + r / 128.f);
            for (int j = 0; j < 6; ++j) {
                if (cfgs[c][2]) { float f; memcpy(&f, &raw[c][r][j], 4); printf(" %.4f", f); }
                else printf(" [%.4f %.4f]", h2f(raw[c][r][j] & 0xffff), h2f(raw[c][r][j] >> 16));
            }
            printf(" ...\n");
        }
        if (!cfgs[c][2]) {
            cudaMemcpyFromSymbol(d2, g_d2, sizeof d2);
            printf("  TS feed-back (identity pick of the first 16 f16 of D as A operand): lane 1:");
            for (int j = 0; j < 8; ++j) printf(" %.4f", d2[1][j]);
            printf("\n");
        }
    }
    return 0;
}
