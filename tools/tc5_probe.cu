// tc5_probe.cu -- phase timers of the sequential tcgen05 block kernel (developer tool, not part of the library).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DB200SR_TC5_PROF -I mobilesuperresolution_b200/csrc tools/tc5_probe.cu -o gpurun_out/tc5_probe
#include <cstdio>
#include <vector>
#include "wdsr_tc5.cuh"
using namespace b200sr;
int main() {
    const int N = 64, H = 96, W = 96, M1P = 144;
    BlockTc5Layout L(M1P);
    std::vector<uint8_t> img(L.total, 0);
    uint8_t *dimg; bf16 *din, *dout;
    cudaMalloc(&dimg, L.total); cudaMemcpy(dimg, img.data(), L.total, cudaMemcpyHostToDevice);
    size_t nb = (size_t)N * H * W * 24 * 2;
    cudaMalloc(&din, nb); cudaMalloc(&dout, nb); cudaMemset(din, 0, nb);
    const int tx = ceil_div(W, 32), ty = ceil_div(H, 16), ntiles = tx * ty * N;
    size_t smem = wdsr_block_tc5_seq_smem(M1P);
    cudaFuncSetAttribute(wdsr_block_tc5_seq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int rep = 0; rep < 2; ++rep) {
        unsigned long long z[64] = {0};
        cudaMemcpyToSymbol(g_tc5_prof, z, sizeof z);
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a);
        wdsr_block_tc5_seq_kernel<<<148, 128, smem>>>(din, dout, dimg, M1P, N, H, W, tx, ty, ntiles);
        cudaEventRecord(b);
        cudaError_t e = cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, a, b);
        unsigned long long p[64]; cudaMemcpyFromSymbol(p, g_tc5_prof, sizeof p);
        const int tiles0 = (ntiles - 1) / 148 + 1;
        printf("rep %d: %s, %.1f us, CTA0 ran %d tiles; cycles per tile: load %llu | G1 %llu (x5) E1 %llu G2 %llu E2 %llu | G3 %llu (x4) E3 %llu\n",
               rep, cudaGetErrorString(e), ms * 1e3, tiles0, p[0] / tiles0, p[1] / tiles0, p[2] / tiles0, p[3] / tiles0, p[4] / tiles0,
               p[5] / tiles0, p[6] / tiles0);
    }
    size_t smem2 = wdsr_block_tc5_smem(M1P);
    cudaFuncSetAttribute(wdsr_block_tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
    for (int rep = 0; rep < 2; ++rep) {
        unsigned long long z[64] = {0};
        cudaMemcpyToSymbol(g_tc5_prof, z, sizeof z);
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a);
        wdsr_block_tc5_kernel<<<148, tc5p::NTHREADS, smem2>>>(din, dout, dimg, M1P, N, H, W, tx, ty, ntiles);
        cudaEventRecord(b);
        cudaError_t e = cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, a, b);
        unsigned long long p[64]; cudaMemcpyFromSymbol(p, g_tc5_prof, sizeof p);
        const int t = (ntiles - 1) / 148 + 1;
        printf("pipelined rep %d: %s, %.1f us (%.0f clk/tile at 1.9GHz); MMA-warp wait clk/tile: A2_FULL %llu D2_EMPTY %llu T2_FULL %llu D3_EMPTY %llu XS_FULL %llu\n",
               rep, cudaGetErrorString(e), ms * 1e3, ms * 1e-3 * 1.9e9 / t, p[0] / t, p[1] / t, p[2] / t, p[3] / t, p[4] / t);
        for (int w = 0; w < 8; ++w)
            printf("   epi warp %d (WG%d): wait D1_FULL %llu D2_FULL %llu T2_EMPTY %llu D3_FULL %llu\n", w + 2, w / 4, p[8 + 4 * w] / t, p[9 + 4 * w] / t, p[10 + 4 * w] / t, p[11 + 4 * w] / t);
    }
    return 0;
}
