// wdsr_bf16_block.cu -- launcher of the fused bf16 tensor-core residual-block kernel.
#include "launch.h"
#include "wdsr_bf16.cuh"

namespace b200sr {

template <int CP, int M2P>
static cudaError_t block_bf16_t(const void *in, void *out, const uint8_t *wimg, int M1P, int N, int H, int W, cudaStream_t st) {
    constexpr int TW = 32, TH = 16, NWARPS = 8;
    auto kern = wdsr_block_bf16_kernel<CP, M2P, TW, TH, NWARPS>;
    const size_t smem = wdsr_block_bf16_smem<CP, M2P, TW, TH>(M1P);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int tx = ceil_div(W, TW), ty = ceil_div(H, TH);
    const int ntiles = tx * ty * N;
    int ctas = 2 * sm_count();  // persistent: two co-resident CTAs per SM, each loops over tiles
    if (ctas > ntiles) ctas = ntiles;
    kern<<<ctas, NWARPS * 32, smem, st>>>((const bf16 *)in, (bf16 *)out, wimg, M1P, N, H, W, tx, ty, ntiles);
    return cudaGetLastError();
}

template <int CP>
static cudaError_t block_bf16_cp(int M2P, const void *in, void *out, const uint8_t *wimg, int M1P, int N, int H, int W,
                                 cudaStream_t st) {
    switch (M2P) {
        case 8: return block_bf16_t<CP, 8>(in, out, wimg, M1P, N, H, W, st);
        case 16: return block_bf16_t<CP, 16>(in, out, wimg, M1P, N, H, W, st);
        case 24: return block_bf16_t<CP, 24>(in, out, wimg, M1P, N, H, W, st);
    }
    return cudaErrorInvalidValue;
}

cudaError_t launch_block_bf16(int CP, int M2P, const void *in, void *out, const uint8_t *wimg, int M1P, int N, int H, int W,
                              cudaStream_t st) {
    switch (CP) {
        case 8: return block_bf16_cp<8>(M2P, in, out, wimg, M1P, N, H, W, st);
        case 16: return block_bf16_cp<16>(M2P, in, out, wimg, M1P, N, H, W, st);
        case 24: return block_bf16_cp<24>(M2P, in, out, wimg, M1P, N, H, W, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace b200sr
