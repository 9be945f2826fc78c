// wdsr_bf16_tail.cu -- launcher of the fused bf16 tensor-core tail (tail 3x3 + skip 5x5 + PixelShuffle + mean).
#include "launch.h"
#include "wdsr_bf16.cuh"

namespace b200sr {

template <typename TIN, typename TOUT, int CP, int S>
static cudaError_t tail_bf16_t(const void *trunk, const void *x, void *y, const uint8_t *wimg, int N, int H, int W, float mean,
                               float out_add, cudaStream_t st) {
    constexpr int TW = 32, TH = 16, NWARPS = 8;
    auto kern = wdsr_tail_bf16_kernel<TIN, TOUT, CP, S, TW, TH, NWARPS>;
    const size_t smem = wdsr_tail_bf16_smem<CP, S, TW, TH>();
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int tx = ceil_div(W, TW), ty = ceil_div(H, TH);
    const int ntiles = tx * ty * N;
    int ctas = 2 * sm_count();
    if (ctas > ntiles) ctas = ntiles;
    kern<<<ctas, NWARPS * 32, smem, st>>>((const bf16 *)trunk, (const TIN *)x, (TOUT *)y, wimg, N, H, W, tx, ty, ntiles, mean,
                                          out_add);
    return cudaGetLastError();
}

template <int CP, int S>
static cudaError_t tail_bf16_io(int xd, int yd, const void *trunk, const void *x, void *y, const uint8_t *wimg, int N, int H,
                                int W, float mean, float out_add, cudaStream_t st) {
    if (xd == kF32 && yd == kF32) return tail_bf16_t<float, float, CP, S>(trunk, x, y, wimg, N, H, W, mean, out_add, st);
    if (xd == kF32 && yd == kBF16) return tail_bf16_t<float, bf16, CP, S>(trunk, x, y, wimg, N, H, W, mean, out_add, st);
    if (xd == kBF16 && yd == kF32) return tail_bf16_t<bf16, float, CP, S>(trunk, x, y, wimg, N, H, W, mean, out_add, st);
    if (xd == kBF16 && yd == kBF16) return tail_bf16_t<bf16, bf16, CP, S>(trunk, x, y, wimg, N, H, W, mean, out_add, st);
    return cudaErrorInvalidValue;
}

template <int CP>
static cudaError_t tail_bf16_s(int S, int xd, int yd, const void *trunk, const void *x, void *y, const uint8_t *wimg, int N,
                               int H, int W, float mean, float out_add, cudaStream_t st) {
    switch (S) {
        case 2: return tail_bf16_io<CP, 2>(xd, yd, trunk, x, y, wimg, N, H, W, mean, out_add, st);
        case 3: return tail_bf16_io<CP, 3>(xd, yd, trunk, x, y, wimg, N, H, W, mean, out_add, st);
        case 4: return tail_bf16_io<CP, 4>(xd, yd, trunk, x, y, wimg, N, H, W, mean, out_add, st);
    }
    return cudaErrorInvalidValue;
}

cudaError_t launch_tail_bf16(int CP, int S, int xd, int yd, const void *trunk, const void *x, void *y, const uint8_t *wimg,
                             int N, int H, int W, float mean, float out_add, cudaStream_t st) {
    switch (CP) {
        case 8: return tail_bf16_s<8>(S, xd, yd, trunk, x, y, wimg, N, H, W, mean, out_add, st);
        case 16: return tail_bf16_s<16>(S, xd, yd, trunk, x, y, wimg, N, H, W, mean, out_add, st);
        case 24: return tail_bf16_s<24>(S, xd, yd, trunk, x, y, wimg, N, H, W, mean, out_add, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace b200sr
