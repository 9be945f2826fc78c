// wdsr_f32.cu -- launchers of the fp32 (FFMA) kernels.
#include "launch.h"
#include "wdsr_f32.cuh"

namespace b200sr {

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (!cached[dev]) {
        int n = 0;
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        cached[dev] = n > 0 ? n : 148;
    }
    return cached[dev];
}

template <typename TIN, typename TOUT, int CP>
static cudaError_t head_t(const void *x, void *trunk, const float *wpack, int N, int H, int W, float mean, cudaStream_t st) {
    const int tx = ceil_div(W, 64), ty = ceil_div(H, 4);
    wdsr_head_kernel<TIN, TOUT, CP><<<tx * ty * N, 256, 0, st>>>((const TIN *)x, (TOUT *)trunk, wpack, N, H, W, mean, tx, ty);
    return cudaGetLastError();
}

template <int CP>
static cudaError_t head_cp(int xd, int td, const void *x, void *trunk, const float *wpack, int N, int H, int W, float mean,
                           cudaStream_t st) {
    if (xd == kF32 && td == kF32) return head_t<float, float, CP>(x, trunk, wpack, N, H, W, mean, st);
    if (xd == kF32 && td == kBF16) return head_t<float, bf16, CP>(x, trunk, wpack, N, H, W, mean, st);
    if (xd == kBF16 && td == kF32) return head_t<bf16, float, CP>(x, trunk, wpack, N, H, W, mean, st);
    if (xd == kBF16 && td == kBF16) return head_t<bf16, bf16, CP>(x, trunk, wpack, N, H, W, mean, st);
    return cudaErrorInvalidValue;
}

cudaError_t launch_head(int CP, int xd, int td, const void *x, void *trunk, const float *wpack, int N, int H, int W,
                        float mean, cudaStream_t st) {
    switch (CP) {
        case 8: return head_cp<8>(xd, td, x, trunk, wpack, N, H, W, mean, st);
        case 16: return head_cp<16>(xd, td, x, trunk, wpack, N, H, W, mean, st);
        case 24: return head_cp<24>(xd, td, x, trunk, wpack, N, H, W, mean, st);
    }
    return cudaErrorInvalidValue;
}

template <int CP, int M2P, int TW, int TH>
static cudaError_t block_f32_t(const float *in, float *out, const float *wpack, int M1P, int N, int H, int W,
                               cudaStream_t st) {
    auto kern = wdsr_block_f32_kernel<CP, M2P, TW, TH>;
    const size_t smem = wdsr_block_f32_smem<CP, M2P, TW, TH>(M1P);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int tx = ceil_div(W, TW), ty = ceil_div(H, TH);
    kern<<<tx * ty * N, TW * TH, smem, st>>>(in, out, wpack, M1P, N, H, W, tx, ty);
    return cudaGetLastError();
}

template <int CP, int M2P>
static cudaError_t block_f32_tile(const float *in, float *out, const float *wpack, int M1P, int N, int H, int W,
                                  cudaStream_t st) {
    // small problems (cfg1: one 64x64 patch) cannot fill the machine with 16x16 tiles -> 8x8 tiles
    const long long big_tiles = (long long)ceil_div(W, 16) * ceil_div(H, 16) * N;
    if (big_tiles < 2ll * sm_count()) return block_f32_t<CP, M2P, 8, 8>(in, out, wpack, M1P, N, H, W, st);
    return block_f32_t<CP, M2P, 16, 16>(in, out, wpack, M1P, N, H, W, st);
}

template <int CP>
static cudaError_t block_f32_cp(int M2P, const float *in, float *out, const float *wpack, int M1P, int N, int H, int W,
                                cudaStream_t st) {
    switch (M2P) {
        case 8: return block_f32_tile<CP, 8>(in, out, wpack, M1P, N, H, W, st);
        case 12: return block_f32_tile<CP, 12>(in, out, wpack, M1P, N, H, W, st);
        case 16: return block_f32_tile<CP, 16>(in, out, wpack, M1P, N, H, W, st);
        case 20: return block_f32_tile<CP, 20>(in, out, wpack, M1P, N, H, W, st);
        case 24: return block_f32_tile<CP, 24>(in, out, wpack, M1P, N, H, W, st);
    }
    return cudaErrorInvalidValue;
}

cudaError_t launch_block_f32(int CP, int M2P, const float *in, float *out, const float *wpack, int M1P, int N, int H, int W,
                             cudaStream_t st) {
    switch (CP) {
        case 8: return block_f32_cp<8>(M2P, in, out, wpack, M1P, N, H, W, st);
        case 16: return block_f32_cp<16>(M2P, in, out, wpack, M1P, N, H, W, st);
        case 24: return block_f32_cp<24>(M2P, in, out, wpack, M1P, N, H, W, st);
    }
    return cudaErrorInvalidValue;
}

template <typename TIN, typename TOUT, int CP, int S>
static cudaError_t tail_f32_t(const float *trunk, const void *x, void *y, const float *wpack, int N, int H, int W, float mean,
                              float out_add, cudaStream_t st) {
    constexpr int TW = 16, TH = 8;
    auto kern = wdsr_tail_f32_kernel<TIN, TOUT, CP, S, TW, TH>;
    const size_t smem = wdsr_tail_f32_smem<CP, S, TW, TH>();
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int tx = ceil_div(W, TW), ty = ceil_div(H, TH);
    kern<<<tx * ty * N, TW * TH, smem, st>>>(trunk, (const TIN *)x, (TOUT *)y, wpack, N, H, W, tx, ty, mean, out_add);
    return cudaGetLastError();
}

template <int CP, int S>
static cudaError_t tail_f32_io(int xd, int yd, const float *trunk, const void *x, void *y, const float *wpack, int N, int H,
                               int W, float mean, float out_add, cudaStream_t st) {
    if (xd == kF32 && yd == kF32) return tail_f32_t<float, float, CP, S>(trunk, x, y, wpack, N, H, W, mean, out_add, st);
    if (xd == kBF16 && yd == kBF16) return tail_f32_t<bf16, bf16, CP, S>(trunk, x, y, wpack, N, H, W, mean, out_add, st);
    return cudaErrorInvalidValue;  // fp32 arithmetic path: x and y share a dtype
}

template <int CP>
static cudaError_t tail_f32_s(int S, int xd, int yd, const float *trunk, const void *x, void *y, const float *wpack, int N,
                              int H, int W, float mean, float out_add, cudaStream_t st) {
    switch (S) {
        case 2: return tail_f32_io<CP, 2>(xd, yd, trunk, x, y, wpack, N, H, W, mean, out_add, st);
        case 3: return tail_f32_io<CP, 3>(xd, yd, trunk, x, y, wpack, N, H, W, mean, out_add, st);
        case 4: return tail_f32_io<CP, 4>(xd, yd, trunk, x, y, wpack, N, H, W, mean, out_add, st);
    }
    return cudaErrorInvalidValue;
}

cudaError_t launch_tail_f32(int CP, int S, int xd, int yd, const float *trunk, const void *x, void *y, const float *wpack,
                            int N, int H, int W, float mean, float out_add, cudaStream_t st) {
    switch (CP) {
        case 8: return tail_f32_s<8>(S, xd, yd, trunk, x, y, wpack, N, H, W, mean, out_add, st);
        case 16: return tail_f32_s<16>(S, xd, yd, trunk, x, y, wpack, N, H, W, mean, out_add, st);
        case 24: return tail_f32_s<24>(S, xd, yd, trunk, x, y, wpack, N, H, W, mean, out_add, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace b200sr
