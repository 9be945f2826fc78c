// conv.cu -- launchers of the generic NHWC convolution kernels (video path).
#include "conv.cuh"
#include "launch.h"

namespace b200sr {

template <int K, typename T>
static cudaError_t conv_f32_t(const ConvArgs &a, cudaStream_t st) {
    auto kern = conv_f32_kernel<K, T>;
    constexpr size_t smem = conv_f32_smem<K>();
    static thread_local SmemOptIn optin;   // per device (launch.h)
    if (cudaError_t e = optin.ensure(kern, smem); e != cudaSuccess) return e;
    const int tx = ceil_div(a.w_, 16), ty = ceil_div(a.h, 8);
    kern<<<dim3(tx * ty * a.n, a.coutp / 32), 256, smem, st>>>(a, tx, ty);
    return cudaGetLastError();
}

template <int K, int NT, typename TIN, typename TOUT>
static cudaError_t conv_bf16_t(const ConvArgs &a, cudaStream_t st) {
    auto kern = conv_bf16_kernel<K, NT, TIN, TOUT>;
    constexpr size_t smem = conv_bf16_smem<K, NT>();
    static thread_local SmemOptIn optin;   // per device (launch.h)
    if (cudaError_t e = optin.ensure(kern, smem); e != cudaSuccess) return e;
    const int tx = ceil_div(a.w_, 16), ty = ceil_div(a.h, 8);
    kern<<<dim3(tx * ty * a.n, a.coutp / (8 * NT)), 128, smem, st>>>(a, tx, ty);
    return cudaGetLastError();
}

template <int K, typename TIN, typename TOUT>
static cudaError_t conv_bf16_nt(int nt, const ConvArgs &a, cudaStream_t st) {
    switch (nt) {
        case 1: return conv_bf16_t<K, 1, TIN, TOUT>(a, st);
        case 2: return conv_bf16_t<K, 2, TIN, TOUT>(a, st);
        case 4: return conv_bf16_t<K, 4, TIN, TOUT>(a, st);
        case 8: return conv_bf16_t<K, 8, TIN, TOUT>(a, st);
    }
    return cudaErrorInvalidValue;
}

template <int K>
static cudaError_t conv_bf16_io(int nt, int in_dtype, int out_dtype, const ConvArgs &a, cudaStream_t st) {
    if (in_dtype == kBF16 && out_dtype == kBF16) return conv_bf16_nt<K, bf16, bf16>(nt, a, st);
    if (in_dtype == kBF16 && out_dtype == kF32) return conv_bf16_nt<K, bf16, float>(nt, a, st);
    if (in_dtype == kF32 && out_dtype == kBF16) return conv_bf16_nt<K, float, bf16>(nt, a, st);
    return cudaErrorInvalidValue;
}

cudaError_t launch_conv(const ConvArgs &a, int k, int nt, int in_dtype, int out_dtype, int precision, cudaStream_t st) {
    if (precision == kF32) {
        if (in_dtype != kF32 || out_dtype != kF32) return cudaErrorInvalidValue;
        switch (k) {
            case 1: return conv_f32_t<1, float>(a, st);
            case 3: return conv_f32_t<3, float>(a, st);
            case 5: return conv_f32_t<5, float>(a, st);
            case 7: return conv_f32_t<7, float>(a, st);
        }
        return cudaErrorInvalidValue;
    }
    switch (k) {
        case 1: return conv_bf16_io<1>(nt, in_dtype, out_dtype, a, st);
        case 3: return conv_bf16_io<3>(nt, in_dtype, out_dtype, a, st);
        case 5: return conv_bf16_io<5>(nt, in_dtype, out_dtype, a, st);
        case 7: return conv_bf16_io<7>(nt, in_dtype, out_dtype, a, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace b200sr
