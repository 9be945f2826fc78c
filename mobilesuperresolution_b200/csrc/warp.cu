// warp.cu -- flow_warp: bilinear gather at (x + fx, y + fy).          models/spynet_arch.py:98-129
//
// The reference builds a pixel mesh, adds the flow, normalises to [-1,1] (`2*v/max(size-1,1) - 1`) and calls
// F.grid_sample(bilinear, padding_mode, align_corners=True), which un-normalises with ((g+1)/2)*(size-1).  The
// kernels replay that round trip in fp32 so the sampling position matches the reference to the last ulp, then
//   zeros : each of the 4 corners contributes only if it lies inside the image
//   border: the coordinate is clamped to [0, size-1] before it is split into corners.
// Coordinates and blend weights are always fp32, also when the features are bf16 (bf16 resolves only 2 px at x~300).
#include "common.cuh"
#include "launch.h"

namespace b200sr {

struct Bilin {
    int x0, y0;
    float w00, w01, w10, w11;  // (y0,x0) (y0,x1) (y1,x0) (y1,x1), already zeroed for out-of-image corners
    bool v00, v01, v10, v11;
};

__device__ __forceinline__ Bilin bilinear_setup(float fx, float fy, int xw, int yh, int W, int H, bool border) {
    const float dw = (float)max(W - 1, 1), dh = (float)max(H - 1, 1);
    const float gx = 2.0f * ((float)xw + fx) / dw - 1.0f;
    const float gy = 2.0f * ((float)yh + fy) / dh - 1.0f;
    float ix = ((gx + 1.f) / 2.f) * (float)(W - 1);
    float iy = ((gy + 1.f) / 2.f) * (float)(H - 1);
    if (border) {
        ix = fminf((float)(W - 1), fmaxf(ix, 0.f));
        iy = fminf((float)(H - 1), fmaxf(iy, 0.f));
    }
    // keep far-away samples finite for the int conversion; they contribute nothing in 'zeros' mode
    ix = fminf(fmaxf(ix, -2.f), (float)W + 1.f);
    iy = fminf(fmaxf(iy, -2.f), (float)H + 1.f);
    const float fx0 = floorf(ix), fy0 = floorf(iy);
    Bilin b;
    b.x0 = (int)fx0;
    b.y0 = (int)fy0;
    const float ex = (fx0 + 1.f) - ix, wx = ix - fx0, ey = (fy0 + 1.f) - iy, wy = iy - fy0;
    const bool vx0 = b.x0 >= 0 && b.x0 < W, vx1 = b.x0 + 1 >= 0 && b.x0 + 1 < W;
    const bool vy0 = b.y0 >= 0 && b.y0 < H, vy1 = b.y0 + 1 >= 0 && b.y0 + 1 < H;
    b.v00 = vy0 && vx0, b.v01 = vy0 && vx1, b.v10 = vy1 && vx0, b.v11 = vy1 && vx1;
    b.w00 = ex * ey, b.w01 = wx * ey, b.w10 = ex * wy, b.w11 = wx * wy;
    return b;
}

// NCHW fp32 (the reference's tensor layout): one thread per pixel, channels looped; lanes = consecutive x so the
// four corner reads of a warp fall in a handful of 128-byte lines when the flow is smooth.
__global__ void __launch_bounds__(256) flow_warp_nchw_kernel(const float *__restrict__ x, const float *__restrict__ flow,
                                                             long long fs_n, long long fs_h, long long fs_w, long long fs_c,
                                                             float *__restrict__ y, int N, int C, int H, int W, int border) {
    const long long P = (long long)N * H * W;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (long long)gridDim.x * blockDim.x) {
        const int xw = (int)(p % W), yh = (int)((p / W) % H), n = (int)(p / ((long long)W * H));
        const float *f = flow + n * fs_n + yh * fs_h + xw * fs_w;
        const Bilin b = bilinear_setup(f[0], f[fs_c], xw, yh, W, H, border != 0);
        const long long o00 = (long long)b.y0 * W + b.x0;
        const float *xp = x + (long long)n * C * H * W;
        float *yp = y + (long long)n * C * H * W + (long long)yh * W + xw;
        for (int c = 0; c < C; ++c) {
            const float *pc = xp + (long long)c * H * W;
            float acc = 0.f;
            if (b.v00) acc += pc[o00] * b.w00;
            if (b.v01) acc += pc[o00 + 1] * b.w01;
            if (b.v10) acc += pc[o00 + W] * b.w10;
            if (b.v11) acc += pc[o00 + W + 1] * b.w11;
            yp[(long long)c * H * W] = acc;
        }
    }
}

cudaError_t launch_flow_warp_nchw(const float *x, const float *flow, long long fs_n, long long fs_h, long long fs_w,
                                  long long fs_c, float *y, int n, int c, int h, int w, int border, cudaStream_t st) {
    const long long P = (long long)n * h * w;
    if (P == 0 || c == 0) return cudaSuccess;
    long long blocks = (P + 255) / 256;
    const long long cap = (long long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    flow_warp_nchw_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, flow, fs_n, fs_h, fs_w, fs_c, y, n, c, h, w, border);
    return cudaGetLastError();
}

// NHWC (video path internal layout): a pixel's C channels are Q = C*esize/16 consecutive 16-byte vectors.  Q lanes
// cooperate on one pixel: the first lane of the group computes position and weights and warp-shuffles them to the
// other Q-1 lanes, then every lane gathers its 16-byte slice of the four corners and blends in fp32.
template <typename T, int Q>
__global__ void __launch_bounds__(256) flow_warp_nhwc_kernel(const T *__restrict__ x, const float *__restrict__ flow,
                                                             T *__restrict__ y, int N, int C, int H, int W, int border) {
    constexpr int VEC = 16 / sizeof(T);
    constexpr bool kShuffle = (Q & (Q - 1)) == 0 && Q <= 32;
    const long long P = (long long)N * H * W;
    const long long total = P * Q;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i - (threadIdx.x & 31) < total;
         i += (long long)gridDim.x * blockDim.x) {
        const bool active = i < total;
        const long long p = active ? i / Q : P - 1;
        const int q = (int)(i % Q);
        const int xw = (int)(p % W), yh = (int)((p / W) % H), n = (int)(p / ((long long)W * H));
        Bilin b = {};
        if (!kShuffle || q == 0) {
            const float *f = flow + ((long long)n * 2 * H + yh) * W + xw;
            b = bilinear_setup(f[0], f[(long long)H * W], xw, yh, W, H, border != 0);
        }
        if constexpr (kShuffle && Q > 1) {
            const int src = (threadIdx.x & 31) & ~(Q - 1);
            unsigned flags = (b.v00 ? 1u : 0u) | (b.v01 ? 2u : 0u) | (b.v10 ? 4u : 0u) | (b.v11 ? 8u : 0u);
            b.x0 = __shfl_sync(0xffffffffu, b.x0, src);
            b.y0 = __shfl_sync(0xffffffffu, b.y0, src);
            b.w00 = __shfl_sync(0xffffffffu, b.w00, src);
            b.w01 = __shfl_sync(0xffffffffu, b.w01, src);
            b.w10 = __shfl_sync(0xffffffffu, b.w10, src);
            b.w11 = __shfl_sync(0xffffffffu, b.w11, src);
            flags = __shfl_sync(0xffffffffu, flags, src);
            b.v00 = flags & 1u, b.v01 = flags & 2u, b.v10 = flags & 4u, b.v11 = flags & 8u;
        }
        if (!active) continue;
        const T *base = x + (((long long)n * H + b.y0) * W + b.x0) * C + q * VEC;
        float acc[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) acc[k] = 0.f;
        auto corner = [&](bool valid, long long off, float wgt) {
            if (!valid) return;
            const uint4 v = *reinterpret_cast<const uint4 *>(base + off);
            if constexpr (sizeof(T) == 4) {
                const float *fv = reinterpret_cast<const float *>(&v);
#pragma unroll
                for (int k = 0; k < 4; ++k) acc[k] += fv[k] * wgt;
            } else {
                const uint32_t *uv = reinterpret_cast<const uint32_t *>(&v);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float2 f2 = unpack_bf16x2(uv[k]);
                    acc[2 * k] += f2.x * wgt;
                    acc[2 * k + 1] += f2.y * wgt;
                }
            }
        };
        corner(b.v00, 0, b.w00);
        corner(b.v01, C, b.w01);
        corner(b.v10, (long long)W * C, b.w10);
        corner(b.v11, (long long)W * C + C, b.w11);
        uint4 o;
        if constexpr (sizeof(T) == 4) {
            o = *reinterpret_cast<uint4 *>(acc);
        } else {
            o.x = pack_bf16x2(acc[0], acc[1]);
            o.y = pack_bf16x2(acc[2], acc[3]);
            o.z = pack_bf16x2(acc[4], acc[5]);
            o.w = pack_bf16x2(acc[6], acc[7]);
        }
        *reinterpret_cast<uint4 *>(y + p * C + q * VEC) = o;
    }
}

template <typename T, int Q>
static cudaError_t warp_nhwc_t(const void *x, const float *flow, void *y, int n, int c, int h, int w, int border,
                               cudaStream_t st) {
    const long long total = (long long)n * h * w * Q;
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    flow_warp_nhwc_kernel<T, Q><<<(unsigned)blocks, 256, 0, st>>>((const T *)x, flow, (T *)y, n, c, h, w, border);
    return cudaGetLastError();
}

template <typename T>
static cudaError_t warp_nhwc_q(int Q, const void *x, const float *flow, void *y, int n, int c, int h, int w, int border,
                               cudaStream_t st) {
    switch (Q) {
        case 1: return warp_nhwc_t<T, 1>(x, flow, y, n, c, h, w, border, st);
        case 2: return warp_nhwc_t<T, 2>(x, flow, y, n, c, h, w, border, st);
        case 3: return warp_nhwc_t<T, 3>(x, flow, y, n, c, h, w, border, st);
        case 4: return warp_nhwc_t<T, 4>(x, flow, y, n, c, h, w, border, st);
        case 6: return warp_nhwc_t<T, 6>(x, flow, y, n, c, h, w, border, st);
        case 8: return warp_nhwc_t<T, 8>(x, flow, y, n, c, h, w, border, st);
        case 16: return warp_nhwc_t<T, 16>(x, flow, y, n, c, h, w, border, st);
    }
    return cudaErrorInvalidValue;
}

cudaError_t launch_flow_warp_nhwc(const void *x, const float *flow_nchw, void *y, int n, int c, int h, int w, int border,
                                  int dtype, cudaStream_t st) {
    if ((long long)n * h * w == 0) return cudaSuccess;
    if (dtype == kF32) {
        if (c % 4) return cudaErrorInvalidValue;
        return warp_nhwc_q<float>(c / 4, x, flow_nchw, y, n, c, h, w, border, st);
    }
    if (c % 8) return cudaErrorInvalidValue;
    return warp_nhwc_q<bf16>(c / 8, x, flow_nchw, y, n, c, h, w, border, st);
}

}  // namespace b200sr
