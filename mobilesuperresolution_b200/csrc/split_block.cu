// split_block.cu -- the fork's searchable block body, Split_Block.forward_body (models/wdsr_b.py:482-496), as ONE kernel:
//
//   x1 = e * x                      e = BinaryConv2d(least_channel = 0) forward weight, w - (w - m), per channel (models/ops.py:18-26)
//   x2 = x - x1
//   x3 = x2 + sum_k p_k * relu(PW_k(relu(DW_k(x1)))) + x1        k = 3, 5, 7; DW_k depthwise k x k, PW_k 1x1 (Conv_sep, :375-402),
//   y  = x2 + e * x3                                             both weight-normed (folded on the host); p = softmax(alpha)
//
// The reference runs it as 3 x (depthwise conv, ReLU, 1x1 conv, ReLU, scale, add) + 2 masks + 3 adds = ~25 kernels with a full
// tensor round trip each; here a CTA stages a 32 x 8 pixel tile of x1 (+ 3-pixel halo, zero outside the image = the convs' padding)
// for all C channels in shared memory and one thread carries one pixel through the whole body in registers (fp32 FFMA: this is the
// 1e-4 parity arm as well as the bf16-storage arm).  Tensors are the reference's NCHW (depthwise convs are per plane).
#include <cstring>

#include "common.cuh"
#include "launch.h"

namespace b200sr {

namespace splitcfg {
constexpr int TW = 32, TH = 8, HALO = 3, SW = TW + 2 * HALO, SH = TH + 2 * HALO, NTHREADS = 256;
__host__ __device__ constexpr int dw_floats(int C) { return C * (9 + 25 + 49); }
// parameter pack (floats): dw3[C][9] | dw5[C][25] | dw7[C][49] | dw_bias[3][C] | pw[3][C in][C out] | pw_bias[3][C] | e[C] | p[4] | g[C]
// g = per-channel multiplier applied to x FIRST (all ones for a stand-alone Split_Block): NAS_MODEL.forward runs `y = self.mask(y)` in
// front of every block (models/wdsr_b.py:116-119), a BinaryConv2d multiply that is fused here instead of costing a tensor round trip
__host__ __device__ constexpr int param_floats(int C) { return dw_floats(C) + 3 * C + 3 * C * C + 3 * C + C + 4 + C; }
__host__ __device__ constexpr size_t smem_bytes(int C) { return (size_t)(C * SH * SW + C * TW * TH) * 4; }
}  // namespace splitcfg

// One branch k of the body for a 32 x 8 pixel tile, in two phases through shared memory:
//   depthwise: a work item = (channel c, tile row r, 8-pixel segment g); a warp takes the 32 items of ONE channel, so the filter taps
//              are warp-uniform constant-bank reads and every staged value loaded feeds up to K taps of 8 pixels (sliding window):
//              K + 7 loads per 8 K FMAs.  ReLU(d) -> ds[c][pixel].
//   pointwise: a thread = a pixel: t[o] = b[o] + sum_c pw[c][o] * ds[c][pixel] with both loops unrolled -- the 1x1 filter is an
//              immediate constant-bank operand of each FFMA -- then s[o] += p * ReLU(t[o])          (models/wdsr_b.py:491-493)
template <int C, int K>
__device__ __forceinline__ void split_branch(const float *__restrict__ xs, float *__restrict__ ds, const float *__restrict__ dw,
                                             const float *__restrict__ dwb, const float *__restrict__ pw, const float *__restrict__ pwb,
                                             float p, int tid, float (&s)[C]) {
    using namespace splitcfg;
    constexpr int P = K / 2;
    const int lane = tid & 31, warp = tid >> 5, r = lane >> 2, g = lane & 3;
    __syncthreads();                       // the previous branch has consumed ds
#pragma unroll 1
    for (int c = warp; c < C; c += NTHREADS / 32) {
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = dwb[c];
        const float *xp = xs + (c * SH + r + HALO - P) * SW + 8 * g + HALO - P;
        const float *wp = dw + c * K * K;
#pragma unroll
        for (int ky = 0; ky < K; ++ky) {
            float v[8 + K - 1];
#pragma unroll
            for (int j = 0; j < 8 + K - 1; ++j) v[j] = xp[ky * SW + j];
#pragma unroll
            for (int kx = 0; kx < K; ++kx) {
                const float wv = wp[ky * K + kx];
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] = fmaf(v[j + kx], wv, acc[j]);
            }
        }
        float *dp = ds + c * (TW * TH) + r * TW + 8 * g;
        *reinterpret_cast<float4 *>(dp) = make_float4(fmaxf(acc[0], 0.f), fmaxf(acc[1], 0.f), fmaxf(acc[2], 0.f), fmaxf(acc[3], 0.f));
        *reinterpret_cast<float4 *>(dp + 4) = make_float4(fmaxf(acc[4], 0.f), fmaxf(acc[5], 0.f), fmaxf(acc[6], 0.f), fmaxf(acc[7], 0.f));
    }
    __syncthreads();
    float t[C];
#pragma unroll
    for (int o = 0; o < C; ++o) t[o] = pwb[o];
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const float d = ds[c * (TW * TH) + tid];
#pragma unroll
        for (int o = 0; o < C; ++o) t[o] = fmaf(d, pw[c * C + o], t[o]);
    }
#pragma unroll
    for (int o = 0; o < C; ++o) s[o] = s[o] + fmaxf(t[o], 0.f) * p;   // x3 = x3 + x_ * pro[i]
}

// The folded filters travel as a KERNEL ARGUMENT (16-24 KB, CUDA >= 12.1 allows 32 KB): they sit in the constant bank, so with the
// channel loop unrolled every filter tap is an immediate constant operand of its FFMA -- no load instruction and no shared-memory
// bandwidth for weights (the first form kept them in shared memory: two LDS per FFMA, shared-memory bound at 198 us per 360p frame).
template <int C> struct SplitParams { float v[splitcfg::param_floats(C)]; };

template <int C, typename T>
__global__ void __launch_bounds__(splitcfg::NTHREADS, 3) split_block_kernel(const T *__restrict__ x, T *__restrict__ y,
                                                                         const __grid_constant__ SplitParams<C> prm_, int N, int H, int W,
                                                                         int tiles_x, int tiles_y) {
    using namespace splitcfg;
    extern __shared__ __align__(16) float smem[];
    float *xs = smem;                             // [C][SH][SW]  x1 = e * x, zero outside the image
    float *ds = smem + C * SH * SW;               // [C][TH * TW]  ReLU(DW_k(x1)) of the current branch
    const int tid = threadIdx.x;
    const int tile = blockIdx.x;
    const int x0 = (tile % tiles_x) * TW, y0 = ((tile / tiles_x) % tiles_y) * TH, n = tile / (tiles_x * tiles_y);
    const float *prm = prm_.v;
    const float *dw3 = prm, *dw5 = dw3 + C * 9, *dw7 = dw5 + C * 25, *dwb = dw7 + C * 49, *pw = dwb + 3 * C, *pwb = pw + 3 * C * C,
                *e = pwb + 3 * C, *p = e + C, *gmask = p + 4;
    const long long plane = (long long)H * W;
    const T *xn = x + (long long)n * C * plane;
    for (int i = tid; i < C * SH * SW; i += NTHREADS) {
        const int q = i % SW, r = (i / SW) % SH, c = i / (SW * SH);
        const int gy = y0 - HALO + r, gx = x0 - HALO + q;
        float v = 0.f;
        if (gy >= 0 && gy < H && gx >= 0 && gx < W) v = (to_f32<T>(xn[c * plane + (long long)gy * W + gx]) * gmask[c]) * e[c];
        xs[i] = v;
    }
    const int tx = tid % TW, ty = tid / TW, gx = x0 + tx, gy = y0 + ty;
    const bool ok = gx < W && gy < H;             // every thread stays for the barriers of the branches
    float s[C];
    const T *xc = xn + (long long)(ok ? gy : 0) * W + (ok ? gx : 0);
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const float xv = to_f32<T>(xc[c * plane]) * gmask[c];
        s[c] = xv - xv * e[c];      // x3 = clone(x2), x2 = x - x1, x1 = e * x
    }
    split_branch<C, 3>(xs, ds, dw3, dwb, pw, pwb, p[0], tid, s);
    split_branch<C, 5>(xs, ds, dw5, dwb + C, pw + C * C, pwb + C, p[1], tid, s);
    split_branch<C, 7>(xs, ds, dw7, dwb + 2 * C, pw + 2 * C * C, pwb + 2 * C, p[2], tid, s);
    if (!ok) return;
    T *yn = y + (long long)n * C * plane + (long long)gy * W + gx;
#pragma unroll
    for (int c = 0; c < C; ++c) {   // x3 += x1; y = x2 + split(x3)   (x re-read: an L1 hit, cheaper than 2 C live registers)
        const float xv = to_f32<T>(xc[c * plane]) * gmask[c], x1 = xv * e[c], x2 = xv - x1;
        yn[c * plane] = from_f32<T>(x2 + (s[c] + x1) * e[c]);
    }
}

template <int C, typename T>
static cudaError_t split_t(const void *x, void *y, const float *params_host, int N, int H, int W, cudaStream_t st) {
    using namespace splitcfg;
    auto kern = split_block_kernel<C, T>;
    static thread_local bool set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes(C));
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) set[dev] = true;
    }
    const int tx = ceil_div(W, TW), ty = ceil_div(H, TH);
    SplitParams<C> prm;
    memcpy(prm.v, params_host, sizeof prm.v);
    kern<<<tx * ty * N, NTHREADS, smem_bytes(C), st>>>((const T *)x, (T *)y, prm, N, H, W, tx, ty);
    return cudaGetLastError();
}

int split_param_floats(int C) { return splitcfg::param_floats(C); }

cudaError_t launch_split_block(int C, int dtype, const void *x, void *y, const float *params, int N, int H, int W, cudaStream_t st) {
    // params: HOST pointer to the packed image (passed to the kernel by value)
#define B200SR_SPLIT_CASE(CC)                                                       \
    if (C == CC) return dtype == kF32 ? split_t<CC, float>(x, y, params, N, H, W, st) \
                                      : split_t<CC, bf16>(x, y, params, N, H, W, st);
    B200SR_SPLIT_CASE(8)
    B200SR_SPLIT_CASE(16)
    B200SR_SPLIT_CASE(24)
    B200SR_SPLIT_CASE(32)
#undef B200SR_SPLIT_CASE
    return cudaErrorInvalidValue;
}

}  // namespace b200sr
