// conv.cuh -- generic stride-1 "same" convolution on NHWC activations, used by the video path:
//   SPyNet BasicModule 7x7 convs 8-32-64-32-16-2          models/spynet_arch.py:10-25
//   BasicVSR trunks 3x3 (67->64, 30 x [64->64, 64->64])    models/basicvsr_arch_origin.py:98-137
//   BasicVSR reconstruction 1x1 / 3x3 (+PixelShuffle(2))   models/basicvsr_arch_origin.py:84-90
//
// y[n, oy, ox, co] = act( bias[co] + sum_{ky,kx,ci} x[n, y+ky-P, x+kx-P, ci] * w[co][ci][ky][kx] ) (+ residual)
// with an optional PixelShuffle(2) folded into the store: (oy,ox,co') = (2y+i, 2x+j, co/4) for co = 4co'+2i+j.
// Activations may live inside wider NHWC tensors (channel stride / offset), so concatenations need no copies.
//
//  * conv_f32_kernel  : true-fp32 FFMA, register-tiled (4 px x 4 co per thread), for the 1e-4 parity arm
//  * conv_bf16_kernel : bf16 operands, fp32 accumulate, mma.sync m16n8k16 implicit GEMM (ldmatrix takes per-row
//                       addresses, so a filter tap is just a shifted window of the staged input tile)
#pragma once
#include "common.cuh"

namespace b200sr {

struct ConvArgs {
    const void *x;         // NHWC input,  channel stride x_cs, first channel x_co
    void *y;               // NHWC output, channel stride y_cs, first channel y_co
    const void *residual;  // optional, same geometry as y (added after the activation-free conv: x + conv(...))
    const void *w;         // packed filters (layout depends on the kernel)
    const float *bias;     // [coutp]
    int n, h, w_, cin, cinp, cout, coutp, x_cs, x_co, y_cs, y_co, r_cs, r_co, act, shuffle;
    int ks = 3;                       // filter size (the tcgen05 launcher picks its instantiation by it)
    int max_ctas = 0;                 // tcgen05 kernels: grid cap (0 = one CTA per SM), b200sr_conv_set_max_ctas
    // tcgen05 3x3 kernel, cout = 3 ("rgb" form, conv_last of BasicVSR_origin): y = fp32 NCHW image n at y + n * y_nstride,
    // y += bilinear x4 upsample (align_corners = False) of the fp32 NCHW low-resolution image n at base + n * base_nstride
    const float *base = nullptr;
    long long base_nstride = 0, y_nstride = 0;
    int x_planar = 0, y_planar = 0;   // tcgen05 kernel only: x (and the residual) / y in the planar-8 layout [n][c/8][h][w][8]
};

__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == 1) return fmaxf(v, 0.f);
    if (act == 2) return v > 0.f ? v : 0.1f * v;
    return v;
}

template <typename T>
__device__ __forceinline__ void store_out(const ConvArgs &a, T *y, const T *res, int n, int py, int px, int co, float v) {
    if (co >= a.cout) return;
    long long o;
    if (a.shuffle == 2) {
        const int c2 = co >> 2, i = (co >> 1) & 1, j = co & 1;
        o = (((long long)n * (2 * a.h) + (2 * py + i)) * (2 * a.w_) + (2 * px + j));
        co = c2;
    } else {
        o = (((long long)n * a.h + py) * a.w_ + px);
    }
    v = apply_act(v, a.act);
    if (res) v += to_f32<T>(res[o * a.r_cs + a.r_co + co]);
    y[o * a.y_cs + a.y_co + co] = from_f32<T>(v);
}

// ------------------------------------------------------------------------------------------------------------------
// fp32: CTA = 16 x 8 pixels x 32 output channels, 256 threads, thread = 4 pixels x 4 channels.
//   weights packed [K*K][cinp][coutp] fp32 (coutp multiple of 32, cinp multiple of 8)
// ------------------------------------------------------------------------------------------------------------------
template <int K, typename T>
__global__ void __launch_bounds__(256) conv_f32_kernel(ConvArgs a, int tiles_x, int tiles_y) {
    constexpr int TW = 16, TH = 8, CK = 8, CO = 32, P = K / 2;
    constexpr int HW_ = TW + K - 1, HH_ = TH + K - 1, HS = HW_ | 1;
    extern __shared__ __align__(16) float smem[];
    float *xs = smem;                  // [CK][HH_][HS]
    float *ws = xs + CK * HH_ * HS;    // [K*K][CK][CO]
    const int tid = threadIdx.x;
    const int tile = blockIdx.x, co0 = blockIdx.y * CO;
    const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, n = tile / (tiles_x * tiles_y);
    const int x0 = tx * TW, y0 = ty * TH;
    const int pxg = tid & 31, cog = tid >> 5;  // 32 pixel lanes x 8 channel groups of 4
    const T *x = reinterpret_cast<const T *>(a.x);
    const float *w = reinterpret_cast<const float *>(a.w);
    float acc[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[j][c] = a.bias[co0 + cog * 4 + c];
    int pofs[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int p = pxg + 32 * j;
        pofs[j] = (p / TW) * HS + (p % TW);
    }
    for (int c0 = 0; c0 < a.cinp; c0 += CK) {
        __syncthreads();
        for (int i = tid; i < HH_ * HW_ * CK; i += 256) {
            const int c = i % CK, q = (i / CK) % HW_, r = i / (CK * HW_);
            const int gy = y0 - P + r, gx = x0 - P + q;
            float v = 0.f;
            if (gy >= 0 && gy < a.h && gx >= 0 && gx < a.w_ && c0 + c < a.cin)
                v = to_f32<T>(x[(((long long)n * a.h + gy) * a.w_ + gx) * a.x_cs + a.x_co + c0 + c]);
            xs[(c * HH_ + r) * HS + q] = v;
        }
        for (int i = tid; i < K * K * CK * (CO / 4); i += 256) {
            const int q = i % (CO / 4), c = (i / (CO / 4)) % CK, tap = i / ((CO / 4) * CK);
            reinterpret_cast<float4 *>(ws)[i] =
                *reinterpret_cast<const float4 *>(w + ((size_t)tap * a.cinp + c0 + c) * a.coutp + co0 + 4 * q);
        }
        __syncthreads();
#pragma unroll 1
        for (int ky = 0; ky < K; ++ky)
#pragma unroll
            for (int kx = 0; kx < K; ++kx)
#pragma unroll
                for (int c = 0; c < CK; ++c) {
                    const float4 wv = *reinterpret_cast<const float4 *>(ws + ((ky * K + kx) * CK + c) * CO + cog * 4);
                    const float *xp = xs + (c * HH_ + ky) * HS + kx;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float v = xp[pofs[j]];
                        acc[j][0] = fmaf(v, wv.x, acc[j][0]);
                        acc[j][1] = fmaf(v, wv.y, acc[j][1]);
                        acc[j][2] = fmaf(v, wv.z, acc[j][2]);
                        acc[j][3] = fmaf(v, wv.w, acc[j][3]);
                    }
                }
    }
    T *y = reinterpret_cast<T *>(a.y);
    const T *res = reinterpret_cast<const T *>(a.residual);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int p = pxg + 32 * j, py = y0 + p / TW, px = x0 + p % TW;
        if (py < a.h && px < a.w_)
#pragma unroll
            for (int c = 0; c < 4; ++c) store_out<T>(a, y, res, n, py, px, co0 + cog * 4 + c, acc[j][c]);
    }
}

template <int K>
constexpr size_t conv_f32_smem() {
    return sizeof(float) * (size_t)(8 * (8 + K - 1) * ((16 + K - 1) | 1) + K * K * 8 * 32);
}

// ------------------------------------------------------------------------------------------------------------------
// bf16 tensor cores: CTA = 16 x 8 pixels x (8*NT) output channels, 4 warps, warp = 2 pixel rows (two m16 tiles).
//   weights packed [K*K][coutp][cinp] bf16 (K-major rows for ldmatrix), cinp multiple of 16, coutp multiple of 8*NT.
//   Input channels are consumed 16 at a time: the halo tile of the chunk is staged [pixel][16 ch] (48-byte pixel
//   stride -> conflict-free ldmatrix), the filter slice of one kernel row [kx][co][16 ch] likewise.
// ------------------------------------------------------------------------------------------------------------------
template <int K, int NT, typename TIN, typename TOUT>
__global__ void __launch_bounds__(128) conv_bf16_kernel(ConvArgs a, int tiles_x, int tiles_y) {
    constexpr int TW = 16, TH = 8, P = K / 2, HW_ = TW + K - 1, HH_ = TH + K - 1, PS = 24;  // PS: smem row stride (elements)
    extern __shared__ __align__(128) uint8_t smem_raw[];
    bf16 *xs = reinterpret_cast<bf16 *>(smem_raw);  // [HH_*HW_][PS]
    bf16 *ws = xs + HH_ * HW_ * PS;                 // [K][8*NT][PS]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const int tile = blockIdx.x, co0 = blockIdx.y * (8 * NT);
    const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, n = tile / (tiles_x * tiles_y);
    const int x0 = tx * TW, y0 = ty * TH;
    const TIN *x = reinterpret_cast<const TIN *>(a.x);
    const bf16 *w = reinterpret_cast<const bf16 *>(a.w);
    const uint32_t xs_u = smem_u32(xs), ws_u = smem_u32(ws);
    float acc[2][NT][4];
#pragma unroll
    for (int mb = 0; mb < 2; ++mb)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const float b0 = a.bias[co0 + nt * 8 + 2 * t], b1 = a.bias[co0 + nt * 8 + 2 * t + 1];
            acc[mb][nt][0] = b0, acc[mb][nt][1] = b1, acc[mb][nt][2] = b0, acc[mb][nt][3] = b1;
        }
    for (int c0 = 0; c0 < a.cinp; c0 += 16) {
        __syncthreads();
        // stage the halo tile of this channel chunk (bf16), zero outside the image / beyond cin
        for (int i = tid; i < HH_ * HW_ * 2; i += 128) {
            const int half = i & 1, hp = i >> 1, r = hp / HW_, q = hp % HW_;
            const int gy = y0 - P + r, gx = x0 - P + q;
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            const int cb = c0 + half * 8;
            if (gy >= 0 && gy < a.h && gx >= 0 && gx < a.w_ && cb < a.cin) {
                const TIN *src = x + (((long long)n * a.h + gy) * a.w_ + gx) * a.x_cs + a.x_co + cb;
                if constexpr (sizeof(TIN) == 2) {
                    if (cb + 8 <= a.cin && ((a.x_cs | a.x_co | cb) & 7) == 0) {
                        v = *reinterpret_cast<const uint4 *>(src);
                    } else {
                        uint32_t *vw = reinterpret_cast<uint32_t *>(&v);
                        for (int e = 0; e < 8; e += 2)
                            vw[e / 2] = pack_bf16x2(cb + e < a.cin ? to_f32<TIN>(src[e]) : 0.f, cb + e + 1 < a.cin ? to_f32<TIN>(src[e + 1]) : 0.f);
                    }
                } else {
                    uint32_t *vw = reinterpret_cast<uint32_t *>(&v);
                    for (int e = 0; e < 8; e += 2)
                        vw[e / 2] = pack_bf16x2(cb + e < a.cin ? to_f32<TIN>(src[e]) : 0.f, cb + e + 1 < a.cin ? to_f32<TIN>(src[e + 1]) : 0.f);
                }
            }
            *reinterpret_cast<uint4 *>(xs + hp * PS + half * 8) = v;
        }
#pragma unroll 1
        for (int ky = 0; ky < K; ++ky) {
            if (ky > 0) __syncthreads();
            for (int i = tid; i < K * 8 * NT * 2; i += 128) {
                const int half = i & 1, co = (i >> 1) % (8 * NT), kx = (i >> 1) / (8 * NT);
                *reinterpret_cast<uint4 *>(ws + (kx * 8 * NT + co) * PS + half * 8) =
                    *reinterpret_cast<const uint4 *>(w + ((size_t)(ky * K + kx) * a.coutp + co0 + co) * a.cinp + c0 + half * 8);
            }
            __syncthreads();
#pragma unroll
            for (int kx = 0; kx < K; ++kx) {
                uint32_t bfr[NT][2];
#pragma unroll
                for (int nt = 0; nt + 1 < NT; nt += 2)
                    ldmatrix_x4(bfr[nt][0], bfr[nt][1], bfr[nt + 1][0], bfr[nt + 1][1],
                                ws_u + ((kx * 8 * NT + (nt + (lane >> 4)) * 8 + (lane & 7)) * PS + ((lane >> 3) & 1) * 8) * 2);
                if constexpr (NT % 2 == 1)
                    ldmatrix_x2(bfr[NT - 1][0], bfr[NT - 1][1], ws_u + ((kx * 8 * NT + (NT - 1) * 8 + (lane & 7)) * PS + ((lane >> 3) & 1) * 8) * 2);
#pragma unroll
                for (int mb = 0; mb < 2; ++mb) {
                    uint32_t af[4];
                    const int hp = (warp * 2 + mb + ky) * HW_ + kx + (lane & 15);
                    ldmatrix_x4(af[0], af[1], af[2], af[3], xs_u + (hp * PS + (lane >> 4) * 8) * 2);
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) mma_16816(acc[mb][nt], af[0], af[1], af[2], af[3], bfr[nt][0], bfr[nt][1]);
                }
            }
        }
    }
    TOUT *y = reinterpret_cast<TOUT *>(a.y);
    const TOUT *res = reinterpret_cast<const TOUT *>(a.residual);
#pragma unroll
    for (int mb = 0; mb < 2; ++mb) {
        const int py = y0 + warp * 2 + mb;
        if (py >= a.h) continue;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int px = x0 + g + half * 8;
            if (px >= a.w_) continue;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                store_out<TOUT>(a, y, res, n, py, px, co0 + nt * 8 + 2 * t, acc[mb][nt][half * 2]);
                store_out<TOUT>(a, y, res, n, py, px, co0 + nt * 8 + 2 * t + 1, acc[mb][nt][half * 2 + 1]);
            }
        }
    }
}

template <int K, int NT>
constexpr size_t conv_bf16_smem() {
    return (size_t)2 * 24 * ((8 + K - 1) * (16 + K - 1) + K * 8 * NT);
}

}  // namespace b200sr
