// wdsr_f32.cuh -- true-fp32 (FFMA) kernels of the WDSR-B forward: head, fused residual block, fused tail.
//
// This is the fp32 arm the north star gates at max-abs 1e-4: single-pass TF32 tensor-core operands miss that
// gate by 10-400x, so the arithmetic is plain fp32 FMA with fp32 accumulation.
//
// Layouts: network input/output NCHW (the reference's tensors); trunk NHWC with CP channels (IN padded to a
// multiple of 8, pad channels are zero).  One CTA = one spatial tile; the tile (+halo) is staged in shared memory
// channel-planar so that lanes (= consecutive pixels) read consecutive banks, and the filter taps are broadcast
// reads of 16-byte weight vectors.
#pragma once
#include "common.cuh"

namespace b200sr {

// ------------------------------------------------------------------------------------------------------------------
// head: trunk = conv3x3(x - mean, Wh) + bh            models/basic_wdsr_b.py:86-87
//   x NCHW (3 channels), zero padding applied AFTER the mean subtraction (pads are 0 in the x-mean domain).
//   wpack: [27][CP] (k = c*9 + ky*3 + kx) then bias[CP], fp32.
// ------------------------------------------------------------------------------------------------------------------
template <typename TIN, typename TOUT, int CP>
__global__ void __launch_bounds__(256) wdsr_head_kernel(const TIN *__restrict__ x, TOUT *__restrict__ trunk,
                                                        const float *__restrict__ wpack, int N, int H, int W, float mean,
                                                        int tiles_x, int tiles_y) {
    // one CTA = 64 x 4 pixels; the (x - mean) halo tile is staged once (coalesced NCHW reads), the NHWC result is staged
    // in shared memory and written back as contiguous 16-byte vectors (one tile row = 64 * CP * sizeof(TOUT) contiguous bytes)
    constexpr int TW = 64, TH = 4, HW_ = TW + 2, HH_ = TH + 2, HS = HW_ + 1;
    __shared__ __align__(16) float ws[27 * CP + CP];
    __shared__ float xin[3][HH_][HS];
    __shared__ __align__(16) TOUT ostage[TW * TH * CP];
    const int tid = threadIdx.x;
    for (int i = tid; i < 27 * CP + CP; i += 256) ws[i] = wpack[i];
    const int tile = blockIdx.x;
    const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, n = tile / (tiles_x * tiles_y);
    const int x0 = tx * TW, y0 = ty * TH;
    for (int i = tid; i < 3 * HH_ * HW_; i += 256) {
        const int c = i / (HH_ * HW_), r = (i / HW_) % HH_, q = i % HW_;
        const int gy = y0 - 1 + r, gx = x0 - 1 + q;
        float v = 0.f;
        if (gy >= 0 && gy < H && gx >= 0 && gx < W) v = to_f32<TIN>(x[(((long long)n * 3 + c) * H + gy) * W + gx]) - mean;
        xin[c][r][q] = v;
    }
    __syncthreads();
    const int lx = tid % TW, ly = tid / TW;
    float acc[CP];
#pragma unroll
    for (int c = 0; c < CP; ++c) acc[c] = ws[27 * CP + c];
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const float v = xin[c][ly + ky][lx + kx];
                const float4 *wr = reinterpret_cast<const float4 *>(&ws[(c * 9 + ky * 3 + kx) * CP]);
#pragma unroll
                for (int q = 0; q < CP / 4; ++q) {
                    const float4 wv = wr[q];
                    acc[4 * q + 0] = fmaf(v, wv.x, acc[4 * q + 0]);
                    acc[4 * q + 1] = fmaf(v, wv.y, acc[4 * q + 1]);
                    acc[4 * q + 2] = fmaf(v, wv.z, acc[4 * q + 2]);
                    acc[4 * q + 3] = fmaf(v, wv.w, acc[4 * q + 3]);
                }
            }
    TOUT *o = ostage + tid * CP;
    if constexpr (sizeof(TOUT) == 4) {
#pragma unroll
        for (int q = 0; q < CP / 4; ++q)
            reinterpret_cast<float4 *>(o)[q] = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
    } else {
#pragma unroll
        for (int q = 0; q < CP / 8; ++q) {
            uint4 v;
            v.x = pack_bf16x2(acc[8 * q + 0], acc[8 * q + 1]);
            v.y = pack_bf16x2(acc[8 * q + 2], acc[8 * q + 3]);
            v.z = pack_bf16x2(acc[8 * q + 4], acc[8 * q + 5]);
            v.w = pack_bf16x2(acc[8 * q + 6], acc[8 * q + 7]);
            reinterpret_cast<uint4 *>(o)[q] = v;
        }
    }
    __syncthreads();
    constexpr int VPP = CP * (int)sizeof(TOUT) / 16;  // 16-byte vectors per pixel
    for (int i = tid; i < TW * TH * VPP; i += 256) {
        const int p = i / VPP, q = i % VPP, px = p % TW, py = p / TW;
        const int gx = x0 + px, gy = y0 + py;
        if (gx < W && gy < H)
            reinterpret_cast<uint4 *>(trunk + (((long long)n * H + gy) * W + gx) * CP)[q] = reinterpret_cast<const uint4 *>(ostage + p * CP)[q];
    }
}

// ------------------------------------------------------------------------------------------------------------------
// fused residual block, fp32:  out = x + conv3x3(conv1x1(relu(conv1x1(x))))      models/basic_wdsr_b.py:96-144
//   The 3x3 zero-pads the REDUCE OUTPUT t2 (not the trunk): t2 of out-of-image halo pixels is forced to 0
//   (SURVEY.md 0-5ii); evaluating the 1x1s on a zero trunk there would give relu(b1)*W2+b2 != 0.
//   wpack (fp32): W1[M1][CP] | b1[M1] | W2[M1][M2P] | b2[M2P] | W3[9][M2P][CP] | b3[CP]
// ------------------------------------------------------------------------------------------------------------------
struct BlockF32Layout {
    int w1, b1, w2, b2, w3, b3, total;  // float offsets into wpack
    __host__ __device__ BlockF32Layout(int CP, int M1, int M2P) {
        w1 = 0;
        b1 = w1 + M1 * CP;
        w2 = b1 + M1;
        b2 = w2 + M1 * M2P;
        w3 = b2 + M2P;
        b3 = w3 + 9 * M2P * CP;
        total = b3 + CP;
    }
};

template <int CP, int M2P, int TW, int TH>
__global__ void __launch_bounds__(TW *TH) wdsr_block_f32_kernel(const float *__restrict__ in, float *__restrict__ out,
                                                                const float *__restrict__ wpack, int M1, int N, int H,
                                                                int W, int tiles_x, int tiles_y) {
    constexpr int HW_ = TW + 2, HH_ = TH + 2, HP = HW_ * HH_;
    constexpr int HPS = HP + 1;  // planar stride
    extern __shared__ __align__(16) float smem[];
    const BlockF32Layout L(CP, M1, M2P);
    float *wsm = smem;                       // weights
    float *xs = wsm + round_up(L.total, 4);  // [CP][HPS]
    float *t2s = xs + CP * HPS;              // [M2P][HPS]
    const int nthreads = TW * TH;
    const int tid = threadIdx.x;

    for (int i = tid; i < L.total; i += nthreads) wsm[i] = wpack[i];

    const int tile = blockIdx.x;
    const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, n = tile / (tiles_x * tiles_y);
    const int x0 = tx * TW - 1, y0 = ty * TH - 1;

    // stage trunk tile + halo, planar; out-of-image -> 0
    for (int i = tid; i < HP * (CP / 4); i += nthreads) {
        const int hp = i / (CP / 4), q = i % (CP / 4);
        const int gy = y0 + hp / HW_, gx = x0 + hp % HW_;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gy >= 0 && gy < H && gx >= 0 && gx < W)
            v = *reinterpret_cast<const float4 *>(in + (((long long)n * H + gy) * W + gx) * CP + 4 * q);
        xs[(4 * q + 0) * HPS + hp] = v.x;
        xs[(4 * q + 1) * HPS + hp] = v.y;
        xs[(4 * q + 2) * HPS + hp] = v.z;
        xs[(4 * q + 3) * HPS + hp] = v.w;
    }
    __syncthreads();

    // phase 1: t2 = W2 * relu(W1 * x + b1) + b2 on every halo pixel
    for (int hp = tid; hp < HP; hp += nthreads) {
        const int gy = y0 + hp / HW_, gx = x0 + hp % HW_;
        const bool inside = gy >= 0 && gy < H && gx >= 0 && gx < W;
        float acc2[M2P];
#pragma unroll
        for (int j = 0; j < M2P; ++j) acc2[j] = wsm[L.b2 + j];
        if (inside) {
            float xv[CP];
#pragma unroll
            for (int c = 0; c < CP; ++c) xv[c] = xs[c * HPS + hp];
            for (int m = 0; m < M1; ++m) {
                const float4 *w1r = reinterpret_cast<const float4 *>(&wsm[L.w1 + m * CP]);
                float t = wsm[L.b1 + m];
#pragma unroll
                for (int q = 0; q < CP / 4; ++q) {
                    float4 wv = w1r[q];
                    t = fmaf(xv[4 * q + 0], wv.x, t);
                    t = fmaf(xv[4 * q + 1], wv.y, t);
                    t = fmaf(xv[4 * q + 2], wv.z, t);
                    t = fmaf(xv[4 * q + 3], wv.w, t);
                }
                t = fmaxf(t, 0.f);
                const float4 *w2r = reinterpret_cast<const float4 *>(&wsm[L.w2 + m * M2P]);
#pragma unroll
                for (int q = 0; q < M2P / 4; ++q) {
                    float4 wv = w2r[q];
                    acc2[4 * q + 0] = fmaf(t, wv.x, acc2[4 * q + 0]);
                    acc2[4 * q + 1] = fmaf(t, wv.y, acc2[4 * q + 1]);
                    acc2[4 * q + 2] = fmaf(t, wv.z, acc2[4 * q + 2]);
                    acc2[4 * q + 3] = fmaf(t, wv.w, acc2[4 * q + 3]);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < M2P; ++j) t2s[j * HPS + hp] = inside ? acc2[j] : 0.f;
    }
    __syncthreads();

    // phase 2: out = x + conv3x3(t2) + b3 on the TWxTH interior
    {
        const int lx = tid % TW, ly = tid / TW;
        const int gx = x0 + 1 + lx, gy = y0 + 1 + ly;
        if (gx < W && gy < H) {
            const int hc = (ly + 1) * HW_ + (lx + 1);
            float acc[CP];
#pragma unroll
            for (int c = 0; c < CP; ++c) acc[c] = wsm[L.b3 + c] + xs[c * HPS + hc];
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
                const int hp = hc + (tap / 3 - 1) * HW_ + (tap % 3 - 1);
#pragma unroll 4
                for (int j = 0; j < M2P; ++j) {
                    const float v = t2s[j * HPS + hp];
                    const float4 *w3r = reinterpret_cast<const float4 *>(&wsm[L.w3 + (tap * M2P + j) * CP]);
#pragma unroll
                    for (int q = 0; q < CP / 4; ++q) {
                        float4 wv = w3r[q];
                        acc[4 * q + 0] = fmaf(v, wv.x, acc[4 * q + 0]);
                        acc[4 * q + 1] = fmaf(v, wv.y, acc[4 * q + 1]);
                        acc[4 * q + 2] = fmaf(v, wv.z, acc[4 * q + 2]);
                        acc[4 * q + 3] = fmaf(v, wv.w, acc[4 * q + 3]);
                    }
                }
            }
            float *o = out + (((long long)n * H + gy) * W + gx) * CP;
#pragma unroll
            for (int q = 0; q < CP / 4; ++q)
                reinterpret_cast<float4 *>(o)[q] = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
        }
    }
}

template <int CP, int M2P, int TW, int TH>
inline size_t wdsr_block_f32_smem(int M1) {
    constexpr int HPS = (TW + 2) * (TH + 2) + 1;
    BlockF32Layout L(CP, M1, M2P);
    return sizeof(float) * (size_t)(round_up(L.total, 4) + (CP + M2P) * HPS);
}

// ------------------------------------------------------------------------------------------------------------------
// fused tail, fp32:   out = PixelShuffle_s( conv3x3(trunk, Wt) + bt + conv5x5(x - mean, Ws) + bs ) (+ mean)
//   models/basic_wdsr_b.py:90-92.   wpack: Wt[9][CP][NO] | Ws[75][NO] (k = c*25+ky*5+kx) | bias[NO] (= bt+bs)
//   NO = 3*s*s padded to a multiple of 4.
// ------------------------------------------------------------------------------------------------------------------
template <typename TIN, typename TOUT, int CP, int S, int TW, int TH>
__global__ void __launch_bounds__(TW *TH) wdsr_tail_f32_kernel(const float *__restrict__ trunk, const TIN *__restrict__ x,
                                                               TOUT *__restrict__ y, const float *__restrict__ wpack, int N,
                                                               int H, int W, int tiles_x, int tiles_y, float mean,
                                                               float out_add) {
    constexpr int NO = round_up(3 * S * S, 4);
    constexpr int HW1 = TW + 2, HP1 = HW1 * (TH + 2), HPS1 = HP1 + 1;
    constexpr int HW2 = TW + 4, HP2 = HW2 * (TH + 4), HPS2 = HP2 + 1;
    constexpr int WT = 9 * CP * NO, WS = 75 * NO;
    extern __shared__ __align__(16) float smem[];
    float *wsm = smem;               // WT + WS + NO
    float *ts = wsm + WT + WS + NO;  // [CP][HPS1]
    float *xs = ts + CP * HPS1;      // [3][HPS2]
    const int nthreads = TW * TH, tid = threadIdx.x;
    for (int i = tid; i < WT + WS + NO; i += nthreads) wsm[i] = wpack[i];

    const int tile = blockIdx.x;
    const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, n = tile / (tiles_x * tiles_y);
    const int x0 = tx * TW, y0 = ty * TH;
    for (int i = tid; i < HP1 * (CP / 4); i += nthreads) {
        const int hp = i / (CP / 4), q = i % (CP / 4);
        const int gy = y0 - 1 + hp / HW1, gx = x0 - 1 + hp % HW1;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gy >= 0 && gy < H && gx >= 0 && gx < W)
            v = *reinterpret_cast<const float4 *>(trunk + (((long long)n * H + gy) * W + gx) * CP + 4 * q);
        ts[(4 * q + 0) * HPS1 + hp] = v.x;
        ts[(4 * q + 1) * HPS1 + hp] = v.y;
        ts[(4 * q + 2) * HPS1 + hp] = v.z;
        ts[(4 * q + 3) * HPS1 + hp] = v.w;
    }
    for (int i = tid; i < 3 * HP2; i += nthreads) {
        const int c = i / HP2, hp = i % HP2;
        const int gy = y0 - 2 + hp / HW2, gx = x0 - 2 + hp % HW2;
        float v = 0.f;
        if (gy >= 0 && gy < H && gx >= 0 && gx < W) v = to_f32<TIN>(x[(((long long)n * 3 + c) * H + gy) * W + gx]) - mean;
        xs[c * HPS2 + hp] = v;
    }
    __syncthreads();

    const int lx = tid % TW, ly = tid / TW;
    const int gx = x0 + lx, gy = y0 + ly;
    if (gx >= W || gy >= H) return;
    float acc[NO];
#pragma unroll
    for (int o = 0; o < NO; ++o) acc[o] = wsm[WT + WS + o];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
        const int hp = (ly + tap / 3) * HW1 + (lx + tap % 3);
#pragma unroll 2
        for (int c = 0; c < CP; ++c) {
            const float v = ts[c * HPS1 + hp];
            const float4 *wr = reinterpret_cast<const float4 *>(&wsm[(tap * CP + c) * NO]);
#pragma unroll
            for (int q = 0; q < NO / 4; ++q) {
                float4 wv = wr[q];
                acc[4 * q + 0] = fmaf(v, wv.x, acc[4 * q + 0]);
                acc[4 * q + 1] = fmaf(v, wv.y, acc[4 * q + 1]);
                acc[4 * q + 2] = fmaf(v, wv.z, acc[4 * q + 2]);
                acc[4 * q + 3] = fmaf(v, wv.w, acc[4 * q + 3]);
            }
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll 5
        for (int tap = 0; tap < 25; ++tap) {
            const float v = xs[c * HPS2 + (ly + tap / 5) * HW2 + (lx + tap % 5)];
            const float4 *wr = reinterpret_cast<const float4 *>(&wsm[WT + (c * 25 + tap) * NO]);
#pragma unroll
            for (int q = 0; q < NO / 4; ++q) {
                float4 wv = wr[q];
                acc[4 * q + 0] = fmaf(v, wv.x, acc[4 * q + 0]);
                acc[4 * q + 1] = fmaf(v, wv.y, acc[4 * q + 1]);
                acc[4 * q + 2] = fmaf(v, wv.z, acc[4 * q + 2]);
                acc[4 * q + 3] = fmaf(v, wv.w, acc[4 * q + 3]);
            }
        }
    // PixelShuffle store: out[n, c, S*gy+i, S*gx+j] = acc[c*S*S + i*S + j] + out_add
    const int OH = S * H, OW = S * W;
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int i = 0; i < S; ++i) {
            TOUT *o = y + (((long long)n * 3 + c) * OH + (S * gy + i)) * OW + S * gx;
#pragma unroll
            for (int j = 0; j < S; ++j) o[j] = from_f32<TOUT>(acc[c * S * S + i * S + j] + out_add);
        }
}

template <int CP, int S, int TW, int TH>
inline size_t wdsr_tail_f32_smem() {
    constexpr int NO = round_up(3 * S * S, 4);
    return sizeof(float) * (size_t)(9 * CP * NO + 75 * NO + NO + CP * ((TW + 2) * (TH + 2) + 1) + 3 * ((TW + 4) * (TH + 4) + 1));
}

}  // namespace b200sr
