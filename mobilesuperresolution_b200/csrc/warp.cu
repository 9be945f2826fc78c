// warp.cu -- flow_warp: bilinear gather at (x + fx, y + fy).          models/spynet_arch.py:98-129
//
// The reference builds a pixel mesh, adds the flow, normalises to [-1,1] (`2*v/max(size-1,1) - 1`) and calls
// F.grid_sample(bilinear, padding_mode, align_corners=True), which un-normalises with ((g+1)/2)*(size-1).  The
// kernels replay that round trip in fp32 so the sampling position matches the reference to the last ulp, then
//   zeros : each of the 4 corners contributes only if it lies inside the image
//   border: the coordinate is clamped to [0, size-1] before it is split into corners.
// Coordinates and blend weights are always fp32, also when the features are bf16 (bf16 resolves only 2 px at x~300).
#include <cstdlib>

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

// NCHW fp32 (the reference's tensor layout).  One CTA owns a 32x8 pixel tile, one thread per pixel: lanes are consecutive x
// (coalesced 128-byte row segments in every channel plane), and the 8 rows of the tile share their corner rows through L1
// (row y0+1 of one pixel row is row y0 of the next), so a smooth flow costs ~1.2 reads of the input from L2/HBM instead of 2.
// The channel loop is unrolled by 8 = 32 independent loads in flight per thread.
constexpr int WTX = 32, WTY = 8;
__global__ void __launch_bounds__(WTX * WTY) flow_warp_nchw_kernel(const float *__restrict__ x, const float *__restrict__ flow,
                                                                  long long fs_n, long long fs_h, long long fs_w, long long fs_c,
                                                                  float *__restrict__ y, int N, int C, int H, int W, int border,
                                                                  int tiles_x, int tiles_y) {
    const long long ntiles = (long long)N * tiles_x * tiles_y;
    const int lx = threadIdx.x & (WTX - 1), ly = threadIdx.x / WTX;
    const long long HW = (long long)H * W;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int tx = (int)(t % tiles_x), ty = (int)((t / tiles_x) % tiles_y), n = (int)(t / ((long long)tiles_x * tiles_y));
        const int xw = tx * WTX + lx, yh = ty * WTY + ly;
        if (xw >= W || yh >= H) continue;
        const float *f = flow + n * fs_n + yh * fs_h + xw * fs_w;
        const Bilin b = bilinear_setup(f[0], f[fs_c], xw, yh, W, H, border != 0);
        // out-of-image corners: weight 0 and a clamped (always legal) address, so the loads need no predicates
        const int x0 = min(max(b.x0, 0), W - 1), x1 = min(max(b.x0 + 1, 0), W - 1);
        const int y0 = min(max(b.y0, 0), H - 1), y1 = min(max(b.y0 + 1, 0), H - 1);
        const float w00 = b.v00 ? b.w00 : 0.f, w01 = b.v01 ? b.w01 : 0.f, w10 = b.v10 ? b.w10 : 0.f, w11 = b.v11 ? b.w11 : 0.f;
        const long long o00 = (long long)y0 * W + x0, o01 = (long long)y0 * W + x1, o10 = (long long)y1 * W + x0, o11 = (long long)y1 * W + x1;
        const float *xp = x + (long long)n * C * HW;
        float *yp = y + (long long)n * C * HW + (long long)yh * W + xw;
        int c = 0;
        for (; c + 8 <= C; c += 8) {
            float v[8][4];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const float *pc = xp + (long long)(c + u) * HW;
                v[u][0] = __ldg(pc + o00), v[u][1] = __ldg(pc + o01), v[u][2] = __ldg(pc + o10), v[u][3] = __ldg(pc + o11);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                // same association as the one-corner-at-a-time form: ((v00*w00 + v01*w01) + v10*w10) + v11*w11
                float acc = 0.f;
                if (b.v00) acc += v[u][0] * w00;
                if (b.v01) acc += v[u][1] * w01;
                if (b.v10) acc += v[u][2] * w10;
                if (b.v11) acc += v[u][3] * w11;
                __stcs(yp + (long long)(c + u) * HW, acc);
            }
        }
        for (; c < C; ++c) {
            const float *pc = xp + (long long)c * HW;
            float acc = 0.f;
            if (b.v00) acc += __ldg(pc + o00) * w00;
            if (b.v01) acc += __ldg(pc + o01) * w01;
            if (b.v10) acc += __ldg(pc + o10) * w10;
            if (b.v11) acc += __ldg(pc + o11) * w11;
            __stcs(yp + (long long)c * HW, acc);
        }
    }
}

cudaError_t launch_flow_warp_nchw(const float *x, const float *flow, long long fs_n, long long fs_h, long long fs_w,
                                  long long fs_c, float *y, int n, int c, int h, int w, int border, cudaStream_t st) {
    const long long P = (long long)n * h * w;
    if (P == 0 || c == 0) return cudaSuccess;
    const int tx = ceil_div(w, WTX), ty = ceil_div(h, WTY);
    long long blocks = (long long)n * tx * ty;
    static const int per_sm = [] { const char *e = getenv("B200SR_WARPN_CTAS_PER_SM"); return e ? atoi(e) : 8; }();   // (developer sweep, 4 x 64 x 720 x 1280: 4 / 8 / 16 / 32 per SM -> 1.9 / 3.1 / 3.0 / 2.8 TB/s)
    const long long cap = (long long)sm_count() * per_sm;
    if (blocks > cap) blocks = cap;
    flow_warp_nchw_kernel<<<(unsigned)blocks, WTX * WTY, 0, st>>>(x, flow, fs_n, fs_h, fs_w, fs_c, y, n, c, h, w, border, tx, ty);
    return cudaGetLastError();
}

// NHWC (video path internal layout): a pixel's C channels are Q = C*esize/16 consecutive 16-byte vectors.
// One CTA owns a 32x8 pixel tile (the four corner rows of neighbouring pixels are re-used through L1, not L2); one warp owns
// one 32-pixel row of it.  Sampling position and blend weights are computed ONCE per pixel with all 32 lanes busy (lane = x),
// then the warp walks its 32 pixels, 32/Q (Q a power of two) or 1 pixel(s) per step: the pixel's setup is broadcast with
// __shfl_sync and every lane gathers its own 16-byte channel slice of the four corners and blends in fp32.
// (The first form of this kernel recomputed the setup in every step with 1/Q of the lanes and was instruction-issue bound.)
template <typename T, int Q>
__global__ void __launch_bounds__(256) flow_warp_nhwc_kernel(const T *__restrict__ x, const float *__restrict__ flow,
                                                             T *__restrict__ y, int N, int C, int H, int W, int border,
                                                             int tiles_x, int tiles_y, int y_cs, int y_co, int x_cs, int x_co) {
    constexpr int VEC = 16 / sizeof(T);
    constexpr bool kPow2 = (Q & (Q - 1)) == 0 && Q <= 32;
    constexpr int PPS = kPow2 ? 32 / Q : 1;            // pixels per step
    constexpr int LPP = kPow2 ? Q : 32;                // lanes that share one pixel in a step
    constexpr int SUB = kPow2 ? 1 : (Q + 31) / 32;     // 16-byte slices per lane when a pixel is wider than a warp step
    const int lane = threadIdx.x & 31, wrow = threadIdx.x >> 5;
    const long long ntiles = (long long)N * tiles_x * tiles_y;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int tx = (int)(t % tiles_x), ty = (int)((t / tiles_x) % tiles_y), n = (int)(t / ((long long)tiles_x * tiles_y));
        const int yh = ty * WTY + wrow;
        if (yh >= H) continue;  // whole warp
        const int xw = tx * WTX + lane;
        Bilin b = {};
        if (xw < W) {
            const float *f = flow + ((long long)n * 2 * H + yh) * W + xw;
            b = bilinear_setup(__ldg(f), __ldg(f + (long long)H * W), xw, yh, W, H, border != 0);
        }
        const unsigned flags = (b.v00 ? 1u : 0u) | (b.v01 ? 2u : 0u) | (b.v10 ? 4u : 0u) | (b.v11 ? 8u : 0u);
        // clamped corner offsets (elements, relative to the image): out-of-image corners get weight 0 and a legal address
        const int cx0 = min(max(b.x0, 0), W - 1), cx1 = min(max(b.x0 + 1, 0), W - 1);
        const int cy0 = min(max(b.y0, 0), H - 1), cy1 = min(max(b.y0 + 1, 0), H - 1);
        const int o00 = cy0 * W + cx0, o01 = cy0 * W + cx1, o10 = cy1 * W + cx0, o11 = cy1 * W + cx1;   // < 2^31: per-image pixel index
        const T *xi = x + (long long)n * H * W * x_cs + x_co;   // x may be a channel window of a wider tensor too
        T *yrow = y + (((long long)n * H + yh) * W + (long long)tx * WTX) * y_cs + y_co;   // y may be a channel window of a wider tensor
        const int npx = min(WTX, W - tx * WTX);
#pragma unroll 2
        for (int s0 = 0; s0 < 32; s0 += PPS) {
            const int src = s0 + lane / LPP;            // which of the warp's 32 pixels this lane works on in this step
            const int p00 = __shfl_sync(0xffffffffu, o00, src), p01 = __shfl_sync(0xffffffffu, o01, src);
            const int p10 = __shfl_sync(0xffffffffu, o10, src), p11 = __shfl_sync(0xffffffffu, o11, src);
            const float w00 = __shfl_sync(0xffffffffu, b.w00, src), w01 = __shfl_sync(0xffffffffu, b.w01, src);
            const float w10 = __shfl_sync(0xffffffffu, b.w10, src), w11 = __shfl_sync(0xffffffffu, b.w11, src);
            const unsigned fl = __shfl_sync(0xffffffffu, flags, src);
            if (src >= npx) continue;
#pragma unroll
            for (int sub = 0; sub < SUB; ++sub) {
                const int q = kPow2 ? (lane % LPP) : lane + 32 * sub;
                if (!kPow2 && q >= Q) break;
                const T *base = xi + q * VEC;
                const uint4 v00 = __ldg(reinterpret_cast<const uint4 *>(base + (long long)p00 * x_cs));
                const uint4 v01 = __ldg(reinterpret_cast<const uint4 *>(base + (long long)p01 * x_cs));
                const uint4 v10 = __ldg(reinterpret_cast<const uint4 *>(base + (long long)p10 * x_cs));
                const uint4 v11 = __ldg(reinterpret_cast<const uint4 *>(base + (long long)p11 * x_cs));
                float acc[VEC];
#pragma unroll
                for (int k = 0; k < VEC; ++k) acc[k] = 0.f;
                auto corner = [&](bool valid, const uint4 &v, float wgt) {   // same order and association as the reference restatement
                    if (!valid) return;
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
                corner(fl & 1u, v00, w00);
                corner(fl & 2u, v01, w01);
                corner(fl & 4u, v10, w10);
                corner(fl & 8u, v11, w11);
                uint4 o;
                if constexpr (sizeof(T) == 4) {
                    o = *reinterpret_cast<uint4 *>(acc);
                } else {
                    o.x = pack_bf16x2(acc[0], acc[1]);
                    o.y = pack_bf16x2(acc[2], acc[3]);
                    o.z = pack_bf16x2(acc[4], acc[5]);
                    o.w = pack_bf16x2(acc[6], acc[7]);
                }
                __stcs(reinterpret_cast<uint4 *>(yrow + (long long)src * y_cs + q * VEC), o);
            }
        }
    }
}

// bf16, Q = C / 8 an even power of two (C = 16, 32, 64, 128): the instruction-lean form.  The general kernel above is issue bound (64 % of
// the issue slots for 3.3-3.8 TB/s at C = 64): ~97 instructions per 16 output bytes, 68 of them the fp32 blend of unpacked bf16 pairs.  Here
//   * a lane owns TWO adjacent 16-byte slices of a pixel (32 bytes): the per-pixel set-up is fetched once per 32 output bytes,
//   * the set-up (four clamped corner offsets, four weights already zeroed for out-of-image corners) goes through shared memory -- two
//     broadcast LDS.128 instead of nine shuffles,
//   * the high bf16 of a pair is used as an fp32 WITHOUT masking its low half (the stray bits are < 2^-7 of a bf16 ulp of the operand;
//     the output is rounded to bf16 anyway); the low one costs one shift.
// Zero weights replace the validity predicates: an out-of-image corner adds v * 0 (finite inputs, like everywhere in this path).
template <int Q>
__global__ void __launch_bounds__(256) flow_warp_nhwc_bf16_lean_kernel(const bf16 *__restrict__ x, const float *__restrict__ flow,
                                                                       bf16 *__restrict__ y, int N, int C, int H, int W, int border, int tiles_x,
                                                                       int tiles_y, int y_cs, int y_co, int x_cs, int x_co) {
    constexpr int LPP = Q / 2;          // lanes per pixel
    constexpr int PPS = 32 / LPP;       // pixels per step
    __shared__ int4 s_off[WTY][32];
    __shared__ float4 s_w[WTY][32];
    const int lane = threadIdx.x & 31, wrow = threadIdx.x >> 5;
    const long long ntiles = (long long)N * tiles_x * tiles_y;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int tx = (int)(t % tiles_x), ty = (int)((t / tiles_x) % tiles_y), n = (int)(t / ((long long)tiles_x * tiles_y));
        const int yh = ty * WTY + wrow;
        if (yh >= H) continue;  // whole warp
        const int xw = tx * WTX + lane;
        Bilin b = {};
        if (xw < W) {
            const float *f = flow + ((long long)n * 2 * H + yh) * W + xw;
            b = bilinear_setup(__ldg(f), __ldg(f + (long long)H * W), xw, yh, W, H, border != 0);
        }
        const int cx0 = min(max(b.x0, 0), W - 1), cx1 = min(max(b.x0 + 1, 0), W - 1);
        const int cy0 = min(max(b.y0, 0), H - 1), cy1 = min(max(b.y0 + 1, 0), H - 1);
        __syncwarp();   // the previous tile's reads of this warp's rows are done
        s_off[wrow][lane] = make_int4((cy0 * W + cx0) * x_cs, (cy0 * W + cx1) * x_cs, (cy1 * W + cx0) * x_cs, (cy1 * W + cx1) * x_cs);
        s_w[wrow][lane] = make_float4(b.v00 ? b.w00 : 0.f, b.v01 ? b.w01 : 0.f, b.v10 ? b.w10 : 0.f, b.v11 ? b.w11 : 0.f);
        __syncwarp();
        const bf16 *xi = x + (long long)n * H * W * x_cs + x_co + (lane % LPP) * 16;
        bf16 *yrow = y + (((long long)n * H + yh) * W + (long long)tx * WTX) * y_cs + y_co + (lane % LPP) * 16;
        const int npx = min(WTX, W - tx * WTX);
#pragma unroll 2
        for (int s0 = 0; s0 < 32; s0 += PPS) {
            const int src = s0 + lane / LPP;
            if (src >= npx) continue;
            const int4 o = s_off[wrow][src];
            const float4 wv = s_w[wrow][src];
            uint4 v[4][2];
            const int off[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint4 *pk = reinterpret_cast<const uint4 *>(xi + off[k]);
                v[k][0] = __ldg(pk), v[k][1] = __ldg(pk + 1);
            }
            const float wk[4] = {wv.x, wv.y, wv.z, wv.w};
            float acc[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k)   // same order and association as the general kernel: ((v00 w00 + v01 w01) + v10 w10) + v11 w11
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint32_t *u = reinterpret_cast<const uint32_t *>(&v[k][h]);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        acc[8 * h + 2 * i] = fmaf(__uint_as_float(u[i] << 16), wk[k], acc[8 * h + 2 * i]);
                        acc[8 * h + 2 * i + 1] = fmaf(__uint_as_float(u[i]), wk[k], acc[8 * h + 2 * i + 1]);
                    }
                }
            uint4 o0, o1;
            o0.x = pack_bf16x2(acc[0], acc[1]), o0.y = pack_bf16x2(acc[2], acc[3]), o0.z = pack_bf16x2(acc[4], acc[5]), o0.w = pack_bf16x2(acc[6], acc[7]);
            o1.x = pack_bf16x2(acc[8], acc[9]), o1.y = pack_bf16x2(acc[10], acc[11]), o1.z = pack_bf16x2(acc[12], acc[13]), o1.w = pack_bf16x2(acc[14], acc[15]);
            uint4 *dst = reinterpret_cast<uint4 *>(yrow + (long long)src * y_cs);
            __stcs(dst, o0);
            __stcs(dst + 1, o1);
        }
    }
}

template <typename T, int Q>
static cudaError_t warp_nhwc_t(const void *x, const float *flow, void *y, int n, int c, int h, int w, int border, int y_cs, int y_co,
                               int x_cs, int x_co, cudaStream_t st) {
    const int tx = ceil_div(w, WTX), ty = ceil_div(h, WTY);
    long long blocks = (long long)n * tx * ty;
    static const int per_sm = [] { const char *e = getenv("B200SR_WARPG_CTAS_PER_SM"); return e ? atoi(e) : 4; }();   // (developer sweep, fp32 C = 64 at 720p: 2 / 4 / 8 / 32 per SM -> 2.6 / 4.4 / 4.0 / 4.2 TB/s)
    const long long cap = (long long)sm_count() * per_sm;
    if (blocks > cap) blocks = cap;
    flow_warp_nhwc_kernel<T, Q><<<(unsigned)blocks, 256, 0, st>>>((const T *)x, flow, (T *)y, n, c, h, w, border, tx, ty, y_cs, y_co, x_cs, x_co);
    return cudaGetLastError();
}

template <typename T>
static cudaError_t warp_nhwc_q(int Q, const void *x, const float *flow, void *y, int n, int c, int h, int w, int border, int y_cs, int y_co,
                               int x_cs, int x_co, cudaStream_t st) {
    switch (Q) {
        case 1: return warp_nhwc_t<T, 1>(x, flow, y, n, c, h, w, border, y_cs, y_co, x_cs, x_co, st);
        case 2: return warp_nhwc_t<T, 2>(x, flow, y, n, c, h, w, border, y_cs, y_co, x_cs, x_co, st);
        case 3: return warp_nhwc_t<T, 3>(x, flow, y, n, c, h, w, border, y_cs, y_co, x_cs, x_co, st);
        case 4: return warp_nhwc_t<T, 4>(x, flow, y, n, c, h, w, border, y_cs, y_co, x_cs, x_co, st);
        case 6: return warp_nhwc_t<T, 6>(x, flow, y, n, c, h, w, border, y_cs, y_co, x_cs, x_co, st);
        case 8: return warp_nhwc_t<T, 8>(x, flow, y, n, c, h, w, border, y_cs, y_co, x_cs, x_co, st);
        case 16: return warp_nhwc_t<T, 16>(x, flow, y, n, c, h, w, border, y_cs, y_co, x_cs, x_co, st);
    }
    return cudaErrorInvalidValue;
}

cudaError_t launch_flow_warp_nhwc(const void *x, const float *flow_nchw, void *y, int n, int c, int h, int w, int border,
                                  int dtype, cudaStream_t st, int y_cs, int y_co, int x_cs, int x_co) {
    if ((long long)n * h * w == 0) return cudaSuccess;
    if (y_cs <= 0) y_cs = c, y_co = 0;
    if (x_cs <= 0) x_cs = c, x_co = 0;
    const int vec = dtype == kF32 ? 4 : 8;     // the windows must keep the 16-byte loads / stores aligned
    if (c % vec || y_cs % vec || y_co % vec || y_co + c > y_cs || x_cs % vec || x_co % vec || x_co + c > x_cs) return cudaErrorInvalidValue;
    if (dtype == kF32) return warp_nhwc_q<float>(c / 4, x, flow_nchw, y, n, c, h, w, border, y_cs, y_co, x_cs, x_co, st);
    const int Q = c / 8;
    // lean bf16 form: two slices per lane need 32-byte aligned windows and int32 element offsets into one image
    if ((Q == 2 || Q == 4 || Q == 8 || Q == 16) && x_cs % 16 == 0 && x_co % 16 == 0 && y_cs % 16 == 0 && y_co % 16 == 0 &&
        (long long)h * w * x_cs < (1ll << 31) && !getenv("B200SR_WARP_GENERAL")) {
        const int tx = ceil_div(w, WTX), ty = ceil_div(h, WTY);
        long long blocks = (long long)n * tx * ty;
        static const int per_sm = [] { const char *e = getenv("B200SR_WARP_CTAS_PER_SM"); return e ? atoi(e) : 4; }();   // (developer sweep: 2 / 3 / 4 / 5 / 6 / 8 / 16 / 32 per SM -> 3.3 / 4.3 / 4.9 / 3.6 / 3.9 / 4.7 / 4.7 / 4.6 TB/s at 8x720x1280x64)
        const long long cap = (long long)sm_count() * per_sm;   // resident CTAs walk the tile list
        if (blocks > cap) blocks = cap;
#define B200SR_WARP_LEAN(QQ)                                                                                                      \
    if (Q == QQ) {                                                                                                                \
        flow_warp_nhwc_bf16_lean_kernel<QQ><<<(unsigned)blocks, 256, 0, st>>>((const bf16 *)x, flow_nchw, (bf16 *)y, n, c, h, w, border, tx, ty, \
                                                                               y_cs, y_co, x_cs, x_co);                           \
        return cudaGetLastError();                                                                                                \
    }
        B200SR_WARP_LEAN(2) B200SR_WARP_LEAN(4) B200SR_WARP_LEAN(8) B200SR_WARP_LEAN(16)
#undef B200SR_WARP_LEAN
    }
    return warp_nhwc_q<bf16>(Q, x, flow_nchw, y, n, c, h, w, border, y_cs, y_co, x_cs, x_co, st);
}

}  // namespace b200sr
