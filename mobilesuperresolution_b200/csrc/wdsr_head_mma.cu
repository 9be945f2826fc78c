// wdsr_head_mma.cu -- WDSR-B head, trunk = conv3x3(x - mean, Wh) + bh (3 -> 24 channels, models/basic_wdsr_b.py:86-87), for the
// planar-8 bf16 trunk [N][3][H][W][8] of the tcgen05 path, on mma.sync.
//
// Why not tcgen05 here (wdsr_tc5_head.cuh, 17 us at cfg2 against 5 us of HBM time): the head is 648 MAC per pixel -- three MMAs per
// 128-pixel M-tile -- so a tcgen05 CTA is one builder -> MMA -> epilogue chain whose hand-offs (~1 k clk per M-tile)
// cost more than the arithmetic.  The legacy tensor pipe needs no hand-off: every warp stages nothing but its CTA's (x - mean) tile
// (NHWC4 bf16, 8 bytes per pixel), builds the im2col A fragments with 4-byte shared-memory loads (k = tap * 4 + channel: a fragment
// register is a pixel's channel pair), runs 2 x m16n8k16 + 1 x m16n8k8 per 8 output channels and stores its C fragments straight to
// the trunk -- a warp-wide 4-byte store covers 8 pixels x 16 bytes = 128 contiguous bytes of one channel plane.  At ~600 MAC/clk/SM
// (tools/hmma_bench.cu) the 960 padded MAC per pixel are ~3 us at cfg2, under the memory time; many small CTAs hide the latencies.
// (ncu: the top stall is the math-pipe throttle of the HMMA pipe, 13.9 us cold.  A dense K = 32 form -- k = channel * 9 + tap, 6 instead of
// 9 HMMA per 16 pixels, fragments assembled from 2-byte loads -- was built and measured SLOWER, 12.0 against 11.3 us: the extra LDS / PRMT
// cost more than the three m16n8k8 saved.)
// Zero padding happens in the (x - mean) domain, exactly like the reference (pads are 0 after the mean subtraction).
#include "common.cuh"
#include "launch.h"

namespace b200sr {

namespace headmma {
constexpr int TW = 32, TH = 16, XW = TW + 2, XH = TH + 2, NTHREADS = 256, CP = 24;
}

template <typename TIN>
__global__ void __launch_bounds__(headmma::NTHREADS, 4)
wdsr_head_mma_kernel(const TIN *__restrict__ x, bf16 *__restrict__ trunk, const float *__restrict__ wh /* [27][CP] (k = c*9+tap) | bias[CP] */,
                     int N, int H, int W, int tiles_x, int tiles_y, float mean) {
    using namespace headmma;
    __shared__ __align__(16) uint2 x4[XH * XW];      // (c0, c1 | c2, 0) bf16 of x - mean, zero outside the image
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const int tile = blockIdx.x;
    const int x0 = (tile % tiles_x) * TW, y0 = ((tile / tiles_x) % tiles_y) * TH, n = tile / (tiles_x * tiles_y);
    const long long plane = (long long)H * W;

    // ---- the tile's input: every load is issued before the first use (one exposed memory latency per CTA)
    constexpr int NIT = (XH * XW + NTHREADS - 1) / NTHREADS;
    float v[NIT][3];
#pragma unroll
    for (int k = 0; k < NIT; ++k) {
        const int i = tid + NTHREADS * k;
        const int gy = y0 - 1 + i / XW, gx = x0 - 1 + i % XW;
        const bool ok = i < XH * XW && gy >= 0 && gy < H && gx >= 0 && gx < W;
        const long long o = ok ? ((long long)n * 3) * plane + (long long)gy * W + gx : 0;
#pragma unroll
        for (int c = 0; c < 3; ++c) v[k][c] = ok ? to_f32<TIN>(x[o + c * plane]) - mean : 0.f;
    }
    // ---- B fragments (bf16) and bias of this lane while the loads are in flight: k = tap * 4 + channel (channel 3 and taps 9.. are zero)
    auto wk = [&](int k, int o) -> float {
        const int tap = k >> 2, c = k & 3;
        return (tap < 9 && c < 3) ? __ldg(wh + (c * 9 + tap) * CP + o) : 0.f;
    };
    uint32_t bfr[3][5];
    float bias[3][2];
#pragma unroll
    for (int nt = 0; nt < 3; ++nt) {
        const int o = 8 * nt + g;                       // B[k][n = g]
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            bfr[nt][2 * s] = pack_bf16x2(wk(16 * s + 2 * t, o), wk(16 * s + 2 * t + 1, o));
            bfr[nt][2 * s + 1] = pack_bf16x2(wk(16 * s + 2 * t + 8, o), wk(16 * s + 2 * t + 9, o));
        }
        bfr[nt][4] = pack_bf16x2(wk(32 + 2 * t, o), wk(32 + 2 * t + 1, o));
        bias[nt][0] = __ldg(wh + 27 * CP + 8 * nt + 2 * t), bias[nt][1] = __ldg(wh + 27 * CP + 8 * nt + 2 * t + 1);
    }
    // byte offsets of this lane's A-fragment words relative to the output pixel's window origin: tap = 4 s + (t >> 1) (+ 2), pair = t & 1
    int aoff[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        int tap = 2 * j + (t >> 1);
        if (tap > 8) tap = 8;                           // zero weights: any finite value
        aoff[j] = ((tap / 3) * XW + tap % 3) * 8 + (t & 1) * 4;
    }
#pragma unroll
    for (int k = 0; k < NIT; ++k) {
        const int i = tid + NTHREADS * k;
        if (i < XH * XW) x4[i] = make_uint2(pack_bf16x2(v[k][0], v[k][1]), pack_bf16x2(v[k][2], 0.f));
    }
    __syncthreads();

    // ---- 32 M-tiles of 16 pixels (tile row ty, half-row xh, 16 px along x); four per warp
    const uint8_t *xb = reinterpret_cast<const uint8_t *>(x4);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int mt = warp * 4 + j, ty = mt >> 1, xh = mt & 1;
        const uint8_t *p0 = xb + (ty * XW + 16 * xh + g) * 8, *p1 = p0 + 64;      // window origins of pixel rows g and g + 8
        auto ld = [](const uint8_t *p) { return *reinterpret_cast<const uint32_t *>(p); };
        float acc[3][4];
#pragma unroll
        for (int nt = 0; nt < 3; ++nt) acc[nt][0] = bias[nt][0], acc[nt][1] = bias[nt][1], acc[nt][2] = bias[nt][0], acc[nt][3] = bias[nt][1];
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const uint32_t a0 = ld(p0 + aoff[2 * s]), a1 = ld(p1 + aoff[2 * s]), a2 = ld(p0 + aoff[2 * s + 1]), a3 = ld(p1 + aoff[2 * s + 1]);
#pragma unroll
            for (int nt = 0; nt < 3; ++nt) mma_16816(acc[nt], a0, a1, a2, a3, bfr[nt][2 * s], bfr[nt][2 * s + 1]);
        }
        {
            const uint32_t a0 = ld(p0 + aoff[4]), a1 = ld(p1 + aoff[4]);
#pragma unroll
            for (int nt = 0; nt < 3; ++nt) mma_1688(acc[nt], a0, a1, bfr[nt][4]);
        }
        const int gy = y0 + ty, gx = x0 + 16 * xh + g;
        if (gy < H) {
            bf16 *o = trunk + (((long long)n * 3 * H + gy) * W + gx) * 8 + 2 * t;     // planar-8 trunk: plane q is H*W*8 elements further
#pragma unroll
            for (int nt = 0; nt < 3; ++nt) {
                if (gx < W) *reinterpret_cast<uint32_t *>(o + nt * plane * 8) = pack_bf16x2(acc[nt][0], acc[nt][1]);
                if (gx + 8 < W) *reinterpret_cast<uint32_t *>(o + nt * plane * 8 + 64) = pack_bf16x2(acc[nt][2], acc[nt][3]);
            }
        }
    }
}

template <typename TIN>
static cudaError_t head_mma_t(const void *x, void *trunk, const float *wh, int N, int H, int W, float mean, cudaStream_t st) {
    using namespace headmma;
    const int tx = ceil_div(W, TW), ty = ceil_div(H, TH);
    wdsr_head_mma_kernel<TIN><<<tx * ty * N, NTHREADS, 0, st>>>((const TIN *)x, (bf16 *)trunk, wh, N, H, W, tx, ty, mean);
    return cudaGetLastError();
}

cudaError_t launch_head_mma(int xd, const void *x, void *trunk, const float *wh, int N, int H, int W, float mean, cudaStream_t st) {
    if (xd == kF32) return head_mma_t<float>(x, trunk, wh, N, H, W, mean, st);
    if (xd == kBF16) return head_mma_t<bf16>(x, trunk, wh, N, H, W, mean, st);
    return cudaErrorInvalidValue;
}

}  // namespace b200sr
