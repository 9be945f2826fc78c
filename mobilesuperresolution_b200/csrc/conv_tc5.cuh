// conv_tc5.cuh -- 3x3 "same" convolution from 64 (or 65..80) bf16 channels, on tcgen05: the BasicVSR propagation trunks
// (ConvResidualBlocks: 30 x [conv-ReLU-conv + x] per frame and direction, models/basicvsr_arch_origin.py:98-137), the upsampler
// convs upconv1 / upconv2 (64 -> 256 + PixelShuffle(2), :87-88), conv_hr (:89) and conv_last + bilinear base (:90-92).
//
//   y[n, oy, ox, co] = act( bias[co] + sum_{ky,kx,ci} x[n, oy+ky-1, ox+kx-1, ci] * w[co][ci][ky][kx] ) (+ residual)
//
// Same producer -> MMA -> epilogue chain as the WDSR tail (wdsr_tc5_tail.cuh), no builders: an implicit GEMM with pixels as M.
//   warp 0      TMA      the filter image by ONE cp.async.bulk (a constant: in flight while the previous kernel drains), then NCH
//                        cp.async.bulk.tensor.4d per tile (one per 8-channel chunk) of a 32 x 10 pixel box (30 x 8 outputs +
//                        1-pixel halo) into chunk-planar shared memory [chunk][pixel][16 B]; out-of-image pixels (and channels past
//                        cin) are zero-filled by the TMA unit (= the conv's zero padding); three (two for NCH = 10) tile buffers.
//                        x is NHWC (every 16-byte pixel row of a chunk is its own TMA request: 320 per chunk) or planar-8
//                        [n][c/8][h][w][8] (a box row is 512 contiguous bytes: 10 requests per chunk) -- the trunk's private tensors
//   warp 1      MMA      per 128-pixel M-tile (four 32-pixel box rows) 9 * NCH / 2 tcgen05.mma (M = 128, N = NOUT, K = 16): a tap is a
//                        constant pixel offset of the A operand's start address (the two box columns right of the 30 outputs
//                        compute don't-care rows), the two 8-channel chunks of a K step are paired through the LBO stride
//   warps 2-5 / 6-9      epilogue of the tile's first / second M-tile: tcgen05.ld -> + bias -> activation -> (+ residual) -> bf16
//                        stores: NHWC pixel rows (channel windows of wider tensors allowed), planar-8 planes, or PixelShuffle(2)
//                        folded into either.  cout = 64 G: a CTA serves one group of 64 output channels (own filter image).
//                        NOUT = 16 ("rgb" form, 64 -> 3): the horizontal taps ride in N (9 accumulator columns, 12 MMAs per M-tile), the
//                        epilogue adds them across lanes, + bias + x4 bilinear base -> fp32 NCHW frame.
// The generic mma.sync kernel (conv.cuh) ran these convolutions at ~110-139 TFLOP/s; they are 90 % of a BasicVSR clip's FLOPs.
// Measured (profiles/r01_conv_tc5_ncu.md): 925 TFLOP/s at 720 x 1280 = 66 % of the sustained bf16 peak -- an SS MMA with N = 64
// reads 6 KB of operands from shared memory for 32 clk of tensor work, so ~2/3 of peak is this operand shape's ceiling.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "conv.cuh"
#include "tc5.cuh"

namespace b200sr {
namespace tc5conv {
constexpr int BW = 32, NTHREADS = 320;
constexpr int CTRL = 256;
enum Bar { TC_FULL = 0 /*3*/, TC_EMPTY = 3 /*3*/, D_FULL = 6, D_EMPTY = 8, W_READY = 10, NBARS = 11 };
// NCH = 8-channel chunks of the input: 8 (64 channels) or 10 (65..80 channels: the trunk's first conv on [x_i | warped feat])
// NOUT = output channels per CTA: 64, or 16 for the 64 -> 3 "rgb" form (conv_last + bilinear base, fp32 NCHW store)
// MT = 128-pixel M-tiles per tile: 2 (30 x 8 outputs), or 1 (30 x 4) for launches of only a few tiles per CTA -- one 180 x 320 frame
//      on a 74-CTA grid is 3.4 rounds of 30 x 8 tiles (4 are paid for) but 6.7 rounds of 30 x 4 tiles (7 half-size ones)
// KS = 3 (3x3, 30 outputs per 32-pixel box row) or 1 (1x1: no halo, no don't-care columns; NCH = 16 for the 128 -> 64 k fusion conv)
template <int NCH, int NOUT = 64, int MT = 2, int KS = 3> struct Cfg {
    static constexpr int HALO = KS / 2, TAPS = KS * KS, TWO = BW - 2 * HALO;
    static constexpr int TH = 4 * MT, BH = TH + 2 * HALO;
    static constexpr int PLANE_PX = BW * BH + 8;            // + 8 zero pixels: the last taps of the last M-tile read past the box
    static constexpr int PLANE = PLANE_PX * 16;             // 5,248 B (3,200 for MT = 1)
    static constexpr int TILE_BUF = NCH * PLANE;            // 41,984 / 52,480 B
    static constexpr int NBUF = NCH <= 8 ? 3 : 2;
    static constexpr int W_SBO = TAPS * NCH * 128;          // weight image [NOUT/8 row groups][TAPS * NCH (tap, chunk) slices][8 rows][16 B]
    static constexpr int W_BYTES = (NOUT / 8) * W_SBO;      // 73,728 / 92,160 B (18,432 for NOUT = 16)
    static constexpr size_t smem_bytes() { return (size_t)CTRL + NBUF * TILE_BUF + W_BYTES + 256; }
};
}  // namespace tc5conv

__device__ __forceinline__ void tma_load_4d_conv(uint32_t dst_saddr, const void *tmap, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst_saddr), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

__device__ __forceinline__ void bulk_load_g2s(uint32_t dst_saddr, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_saddr), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}

// x4 bilinear upsample (align_corners = False) of one plane of the low-resolution image at HR pixel (oy, ox): the arithmetic of
// vsr_base_add_kernel (video_glue.cu) / F.interpolate(x_i, scale_factor=4, mode='bilinear')   models/basicvsr_arch_origin.py:91
__device__ __forceinline__ float bilinear_x4(const float *__restrict__ p, int h, int w, int oy, int ox) {
    const float sy = fmaxf(0.25f * ((float)oy + 0.5f) - 0.5f, 0.f), sx = fmaxf(0.25f * ((float)ox + 0.5f) - 0.5f, 0.f);
    const int y0 = min((int)sy, h - 1), x0 = min((int)sx, w - 1);
    const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
    const float ly = sy - (float)y0, lx = sx - (float)x0;
    return (1.f - ly) * ((1.f - lx) * p[y0 * w + x0] + lx * p[y0 * w + x1]) + ly * ((1.f - lx) * p[y1 * w + x0] + lx * p[y1 * w + x1]);
}

template <int NCH, int NOUT, int MT, int KS>
__global__ void __launch_bounds__(tc5conv::NTHREADS, 1)
conv3x3_c64_tc5_kernel(const __grid_constant__ CUtensorMap tmap_x, ConvArgs a, const uint8_t *__restrict__ wimg, int tiles_x, int tiles_y,
                       int ntiles) {
    using namespace tc5conv;
    using C = Cfg<NCH, NOUT, MT, KS>;
    constexpr int TILE_BUF = C::TILE_BUF, NBUF = C::NBUF, W_SBO = C::W_SBO, W_BYTES = C::W_BYTES, TH = C::TH, BH = C::BH, PLANE = C::PLANE;
    constexpr int HALO = C::HALO, TAPS = C::TAPS, TWO = C::TWO;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *ctrl = smem_raw;
    uint8_t *tc = smem_raw + CTRL;           // NBUF x TILE_BUF
    uint8_t *wsm = tc + NBUF * TILE_BUF;     // W_BYTES
    float *bias_s = reinterpret_cast<float *>(wsm + W_BYTES);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bars = smem_u32(ctrl);
    auto bar = [&](int b) { return bars + 8u * (uint32_t)b; };
    const uint32_t tc_u = smem_u32(tc), w_u = smem_u32(wsm);
    const int H = a.h, W = a.w_;

    // Launched with programmatic stream serialization: the set-up below (barriers, TMEM, the constant weight image) may overlap the
    // previous kernel's tail; every read of an activation tensor sits behind griddepcontrol.wait, every store behind those reads.
    tc5::pdl_launch_dependents();
    if (tid == 0) {
        for (int b = 0; b < NBUF; ++b) {
            tc5::mbar_init(bar(TC_FULL + b), 1);
            tc5::mbar_init(bar(TC_EMPTY + b), 1);
        }
        for (int e = 0; e < 2; ++e) {
            tc5::mbar_init(bar(D_FULL + e), 1);
            tc5::mbar_init(bar(D_EMPTY + e), 128);
        }
        tc5::mbar_init(bar(W_READY), 1);
        tc5::mbar_init_fence();
        tc5::tma_prefetch_desc(&tmap_x);
    }
    __syncwarp();
    if (warp == 0) tc5::tmem_alloc(smem_u32(ctrl + 240), 128);
    // cout = 64 G: output-channel group `grp` (its own weight image and bias slice) is fixed per CTA, the spatial tiles of a group
    // are dealt round-robin to the group's CTAs (gridDim.x is a multiple of G)
    const int G = NOUT == 64 ? a.cout >> 6 : 1, grp = (int)blockIdx.x % G, rank = (int)blockIdx.x / G, nranks = (int)gridDim.x / G;
    wimg += (size_t)grp * W_BYTES;
    if (tid < NOUT) bias_s[tid] = a.bias[grp * 64 + tid];
    for (int i = tid; i < NBUF * NCH * 8; i += NTHREADS)   // the 8 pad pixels of every plane stay zero (TMA never writes them)
        *reinterpret_cast<uint4 *>(tc + (i / 8) * PLANE + (BW * BH + i % 8) * 16) = make_uint4(0u, 0u, 0u, 0u);
    tc5::fence_proxy_async();
    tc5::fence_before_sync();
    __syncthreads();
    tc5::fence_after_sync();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t *>(ctrl + 240);
    const int nmine = rank < ntiles ? (ntiles - 1 - rank) / nranks + 1 : 0;
    auto tile_origin = [&](int it, int &x0, int &y0, int &n) {
        const int tile = rank + it * nranks;
        x0 = (tile % tiles_x) * TWO;
        y0 = ((tile / tiles_x) % tiles_y) * TH;
        n = tile / (tiles_x * tiles_y);
    };

    if (warp == 0) {
        // ============================== TMA producer ==============================
        if (tc5::elect_one()) {
            // the filter image is a constant: one bulk copy, in flight while the previous kernel drains and the first tile loads
            tc5::mbar_arrive_expect_tx(bar(W_READY), W_BYTES);
            bulk_load_g2s(w_u, wimg, W_BYTES, bar(W_READY));
            tc5::pdl_wait();
            for (int it = 0; it < nmine; ++it) {
                int x0, y0, n;
                tile_origin(it, x0, y0, n);
                const int b = it % NBUF;
                tc5::mbar_wait(bar(TC_EMPTY + b), ((it / NBUF) & 1) ^ 1);
                tc5::mbar_arrive_expect_tx(bar(TC_FULL + b), NCH * BW * BH * 16);
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    const uint32_t dst = tc_u + b * TILE_BUF + c * PLANE;
                    if (a.x_planar) tma_load_4d_conv(dst, &tmap_x, bar(TC_FULL + b), 4 * (x0 - HALO), y0 - HALO, c, n);   // (uint32 of a row, row, plane, image)
                    else tma_load_4d_conv(dst, &tmap_x, bar(TC_FULL + b), 8 * c, x0 - HALO, y0 - HALO, n);                 // (channel, x, y, image)
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ============================== MMA issuer ==============================
        const bool leader = tc5::elect_one();
        const uint32_t idesc = tc5::idesc_bf16_f32(128, NOUT);
        const uint64_t bw = tc5::smem_desc(w_u, 128, W_SBO), a0d = tc5::smem_desc(tc_u, PLANE, 128);   // A: chunk pairs through LBO = plane stride
        tc5::mbar_wait(bar(W_READY), 0);
        // M-tiles are numbered j = MT * it + m across the CTA's tiles; accumulator (and epilogue warpgroup) e = j & 1, k-th use = j >> 1
        for (int j = 0; j < MT * nmine; ++j) {
            const int it = j / MT, m = j % MT, b = it % NBUF, e = j & 1;
            if (m == 0) tc5::mbar_wait(bar(TC_FULL + b), (it / NBUF) & 1);
            tc5::mbar_wait(bar(D_EMPTY + e), ((j >> 1) & 1) ^ 1);
            tc5::fence_after_sync();
            if (leader) {
                const uint32_t d = tmem + e * NOUT;
                const uint64_t abase = a0d + (uint64_t)((b * TILE_BUF + m * 128 * 16) >> 4);
                if constexpr (NOUT == 16) {
                    // "rgb" form (64 -> 3): an N = 16 MMA costs what an N = 64 one does (the 4 KB A fetch bounds it), so the three HORIZONTAL
                    // taps ride in N instead of K: accumulator column dx * 3 + c of box column j = sum over (dy, ci) of x[by + dy][j][ci] *
                    // w[c][ci][dy][dx] -- 3 * NCH / 2 = 12 MMAs instead of 36 -- and the epilogue adds columns (dx, c) of lanes j + dx
#pragma unroll
                    for (int i = 0; i < KS * (NCH / 2); ++i) {   // (tap row dy, chunks 2 cp, 2 cp + 1)
                        const int dy = i / (NCH / 2), cp = i % (NCH / 2);
                        const int aoff = 2 * cp * PLANE + dy * BW * 16;
                        tc5::mma_ss(d, abase + (uint64_t)(aoff >> 4), bw + (uint64_t)(8 * (dy * NCH + 2 * cp)), idesc, i > 0);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < TAPS * (NCH / 2); ++i) {   // (tap t, chunks 2 cp, 2 cp + 1)
                        const int t = i / (NCH / 2), cp = i % (NCH / 2), dy = t / KS, dx = t % KS;
                        const int aoff = 2 * cp * PLANE + (dy * BW + dx) * 16;
                        tc5::mma_ss(d, abase + (uint64_t)(aoff >> 4), bw + (uint64_t)(8 * (t * NCH + 2 * cp)), idesc, i > 0);
                    }
                }
                tc5::commit(bar(D_FULL + e));
                if (m == MT - 1) tc5::commit(bar(TC_EMPTY + b));
            }
            __syncwarp();
        }
        if (nmine > 0) {   // every MMA retired
            const int jl = MT * nmine - 1;
            tc5::mbar_wait(bar(D_FULL + (jl & 1)), (jl >> 1) & 1);
        }
    } else {
        // ============================== epilogue: M-tile e of every tile ==============================
        const int e = (warp - 2) >> 2;
        const int row = (warp & 3) * 32 + lane;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        bf16 *y = reinterpret_cast<bf16 *>(a.y);
        const bf16 *res = reinterpret_cast<const bf16 *>(a.residual);
        const int act = a.act;
        tc5::pdl_wait();
        for (int j = e; j < MT * nmine; j += 2) {
            const int it = j / MT, m = j % MT;
            int x0, y0, n;
            tile_origin(it, x0, y0, n);
            const int p = m * 128 + row, by = p >> 5, bx = p & 31;
            const int gy = y0 + by, gx = x0 + bx;
            const bool ok = bx < TWO && gx < W && gy < H;
            const long long pix = ((long long)n * H + gy) * W + gx;
            const long long hw = (long long)H * W, ppix = (long long)n * 8 * hw + (long long)gy * W + gx;   // planar-8: [n][q][H][W][8]
            uint4 rv[8];
            if (res && ok) {   // residual of this pixel (laid out like x): in flight while the accumulator is waited for
                const uint4 *rp = reinterpret_cast<const uint4 *>(a.x_planar ? res + ppix * 8 : res + pix * a.r_cs + a.r_co + 64 * grp);
                const long long rstep = a.x_planar ? hw : 1;
#pragma unroll
                for (int q = 0; q < 8; ++q) rv[q] = rp[q * rstep];   // plain loads: y may be the residual tensor itself (in-place add)
            }
            tc5::mbar_wait(bar(D_FULL + e), (j >> 1) & 1);
            tc5::fence_after_sync();
            if constexpr (NOUT == 16) {
                // "rgb" form: 3 of the 16 accumulator columns + bias + x4 bilinear base -> fp32 NCHW (lanes = consecutive x: coalesced)
                uint32_t v[16];
                tc5::tmem_ld16(tmem + lane_base + e * NOUT, v);
                tc5::tmem_wait_ld();
                tc5::fence_before_sync();
                tc5::mbar_arrive_relaxed(bar(D_EMPTY + e));
                // output column o = lane: taps dx = 0, 1, 2 were accumulated at box columns o, o + 1, o + 2 (lanes of this warp = one box row)
                float s3[3];
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    s3[c] = __uint_as_float(v[c]) + __shfl_down_sync(0xffffffffu, __uint_as_float(v[3 + c]), 1) +
                            __shfl_down_sync(0xffffffffu, __uint_as_float(v[6 + c]), 2);
                if (ok) {
                    const int lh = H >> 2, lw = W >> 2;
                    const float *bp = a.base + (long long)n * a.base_nstride;
                    float *yp = reinterpret_cast<float *>(a.y) + (long long)n * a.y_nstride + (long long)gy * W + gx;
#pragma unroll
                    for (int c = 0; c < 3; ++c)
                        yp[c * hw] = s3[c] + bias_s[c] + bilinear_x4(bp + (long long)c * lh * lw, lh, lw, gy, gx);
                }
                continue;
            }
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {   // the accumulator row in two halves of 32 channels (register budget)
                uint32_t v[32];
                tc5::tmem_ld32(tmem + lane_base + e * NOUT + 32 * hh, v);
                tc5::tmem_wait_ld();
                if (hh == 1) {
                    tc5::fence_before_sync();
                    tc5::mbar_arrive_relaxed(bar(D_EMPTY + e));
                }
                if (ok && a.shuffle == 2) {
                    // PixelShuffle(2) in the store: channel co = 64 grp + c goes to sub-pixel (i, j) = ((c >> 1) & 1, c & 1), channel
                    // co / 4 -- this half's 32 channels are 8 contiguous channels (16 bytes) of each of the four HR pixels
#pragma unroll
                    for (int sub = 0; sub < 4; ++sub) {
                        const long long hp = ((long long)n * 2 * H + 2 * gy + (sub >> 1)) * (2 * W) + 2 * gx + (sub & 1);
                        float f[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) f[j] = apply_act(__uint_as_float(v[4 * j + sub]) + bias_s[32 * hh + 4 * j + sub], act);
                        // NHWC: 8 of the HR pixel's channels; planar-8 [n][cout/32][2H][2W][8]: plane 2 grp + hh of the HR image
                        bf16 *dst = a.y_planar ? y + (((long long)n * (a.cout >> 5) + 2 * grp + hh) * (4 * hw) + (hp - (long long)n * 4 * hw)) * 8
                                               : y + hp * a.y_cs + a.y_co + 16 * grp + 8 * hh;
                        *reinterpret_cast<uint4 *>(dst) =
                            make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
                    }
                } else if (ok) {
                    uint4 *yp = reinterpret_cast<uint4 *>(a.y_planar ? y + ppix * 8 : y + pix * a.y_cs + a.y_co + 64 * grp);
                    const long long ystep = a.y_planar ? hw : 1;
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4) {
                        const int q = 4 * hh + q4;
                        float f[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) f[j] = apply_act(__uint_as_float(v[q4 * 8 + j]) + bias_s[q * 8 + j], act);
                        if (res) {
                            const uint32_t *rw = reinterpret_cast<const uint32_t *>(&rv[q]);
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const float2 r2 = unpack_bf16x2(rw[j]);
                                f[2 * j] += r2.x, f[2 * j + 1] += r2.y;
                            }
                        }
                        yp[q * ystep] = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
                    }
                }
            }
        }
    }
    tc5::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc5::tmem_free(tmem, 128);
}

}  // namespace b200sr
