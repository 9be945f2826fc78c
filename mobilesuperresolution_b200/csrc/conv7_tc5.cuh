// conv7_tc5.cuh -- 7x7 "same" convolution, bf16, on tcgen05: SPyNet's BasicModule layers 8->32, 32->64, 64->32, 32->16
// (models/spynet_arch.py:17-22; 97 % of SPyNet's FLOPs -- the 16->2 flow head stays on the mma.sync kernel of conv.cuh).
//
//   y[n, oy, ox, co] = act( bias[co] + sum_{ky,kx,ci} x[n, oy+ky-3, ox+kx-3, ci] * w[co][ci][ky][kx] )
//
// Implicit GEMM with pixels as M, like conv_tc5.cuh, but a 7x7 filter bank (64 x 32 x 49 bf16 = 200 KB) does not fit next to the
// activation tile, so the filters are STREAMED: one tap row (7 taps, <= 28 KB) per stage of a 3-stage ring, and each stage is used by
// the four M-tiles of a 26 x 16 pixel output tile before it is released (four accumulators per tile, two tiles in flight in TMEM).
//   warp 0      TMA      NCH cp.async.bulk.tensor.4d per tile (one per 8-channel chunk) of a 32 x 22 pixel box (26 x 16 outputs +
//                        3-pixel halo) into chunk-planar shared memory [chunk][pixel][16 B]; out-of-image pixels and channels past
//                        cin are zero-filled by the TMA unit.  x is NHWC or planar-8 [n][c/8][h][w][8] (SPyNet's private tensors).
//   warp 1      weights  one cp.async.bulk (global -> shared, contiguous) per tap row into the ring, free running
//   warp 2      MMA      per tap row: 4 M-tiles x NG tap groups x NCH/2 tcgen05.mma (M = 128, N = G cout, K = 16; Cfg): a tap group is a
//                        constant pixel offset of the A operand, the two chunks of a K step are paired through the LBO stride
//   warps 4-7            epilogue: tcgen05.ld -> + bias -> activation -> bf16 stores (NHWC pixel rows or planar-8 planes)
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "conv.cuh"
#include "tc5.cuh"

namespace b200sr {
namespace tc5conv7 {
constexpr int TWO = 26, TH = 16, BW = 32, BH = TH + 6, NMT = 4, NTHREADS = 256, NSTAGE = 3;
constexpr int PLANE_PX = BW * BH + 8;            // + 8 zero pixels: the last taps of the last M-tile read past the box
constexpr int PLANE = PLANE_PX * 16;             // 11,392 B
constexpr int CTRL = 256;
enum Bar { TILE_FULL = 0, TILE_EMPTY = 1, W_FULL = 2 /*3*/, W_EMPTY = 5 /*3*/, D_FULL = 8 /*2*/, D_EMPTY = 10 /*2*/, NBARS = 12 };
// G horizontal taps ride in the N dimension (an MMA of N <= 64 costs ~41-48 clk whatever N is: its 4 KB A fetch bounds it): accumulator
// column b * NOUT + co of box column r = sum over tap groups j and input channels of x[r + G j] * w[co][ci][ky][G j + b], and the epilogue
// forms out[o] = sum_b D_b[o + b] with G - 1 shuffles per value (a warp = one 32-pixel box row, outputs at its first 26 columns).
// 64 -> 32: G = 2 (28 -> 16 MMAs per tap row and M-tile); 32 -> 16: G = 4 (14 -> 4); 64 outputs: G = 1 (TMEM is full at 2 x 4 x 64).
template <int NCH, int NOUT> struct Cfg {
    static constexpr int G = NOUT == 32 ? 2 : NOUT == 16 ? 4 : 1;
    static constexpr int NG = (7 + G - 1) / G;                // tap groups per row (the last one padded with zero weights)
    static constexpr int ND = G * NOUT;                       // accumulator columns per M-tile (64 for every layer)
    static constexpr int TILE_BUF = NCH * PLANE;
    static constexpr int W_SBO = NG * NCH * 128;              // stage image [ND/8 row groups][NG tap groups x NCH chunks][8 rows][16 B]
    static constexpr int STAGE = (ND / 8) * W_SBO;            // <= 32,768 B
    static constexpr int TMEM_COLS = 2 * NMT * ND <= 128 ? 128 : 2 * NMT * ND <= 256 ? 256 : 512;
    static constexpr size_t smem_bytes() { return (size_t)CTRL + TILE_BUF + NSTAGE * STAGE + 256; }
};
}  // namespace tc5conv7

__device__ __forceinline__ void bulk_load(uint32_t dst_saddr, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_saddr), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tma_load_4d_c7(uint32_t dst_saddr, const void *tmap, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst_saddr), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

template <int NCH, int NOUT>
__global__ void __launch_bounds__(tc5conv7::NTHREADS, 1)
conv7x7_tc5_kernel(const __grid_constant__ CUtensorMap tmap_x, ConvArgs a, const uint8_t *__restrict__ wimg, int tiles_x, int tiles_y,
                   int ntiles) {
    using namespace tc5conv7;
    using C = Cfg<NCH, NOUT>;
    constexpr int TILE_BUF = C::TILE_BUF, W_SBO = C::W_SBO, STAGE = C::STAGE, G = C::G, NG = C::NG, ND = C::ND;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *ctrl = smem_raw;
    uint8_t *tc = smem_raw + CTRL;           // TILE_BUF
    uint8_t *wsm = tc + TILE_BUF;            // NSTAGE x STAGE
    float *bias_s = reinterpret_cast<float *>(wsm + NSTAGE * STAGE);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bars = smem_u32(ctrl);
    auto bar = [&](int b) { return bars + 8u * (uint32_t)b; };
    const uint32_t tc_u = smem_u32(tc), w_u = smem_u32(wsm);
    const int H = a.h, W = a.w_;

    // programmatic stream serialization: set-up overlaps the previous kernel's tail; activation reads sit behind griddepcontrol.wait
    tc5::pdl_launch_dependents();
    if (tid == 0) {
        tc5::mbar_init(bar(TILE_FULL), 1);
        tc5::mbar_init(bar(TILE_EMPTY), 1);
        for (int s = 0; s < NSTAGE; ++s) {
            tc5::mbar_init(bar(W_FULL + s), 1);
            tc5::mbar_init(bar(W_EMPTY + s), 1);
        }
        for (int e = 0; e < 2; ++e) {
            tc5::mbar_init(bar(D_FULL + e), 1);
            tc5::mbar_init(bar(D_EMPTY + e), 128);
        }
        tc5::mbar_init_fence();
        tc5::tma_prefetch_desc(&tmap_x);
    }
    __syncwarp();
    if (warp == 0) tc5::tmem_alloc(smem_u32(ctrl + 240), C::TMEM_COLS);
    if (tid < NOUT) bias_s[tid] = a.bias[tid];
    for (int i = tid; i < NCH * 8; i += NTHREADS)   // the 8 pad pixels of every plane stay zero (TMA never writes them)
        *reinterpret_cast<uint4 *>(tc + (i / 8) * PLANE + (BW * BH + i % 8) * 16) = make_uint4(0u, 0u, 0u, 0u);
    tc5::fence_proxy_async();
    tc5::fence_before_sync();
    __syncthreads();
    tc5::fence_after_sync();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t *>(ctrl + 240);
    const int nmine = (int)blockIdx.x < ntiles ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    auto tile_origin = [&](int it, int &x0, int &y0, int &n) {
        const int tile = blockIdx.x + it * gridDim.x;
        x0 = (tile % tiles_x) * TWO;
        y0 = ((tile / tiles_x) % tiles_y) * TH;
        n = tile / (tiles_x * tiles_y);
    };

    if (warp == 0) {
        // ============================== activation tile producer ==============================
        if (tc5::elect_one()) {
            tc5::pdl_wait();
            for (int it = 0; it < nmine; ++it) {
                int x0, y0, n;
                tile_origin(it, x0, y0, n);
                tc5::mbar_wait(bar(TILE_EMPTY), (it & 1) ^ 1);
                tc5::mbar_arrive_expect_tx(bar(TILE_FULL), NCH * BW * BH * 16);
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    const uint32_t dst = tc_u + c * PLANE;
                    if (a.x_planar) tma_load_4d_c7(dst, &tmap_x, bar(TILE_FULL), 4 * (x0 - 3), y0 - 3, c, n);   // (uint32 of a row, row, plane, image)
                    else tma_load_4d_c7(dst, &tmap_x, bar(TILE_FULL), 8 * c, x0 - 3, y0 - 3, n);                 // (channel, x, y, image)
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ============================== filter tap-row producer (constants: no dependence on the previous kernel) ==============================
        if (tc5::elect_one()) {
            for (int g = 0; g < 7 * nmine; ++g) {
                const int s = g % NSTAGE;
                tc5::mbar_wait(bar(W_EMPTY + s), ((g / NSTAGE) & 1) ^ 1);
                tc5::mbar_arrive_expect_tx(bar(W_FULL + s), STAGE);
                bulk_load(w_u + s * STAGE, wimg + (size_t)(g % 7) * STAGE, STAGE, bar(W_FULL + s));
            }
        }
        __syncwarp();
    } else if (warp == 2) {
        // ============================== MMA issuer ==============================
        const bool leader = tc5::elect_one();
        const uint32_t idesc = tc5::idesc_bf16_f32(128, ND);
        const uint64_t a0d = tc5::smem_desc(tc_u, PLANE, 128);   // A: chunk pairs through LBO = plane stride
        for (int it = 0; it < nmine; ++it) {
            const int set = it & 1;
            tc5::mbar_wait(bar(TILE_FULL), it & 1);
            tc5::mbar_wait(bar(D_EMPTY + set), ((it >> 1) & 1) ^ 1);
            for (int ky = 0; ky < 7; ++ky) {
                const int g = it * 7 + ky, s = g % NSTAGE;
                tc5::mbar_wait(bar(W_FULL + s), (g / NSTAGE) & 1);
                tc5::fence_after_sync();
                if (leader) {
                    const uint64_t bw = tc5::smem_desc(w_u + s * STAGE, 128, W_SBO);
#pragma unroll 1
                    for (int m = 0; m < NMT; ++m) {
                        const uint32_t d = tmem + (set * NMT + m) * ND;
                        const uint64_t abase = a0d + (uint64_t)((((4 * m + ky) * BW) * 16) >> 4);
#pragma unroll
                        for (int i = 0; i < NG * (NCH / 2); ++i) {   // (tap group j = taps G j .. G j + G - 1, chunks 2 cp, 2 cp + 1)
                            const int j = i / (NCH / 2), cp = i % (NCH / 2);
                            const int aoff = 2 * cp * PLANE + G * j * 16;
                            tc5::mma_ss(d, abase + (uint64_t)(aoff >> 4), bw + (uint64_t)(8 * (j * NCH + 2 * cp)), idesc, (ky | i) != 0);
                        }
                    }
                    tc5::commit(bar(W_EMPTY + s));
                    if (ky == 6) {
                        tc5::commit(bar(D_FULL + set));
                        tc5::commit(bar(TILE_EMPTY));
                    }
                }
                __syncwarp();
            }
        }
        if (nmine > 0) tc5::mbar_wait(bar(D_FULL + ((nmine - 1) & 1)), ((nmine - 1) >> 1) & 1);  // every MMA retired
    } else if (warp >= 4) {
        // ============================== epilogue ==============================
        const int row = (warp & 3) * 32 + lane;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        bf16 *y = reinterpret_cast<bf16 *>(a.y);
        const int act = a.act;
        const long long hw = (long long)H * W;
        for (int it = 0; it < nmine; ++it) {
            int x0, y0, n;
            tile_origin(it, x0, y0, n);
            const int set = it & 1;
            tc5::mbar_wait(bar(D_FULL + set), (it >> 1) & 1);
            tc5::fence_after_sync();
#pragma unroll 1
            for (int m = 0; m < NMT; ++m) {
                const int p = m * 128 + row, by = p >> 5, bx = p & 31;
                const int gy = y0 + by, gx = x0 + bx;
                const bool ok = bx < TWO && gx < W && gy < H;
                const long long pix = ((long long)n * H + gy) * W + gx;
                const long long ppix = (long long)n * (NOUT / 8) * hw + (long long)gy * W + gx;   // planar-8: [n][q][H][W][8]
                constexpr int HC = NOUT < 32 ? NOUT : 32;   // channels per tcgen05.ld
#pragma unroll
                for (int hh = 0; hh < NOUT / HC; ++hh) {
                    uint32_t v[HC];
                    const uint32_t taddr = tmem + lane_base + (set * NMT + m) * ND + HC * hh;
                    if constexpr (HC == 32) tc5::tmem_ld32(taddr, v);
                    else tc5::tmem_ld16(taddr, v);
                    tc5::tmem_wait_ld();
#pragma unroll
                    for (int b = 1; b < G; ++b) {   // + taps G j + b, accumulated b box columns to the right (all 32 lanes shuffle)
                        uint32_t vb[HC];
                        if constexpr (HC == 32) tc5::tmem_ld32(taddr + b * NOUT, vb);
                        else tc5::tmem_ld16(taddr + b * NOUT, vb);
                        tc5::tmem_wait_ld();
#pragma unroll
                        for (int q = 0; q < HC; ++q) v[q] = __float_as_uint(__uint_as_float(v[q]) + __shfl_down_sync(0xffffffffu, __uint_as_float(vb[q]), b));
                    }
                    if (ok) {
                        uint4 *yp = reinterpret_cast<uint4 *>(a.y_planar ? y + ppix * 8 : y + pix * a.y_cs + a.y_co);
                        const long long ystep = a.y_planar ? hw : 1;
#pragma unroll
                        for (int q4 = 0; q4 < HC / 8; ++q4) {
                            const int q = (HC / 8) * hh + q4;
                            float f[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) f[j] = apply_act(__uint_as_float(v[q4 * 8 + j]) + bias_s[q * 8 + j], act);
                            yp[q * ystep] = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
                        }
                    }
                }
            }
            tc5::fence_before_sync();
            tc5::mbar_arrive_relaxed(bar(D_EMPTY + set));
        }
    }
    tc5::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc5::tmem_free(tmem, C::TMEM_COLS);
}

}  // namespace b200sr
