// wdsr_tc5_head.cuh -- WDSR-B head on tcgen05:  trunk = conv3x3(x - mean, Wh) + bh  (3 -> 24 channels)   models/basic_wdsr_b.py:86-87
//
// Same producer -> MMA -> epilogue chain as the tcgen05 tail (wdsr_tc5_tail.cuh), without TMA: builder warps stage (x - mean) as
// an NHWC4 bf16 tile and write the im2col rows of one 128-pixel M-tile (a pixel's 3x3 window x 4 channels = nine 8-byte loads,
// five 16-byte stores, K = 40); three tcgen05.mma (N = 32) per M-tile; the epilogue adds the bias and stores 48-byte NHWC pixels.
// Zero padding happens in the (x - mean) domain, exactly like the reference (pads are 0 after the mean subtraction).
#pragma once
#include "common.cuh"
#include "tc5.cuh"

namespace b200sr {

struct HeadTc5Layout {  // weight image: [4 groups][6 chunks][8 rows][8] bf16 | bias f32[32]
    int w, bias, total;
    __host__ __device__ HeadTc5Layout() { w = 0, bias = 4 * 6 * 128, total = bias + 128; }
};

namespace tc5head {
constexpr int TW = 32, TH = 8, NTHREADS = 320;
#ifndef B200SR_HEAD_CTAS
#define B200SR_HEAD_CTAS 3
#endif
constexpr int CTAS_PER_SM = B200SR_HEAD_CTAS;   // each CTA is a latency-bound builder -> MMA -> epilogue chain: co-resident CTAs fill each other's bubbles
constexpr int A_BUF = 6 * 2048;              // 5 window chunks + one zero chunk, [chunk][128 px][16 B]
constexpr int XW = TW + 2, XH = TH + 2;
constexpr int X4_BUF = XH * XW * 8;
constexpr int CTRL = 128;
enum Bar { A_FULL = 0, A_EMPTY = 2, D_FULL = 4, D_EMPTY = 6 };
__host__ __device__ inline size_t smem_bytes() { return (size_t)CTRL + 2 * A_BUF + X4_BUF + 64 + HeadTc5Layout().total; }
}  // namespace tc5head

template <typename TIN>
__global__ void __launch_bounds__(tc5head::NTHREADS, tc5head::CTAS_PER_SM)
wdsr_head_tc5_kernel(const TIN *__restrict__ x, bf16 *__restrict__ trunk, const uint8_t *__restrict__ wimg, int N, int H, int W, int tiles_x,
                     int tiles_y, int ntiles, float mean) {
    using namespace tc5head;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const HeadTc5Layout L;
    uint8_t *ctrl = smem_raw;
    uint8_t *ab = smem_raw + CTRL;                              // 2 x A_BUF
    uint8_t *x4 = ab + 2 * A_BUF;                               // X4_BUF (+ pad to 16)
    uint8_t *wsm = x4 + ((X4_BUF + 63) / 64) * 64;              // L.total
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bars = smem_u32(ctrl);
    auto bar = [&](int b) { return bars + 8u * (uint32_t)b; };
    if (tid == 0) {
        for (int e = 0; e < 2; ++e) {
            tc5::mbar_init(bar(A_FULL + e), 128);
            tc5::mbar_init(bar(A_EMPTY + e), 1);
            tc5::mbar_init(bar(D_FULL + e), 1);
            tc5::mbar_init(bar(D_EMPTY + e), 128);
        }
        tc5::mbar_init_fence();
    }
    __syncwarp();
    if (warp == 0) tc5::tmem_alloc(smem_u32(ctrl + 96), 64);
    for (int i = tid; i < L.total / 16; i += NTHREADS) cp_async16(wsm + i * 16, wimg + i * 16, 16);
    cp_async_commit();
    for (int i = tid; i < 2 * A_BUF / 16; i += NTHREADS) *reinterpret_cast<uint4 *>(ab + i * 16) = make_uint4(0u, 0u, 0u, 0u);
    cp_async_wait<0>();
    tc5::fence_proxy_async();
    tc5::fence_before_sync();
    __syncthreads();
    tc5::fence_after_sync();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t *>(ctrl + 96);
    const int nmine = (int)blockIdx.x < ntiles ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    auto tile_origin = [&](int it, int &x0, int &y0, int &n) {
        const int tile = blockIdx.x + it * gridDim.x;
        x0 = (tile % tiles_x) * TW;
        y0 = ((tile / tiles_x) % tiles_y) * TH;
        n = tile / (tiles_x * tiles_y);
    };

    if (warp == 1) {
        const bool leader = tc5::elect_one();
        const uint32_t idesc = tc5::idesc_bf16_f32(128, 32);
        const uint64_t bw = tc5::smem_desc(smem_u32(wsm) + L.w, 128, 6 * 128), a0 = tc5::smem_desc(smem_u32(ab), 2048, 128);
        for (int g = 0; g < 2 * nmine; ++g) {
            const int e = g & 1;
            tc5::mbar_wait(bar(A_FULL + e), (g >> 1) & 1);
            tc5::mbar_wait(bar(D_EMPTY + e), ((g >> 1) & 1) ^ 1);
            tc5::fence_after_sync();
            if (leader) {
                const uint64_t a = a0 + (uint64_t)((e * A_BUF) >> 4);
#pragma unroll
                for (int i = 0; i < 3; ++i) tc5::mma_ss(tmem + e * 32, a + (uint64_t)((2 * i * 2048) >> 4), bw + (uint64_t)(16 * i), idesc, i > 0);
                tc5::commit(bar(D_FULL + e));
                tc5::commit(bar(A_EMPTY + e));
            }
            __syncwarp();
        }
        if (nmine > 0) tc5::mbar_wait(bar(D_FULL + 1), ((2 * nmine - 1) >> 1) & 1);
    } else if (warp >= 2 && warp < 6) {
        const int bt = tid - 64;
        // The (x - mean) halo tile of the NEXT tile is fetched into registers while the current tile's two M-tiles are built: one
        // exposed global-memory latency per CTA instead of one per tile (it was ~2/3 of this kernel's time).
        constexpr int NIT = (XH * XW + 127) / 128;
        float v[NIT][3];
        auto fetch = [&](int it) {
            int x0, y0, n;
            tile_origin(it, x0, y0, n);
#pragma unroll
            for (int k = 0; k < NIT; ++k) {
                const int i = bt + 128 * k;
                const int gy = y0 - 1 + i / XW, gx = x0 - 1 + i % XW;
                const bool ok = i < XH * XW && gy >= 0 && gy < H && gx >= 0 && gx < W;
                const long long o = ok ? (((long long)n * 3) * H + gy) * W + gx : 0;
#pragma unroll
                for (int c = 0; c < 3; ++c) v[k][c] = ok ? to_f32<TIN>(x[o + (long long)c * H * W]) - mean : 0.f;
            }
        };
        if (nmine > 0) fetch(0);
        for (int g = 0; g < 2 * nmine; ++g) {
            const int it = g >> 1, h = g & 1, e = g & 1;
            if (h == 0) {
                asm volatile("bar.sync 1, 128;" ::: "memory");  // everyone finished reading the previous tile's x4
#pragma unroll
                for (int k = 0; k < NIT; ++k) {
                    const int i = bt + 128 * k;
                    if (i < XH * XW) *reinterpret_cast<uint2 *>(x4 + i * 8) = make_uint2(pack_bf16x2(v[k][0], v[k][1]), pack_bf16x2(v[k][2], 0.f));
                }
                if (it + 1 < nmine) fetch(it + 1);
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
            tc5::mbar_wait(bar(A_EMPTY + e), ((g >> 1) & 1) ^ 1);
            const int ly = 4 * h + (bt >> 5), lx = bt & 31;
            uint2 wv[10];
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) wv[ky * 3 + kx] = *reinterpret_cast<const uint2 *>(x4 + ((ly + ky) * XW + lx + kx) * 8);
            wv[9] = make_uint2(0u, 0u);
            uint8_t *dst = ab + e * A_BUF + bt * 16;
#pragma unroll
            for (int j = 0; j < 5; ++j) *reinterpret_cast<uint4 *>(dst + j * 2048) = make_uint4(wv[2 * j].x, wv[2 * j].y, wv[2 * j + 1].x, wv[2 * j + 1].y);
            tc5::fence_proxy_async();
            tc5::mbar_arrive(bar(A_FULL + e));
        }
    } else if (warp >= 6) {
        const int row = (warp & 3) * 32 + lane;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const float *bias = reinterpret_cast<const float *>(wsm + L.bias);
        for (int g = 0; g < 2 * nmine; ++g) {
            const int it = g >> 1, h = g & 1, e = g & 1;
            int x0, y0, n;
            tile_origin(it, x0, y0, n);
            tc5::mbar_wait(bar(D_FULL + e), (g >> 1) & 1);
            tc5::fence_after_sync();
            uint32_t v[32];
            tc5::tmem_ld32(tmem + lane_base + e * 32, v);
            tc5::tmem_wait_ld();
            tc5::fence_before_sync();
            tc5::mbar_arrive_relaxed(bar(D_EMPTY + e));
            const int gy = y0 + 4 * h + (row >> 5), gx = x0 + (row & 31);
            if (gy < H && gx < W) {
                bf16 *o = trunk + (((long long)n * 3 * H + gy) * W + gx) * 8;   // planar-8 trunk [N][3][H][W][8] (tma_map.h)
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    const float4 ba = *reinterpret_cast<const float4 *>(bias + q * 8), bb = *reinterpret_cast<const float4 *>(bias + q * 8 + 4);
                    uint4 ov;
                    ov.x = pack_bf16x2(__uint_as_float(v[q * 8 + 0]) + ba.x, __uint_as_float(v[q * 8 + 1]) + ba.y);
                    ov.y = pack_bf16x2(__uint_as_float(v[q * 8 + 2]) + ba.z, __uint_as_float(v[q * 8 + 3]) + ba.w);
                    ov.z = pack_bf16x2(__uint_as_float(v[q * 8 + 4]) + bb.x, __uint_as_float(v[q * 8 + 5]) + bb.y);
                    ov.w = pack_bf16x2(__uint_as_float(v[q * 8 + 6]) + bb.z, __uint_as_float(v[q * 8 + 7]) + bb.w);
                    *reinterpret_cast<uint4 *>(o + (long long)q * H * W * 8) = ov;   // 32 lanes = 512 contiguous bytes
                }
            }
        }
    }
    tc5::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc5::tmem_free(tmem, 64);
}

}  // namespace b200sr
