// wdsr_tc5_tail.cuh -- fused WDSR-B tail on tcgen05:
//   y = PixelShuffle_s( conv3x3(trunk, Wt) + conv5x5(x - mean, Ws) + (bt + bs) ) + out_add          models/basic_wdsr_b.py:90-92
//
// One persistent CTA per SM (see NBUF below for the two-CTA experiment), 32 x 8 LR-pixel tiles = two M-tiles of 128 pixels (4 rows x 32).  No stage depends on another
// CTA-local result, so the pipeline is a plain producer -> MMA -> epilogue chain:
//   warp 0      TMA      nine cp.async.bulk.tensor.5d per tile: three x-shifted copies (one per horizontal tap) of the three
//                        8-channel planes of the trunk tile + 1-row halo, so every 3x3 tap is a constant address offset;
//                        out-of-image pixels are zero-filled by the TMA unit (= the conv's zero padding); double buffered
//   warps 2-5   builder  stages (x - mean) as an NHWC4 bf16 tile (the next tile's pixels are prefetched into registers), then writes the
//                        skip operand ONCE PER TILE: three x-shifted copies (shift 0, 2, 4) of the tile + 2-row halo whose 16-byte
//                        entries hold TWO horizontally adjacent pixels [x | x+1] x 4 channels, so a K = 8 chunk covers two taps of a
//                        window row and every (ky, tap pair) is a constant address offset -- no im2col.  (The first form built a K = 104
//                        im2col per M-tile, 25 loads + 13 stores per pixel: ncu showed the builder warps 94 % busy and everybody else
//                        waiting for them, profiles/r02_tail_head_ncu.md.)
//   warp 1      MMA      14 (3x3: 27 (tap, chunk) slices paired through the LBO stride) + 8 (skip: 5 window rows x 3 tap pairs = 15
//                        chunks) tcgen05.mma, N = 3 s^2 padded to 16
//   warps 6-9   epilogue tcgen05.ld -> + bias + mean -> PixelShuffle store: lane = LR pixel, 32 lanes x s outputs are one
//                        contiguous segment of an HR row
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "tc5.cuh"

namespace b200sr {

struct TailTc5Layout {  // weight image (bytes); rows = output channel (NOP = 3 s^2 padded to 16)
    int wt, ws, bias, total, sbo_t, sbo_s;
    __host__ __device__ TailTc5Layout(int NOP) {
        sbo_t = 28 * 128;  // 27 (dx, c, dy) chunks + zero chunk
        sbo_s = 16 * 128;  // 15 (window row, tap pair) chunks + zero chunk: 8 instructions x 2 halves
        wt = 0;
        ws = wt + (NOP / 8) * sbo_t;
        bias = ws + (NOP / 8) * sbo_s;
        total = bias + NOP * 4;
    }
};

namespace tc5tail {
// NBUF = trunk-tile / skip-operand buffers per CTA.  2 (default): one CTA per SM, double buffered.  1 (-DB200SR_TAIL_NBUF=1 through
// B200SR_NVCC_EXTRA): TWO co-resident CTAs per SM, each single buffered (102 KB of shared memory, 128 of the 512 TMEM columns, <= 96
// registers) -- the test of "a second, independent TMA -> builder -> MMA -> epilogue chain on the same SM fills the first one's bubbles".
// Measured, parity-green: 34.3 us against 30.3 us per launch at cfg2 (graph replay): SLOWER, like every deeper ring inside one chain was
// neutral -- the tail is not bound by the latency of its chain (profiles/r02_tail_head_ncu.md).
#ifndef B200SR_TAIL_NBUF
#define B200SR_TAIL_NBUF 2
#endif
constexpr int NBUF = B200SR_TAIL_NBUF, CTAS_PER_SM = NBUF == 1 ? 2 : 1;
__host__ __device__ constexpr int buf_of(int it) { return NBUF == 2 ? (it & 1) : 0; }
__host__ __device__ constexpr int phase_of(int it) { return NBUF == 2 ? ((it >> 1) & 1) : (it & 1); }
constexpr int TW = 32, TH = 8, NTHREADS = 320;
constexpr int PLANE = (TH + 2) * TW * 16;  // 5,120 B: 10 rows x 32 px x 16 B
constexpr int TC_BUF = 9 * PLANE;          // 46,080 B per tile (3 copies x 3 planes)
constexpr int XW = TW + 5, XH = TH + 4;    // x tile with 2-pixel halo (+ 1 column: the zero-weight partner of the fifth tap), 8 bytes per pixel
constexpr int X4_BUF = (XH * XW * 8 + 15) / 16 * 16;
constexpr int SK_COPY = XH * TW * 16;      // 6,144 B: one x-shifted copy, 12 rows x 32 entries x 16 B ([pixel | right neighbour] x 4 channels)
constexpr int SK_BUF = 3 * SK_COPY;        // 18,432 B per tile
constexpr int SK_STRIDE = SK_BUF;
constexpr int CTRL = 256;
enum Bar { TC_FULL = 0, TC_EMPTY = 2, SK_FULL = 4, SK_EMPTY = 6, D_FULL = 8, D_EMPTY = 10, NBARS = 12 };
__host__ __device__ inline size_t smem_bytes(int NOP) {
    return (size_t)CTRL + NBUF * TC_BUF + NBUF * SK_STRIDE + X4_BUF + (size_t)TailTc5Layout(NOP).total;
}
}  // namespace tc5tail

template <typename TIN, typename TOUT, int S>
__global__ void __launch_bounds__(tc5tail::NTHREADS, tc5tail::CTAS_PER_SM)
wdsr_tail_tc5_kernel(const __grid_constant__ CUtensorMap tmap_trunk, const TIN *__restrict__ x, TOUT *__restrict__ y,
                     const uint8_t *__restrict__ wimg, int N, int H, int W, int tiles_x, int tiles_y, int ntiles, float mean, float out_add) {
    using namespace tc5tail;
    constexpr int NO = 3 * S * S, NOP = round_up(NO, 16);
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const TailTc5Layout L(NOP);
    uint8_t *ctrl = smem_raw;
    uint8_t *tc = smem_raw + CTRL;          // NBUF x TC_BUF
    uint8_t *sk = tc + NBUF * TC_BUF;       // NBUF x SK_STRIDE
    uint8_t *x4 = sk + NBUF * SK_STRIDE;    // X4_BUF
    uint8_t *wsm = x4 + X4_BUF;          // L.total
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bars = smem_u32(ctrl);
    auto bar = [&](int b) { return bars + 8u * (uint32_t)b; };
    const uint32_t tc_u = smem_u32(tc), sk_u = smem_u32(sk), w_u = smem_u32(wsm);

    if (tid == 0) {
        for (int e = 0; e < 2; ++e) {
            tc5::mbar_init(bar(TC_FULL + e), 1);
            tc5::mbar_init(bar(TC_EMPTY + e), 1);
            tc5::mbar_init(bar(SK_FULL + e), 128);
            tc5::mbar_init(bar(SK_EMPTY + e), 1);
            tc5::mbar_init(bar(D_FULL + e), 1);
            tc5::mbar_init(bar(D_EMPTY + e), 128);
        }
        tc5::mbar_init_fence();
        tc5::tma_prefetch_desc(&tmap_trunk);
    }
    __syncwarp();
    if (warp == 0) tc5::tmem_alloc(smem_u32(ctrl + 240), 128);
    for (int i = tid; i < L.total / 16; i += NTHREADS) cp_async16(wsm + i * 16, wimg + i * 16, 16);
    cp_async_commit();
    for (int i = tid; i < (NBUF * SK_STRIDE + X4_BUF) / 16; i += NTHREADS) *reinterpret_cast<uint4 *>(sk + i * 16) = make_uint4(0u, 0u, 0u, 0u);
    // the zero-weight dummy half of the 14th 3x3 instruction reads one row past the last plane of a buffer: keep it finite
    // before buffer 1 has ever been loaded (the last buffer overflows into the zeroed -- later: finite -- skip area)
    if (NBUF == 2)
        for (int i = tid; i < 512 / 16; i += NTHREADS) *reinterpret_cast<uint4 *>(tc + TC_BUF + i * 16) = make_uint4(0u, 0u, 0u, 0u);
    cp_async_wait<0>();
    tc5::fence_proxy_async();
    tc5::fence_before_sync();
    __syncthreads();
    tc5::fence_after_sync();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t *>(ctrl + 240);
    const int nmine = (int)blockIdx.x < ntiles ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    auto tile_origin = [&](int it, int &x0, int &y0, int &n) {
        const int tile = blockIdx.x + it * gridDim.x;
        x0 = (tile % tiles_x) * TW;
        y0 = ((tile / tiles_x) % tiles_y) * TH;
        n = tile / (tiles_x * tiles_y);
    };

    if (warp == 0) {
        // ============================== TMA producer ==============================
        if (tc5::elect_one()) {
            for (int it = 0; it < nmine; ++it) {
                int x0, y0, n;
                tile_origin(it, x0, y0, n);
                const int b = buf_of(it);
                tc5::mbar_wait(bar(TC_EMPTY + b), phase_of(it) ^ 1);
                tc5::mbar_arrive_expect_tx(bar(TC_FULL + b), TC_BUF);
#pragma unroll
                for (int d = 0; d < 3; ++d)
#pragma unroll
                    for (int c = 0; c < 3; ++c)
                        tc5::tma_load_plane(tc_u + b * TC_BUF + (d * 3 + c) * PLANE, &tmap_trunk, bar(TC_FULL + b), x0 - 1 + d, y0 - 1, c, n);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ============================== MMA issuer ==============================
        const bool leader = tc5::elect_one();
        const uint32_t idesc = tc5::idesc_bf16_f32(128, NOP);
        const uint64_t bwt = tc5::smem_desc(w_u + L.wt, 128, L.sbo_t), bws = tc5::smem_desc(w_u + L.ws, 128, L.sbo_s);
        const uint64_t at0 = tc5::smem_desc(tc_u, 0, 128), as0 = tc5::smem_desc(sk_u, 0, 128);
        for (int g = 0; g < 2 * nmine; ++g) {
            const int it = g >> 1, h = g & 1, b = buf_of(it), e = g & 1;
            if (h == 0) {
                tc5::mbar_wait(bar(TC_FULL + b), phase_of(it));
                tc5::mbar_wait(bar(SK_FULL + b), phase_of(it));
            }
            tc5::mbar_wait(bar(D_EMPTY + e), ((g >> 1) & 1) ^ 1);
            tc5::fence_after_sync();
            if (leader) {
                const uint32_t d = tmem + e * 64;
                const uint64_t abase = at0 + (uint64_t)((b * TC_BUF + 4 * h * 512) >> 4);
#pragma unroll
                for (int i = 0; i < 14; ++i) {  // chunk order q = (dx * 3 + c) * 3 + dy  ->  offset (dx*3+c) * PLANE + dy * 512
                    const int q0 = 2 * i, q1 = 2 * i + 1;
                    const int a0 = (q0 / 3) * PLANE + (q0 % 3) * 512;
                    const int a1 = q1 < 27 ? (q1 / 3) * PLANE + (q1 % 3) * 512 : a0 + 512;
                    tc5::mma_ss(d, abase + (uint64_t)(a0 >> 4) + ((uint64_t)((a1 - a0) >> 4) << 16), bwt + (uint64_t)(16 * i), idesc, i > 0);
                }
                // skip 5x5: chunk (ky, j) = copy j (shift 2j) at window row ky.  Instructions 0..4: (ky, 0) | (ky, 1) paired through the
                // copy stride; 5..7: (0,2) | (1,2), (2,2) | (3,2), (4,2) | zero weights, paired through the row stride
                const uint64_t sbase = as0 + (uint64_t)((b * SK_STRIDE + 4 * h * 512) >> 4);
#pragma unroll
                for (int i = 0; i < 5; ++i) tc5::mma_ss(d, sbase + (uint64_t)((i * 512) >> 4) + ((uint64_t)(SK_COPY >> 4) << 16), bws + (uint64_t)(16 * i), idesc, true);
#pragma unroll
                for (int i = 0; i < 3; ++i)
                    tc5::mma_ss(d, sbase + (uint64_t)((2 * SK_COPY + 2 * i * 512) >> 4) + ((uint64_t)(512 >> 4) << 16), bws + (uint64_t)(16 * (5 + i)), idesc, true);
                tc5::commit(bar(D_FULL + e));
                if (h == 1) {
                    tc5::commit(bar(SK_EMPTY + b));
                    tc5::commit(bar(TC_EMPTY + b));
                }
            }
            __syncwarp();
        }
        if (nmine > 0) tc5::mbar_wait(bar(D_FULL + 1), ((2 * nmine - 1) >> 1) & 1);  // every MMA retired
    } else if (warp < 6) {
        // ============================== builders: x - mean tile, pair-packed shifted copies for the skip ==============================
        const int bt = tid - 64;  // 0..127
        constexpr int NIT = (XH * XW + 127) / 128;
        float v[NIT][3];
        auto fetch = [&](int it) {   // global loads of tile `it`'s (x - mean) pixels into registers (consumed one tile later)
            int x0, y0, n;
            tile_origin(it, x0, y0, n);
#pragma unroll
            for (int k = 0; k < NIT; ++k) {
                const int i = bt + 128 * k;
                const int gy = y0 - 2 + i / XW, gx = x0 - 2 + i % XW;
                const bool ok = i < XH * XW && gy >= 0 && gy < H && gx >= 0 && gx < W;
                const long long o = ok ? (((long long)n * 3) * H + gy) * W + gx : 0;
#pragma unroll
                for (int c = 0; c < 3; ++c) v[k][c] = ok ? to_f32<TIN>(x[o + (long long)c * H * W]) - mean : 0.f;
            }
        };
        if (nmine > 0) fetch(0);
        for (int it = 0; it < nmine; ++it) {
            const int b = buf_of(it);
#pragma unroll
            for (int k = 0; k < NIT; ++k) {
                const int i = bt + 128 * k;
                if (i < XH * XW) *reinterpret_cast<uint2 *>(x4 + i * 8) = make_uint2(pack_bf16x2(v[k][0], v[k][1]), pack_bf16x2(v[k][2], 0.f));
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (it + 1 < nmine) fetch(it + 1);       // in flight while this tile's operand is written
            tc5::mbar_wait(bar(SK_EMPTY + b), phase_of(it) ^ 1);
            uint8_t *dstb = sk + b * SK_STRIDE;
#pragma unroll
            for (int k = 0; k < 3 * XH * TW / 128; ++k) {   // 9 entries per thread: copy j, row yy, column xx
                const int e = bt + 128 * k, j = e / (XH * TW), r = e % (XH * TW), yy = r / TW, xx = r % TW;
                const uint8_t *src = x4 + (yy * XW + xx + 2 * j) * 8;
                const uint2 lo = *reinterpret_cast<const uint2 *>(src), hi = *reinterpret_cast<const uint2 *>(src + 8);
                *reinterpret_cast<uint4 *>(dstb + j * SK_COPY + r * 16) = make_uint4(lo.x, lo.y, hi.x, hi.y);
            }
            tc5::fence_proxy_async();
            tc5::mbar_arrive(bar(SK_FULL + b));
            asm volatile("bar.sync 1, 128;" ::: "memory");  // everyone finished reading x4 before the next tile overwrites it
        }
    } else {
        // ============================== epilogue ==============================
        const int row = (warp & 3) * 32 + lane;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const float *bias = reinterpret_cast<const float *>(wsm + L.bias);
        const int OH = S * H, OW = S * W;
        int x0 = 0, y0 = 0, n = 0;
        for (int g = 0; g < 2 * nmine; ++g) {
            const int it = g >> 1, h = g & 1, e = g & 1;
            if (h == 0) tile_origin(it, x0, y0, n);   // (two integer divisions: once per tile, not per M-tile)
            tc5::mbar_wait(bar(D_FULL + e), (g >> 1) & 1);
            tc5::fence_after_sync();
            uint32_t v[NOP];
#pragma unroll
            for (int c = 0; c < NOP; c += 16) tc5::tmem_ld16(tmem + lane_base + e * 64 + c, *reinterpret_cast<uint32_t(*)[16]>(&v[c]));
            tc5::tmem_wait_ld();
            tc5::fence_before_sync();
            tc5::mbar_arrive_relaxed(bar(D_EMPTY + e));
            const int gy = y0 + 4 * h + (row >> 5), gx = x0 + (row & 31);
            if (gy < H && gx < W) {
                TOUT *o0 = y + (((long long)n * 3) * OH + S * gy) * OW + S * gx;   // (c, i) add constant strides to one base
                const long long cstride = (long long)OH * OW;
#pragma unroll
                for (int c = 0; c < 3; ++c)
#pragma unroll
                    for (int i = 0; i < S; ++i) {
                        TOUT *o = o0 + c * cstride + i * OW;
                        float r[S];
                        if constexpr (S == 4) {
                            const float4 bb = *reinterpret_cast<const float4 *>(bias + c * 16 + i * 4);  // broadcast read
                            r[0] = __uint_as_float(v[c * 16 + i * 4 + 0]) + bb.x + out_add;
                            r[1] = __uint_as_float(v[c * 16 + i * 4 + 1]) + bb.y + out_add;
                            r[2] = __uint_as_float(v[c * 16 + i * 4 + 2]) + bb.z + out_add;
                            r[3] = __uint_as_float(v[c * 16 + i * 4 + 3]) + bb.w + out_add;
                        } else {
#pragma unroll
                            for (int j = 0; j < S; ++j) {
                                const int ch = c * S * S + i * S + j;
                                r[j] = __uint_as_float(v[ch]) + bias[ch] + out_add;
                            }
                        }
                        if constexpr (sizeof(TOUT) == 1) {
                            // 8-bit frame: (sr * 255).round().clamp(0, 255)   common/metrics.py:12 (round half to even, like torch.round)
                            uint32_t pk = 0;
#pragma unroll
                            for (int j = 0; j < S; ++j) pk |= (uint32_t)min(max(__float2int_rn(r[j] * 255.f), 0), 255) << (8 * j);
                            if constexpr (S == 4) *reinterpret_cast<uint32_t *>(o) = pk;
                            else if constexpr (S == 2) *reinterpret_cast<uint16_t *>(o) = (uint16_t)pk;
                            else {
#pragma unroll
                                for (int j = 0; j < S; ++j) o[j] = (TOUT)((pk >> (8 * j)) & 255u);
                            }
                        } else if constexpr (S == 4 && sizeof(TOUT) == 2) {
                            *reinterpret_cast<uint2 *>(o) = make_uint2(pack_bf16x2(r[0], r[1]), pack_bf16x2(r[2], r[3]));
                        } else if constexpr (S == 4 && sizeof(TOUT) == 4) {
                            *reinterpret_cast<float4 *>(o) = make_float4(r[0], r[1], r[2], r[3]);
                        } else if constexpr (S == 2 && sizeof(TOUT) == 2) {
                            *reinterpret_cast<uint32_t *>(o) = pack_bf16x2(r[0], r[1]);
                        } else if constexpr (S == 2 && sizeof(TOUT) == 4) {
                            *reinterpret_cast<float2 *>(o) = make_float2(r[0], r[1]);
                        } else {
#pragma unroll
                            for (int j = 0; j < S; ++j) o[j] = from_f32<TOUT>(r[j]);
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
