// split_block_tc.cu -- bf16 arm of the fork's searchable block body, Split_Block.forward_body (models/wdsr_b.py:482-496), with the
// pointwise half on the tensor cores.  Same algebra as split_block.cu (which stays the fp32 / 1e-4 arm and the fallback for widths
// that are not a multiple of 8):
//
//   x1 = e * (g * x),  x2 = g * x - x1,   y = x2 + e * (x2 + sum_k p_k * relu(PW_k(relu(DW_k(x1)))) + x1)        k = 3, 5, 7
//
// What changed against the FFMA kernel (82 us per 360p frame, 7.8 k instructions per pixel for 3.7 k FMAs):
//   * the three depthwise filters run in ONE pass over the staged tile: a work item (channel, tile row, 8-pixel segment) loads its
//     7 x 16 window once (14 LDS.128 of bf16, unpacked in registers) and feeds all 83 taps of the 3x3 / 5x5 / 7x7 filters (664 FMAs) -- the separate passes loaded
//     the window three times with scalar loads;
//   * relu(DW_k) is stored as bf16 [k][channel][pixel] and the 1x1 convolutions are mma.sync m16n8k16 / m16n8k8 (bf16 in, fp32
//     accumulate): the A fragments come straight out of the channel-major image with ldmatrix.trans, the B fragments (the folded
//     1x1 filters, bf16) are packed per lane on the host;
//   * the folded parameters live in a device-memory image (SplitTcImage) instead of travelling as an 18 KB kernel argument: the
//     per-lane B fragments and biases would be DIVERGENT constant-bank reads (one serialised access per distinct address: the first
//     form of this kernel, with everything in the constant bank, ran at 128 us); the filter taps are broadcast LDS.128 reads of a
//     shared-memory copy.  (The same tap copy in the FFMA kernel made it slower, 94 -> 101 us: there the third resident CTA per SM
//     that the extra shared memory costs matters more than the warp-uniform LDC per tap.);
//   * persistent CTAs: the raw bf16 halo tile of the next work item streams in with cp.async while this one is computed (e * g
//     is folded into the taps), and the result leaves through shared memory as 16-byte stores.
// Tensors are the reference's NCHW in bf16.  W must be a multiple of 8 (16-byte rows); b200sr.cu routes everything else to
// split_block.cu.
#include <cstring>
#include <vector>

#include "common.cuh"
#include "launch.h"

namespace b200sr {

namespace splittc {
constexpr int TW = 32, TH = 8, HALO = 3, SH = TH + 2 * HALO, SWP = 40, NTHREADS = 256;   // staged column j <-> gx = x0 - 4 + j
constexpr int XR_ROW = SWP * 2;     // bytes per staged row (bf16)
constexpr int TAPF = 112;           // floats per channel of the tap image: w7 rows padded to 8 (56) | w5 rows padded to 8 (40) | w3 rows padded to 4 (12) | b3 b5 b7 0
constexpr int DS_PITCH = 528;       // bytes per channel row of relu(DW_k): 256 px bf16 + 16 (ldmatrix rows land in distinct bank groups)
constexpr int ST_PITCH = 260;       // floats per channel row of the staged branch sum (conflict-free fragment writes)
__host__ __device__ constexpr size_t xr_bytes(int C) { return (size_t)C * SH * XR_ROW; }
__host__ __device__ constexpr size_t ds_bytes(int C) { return (size_t)3 * C * DS_PITCH; }
__host__ __device__ constexpr size_t tap_bytes(int C) { return (size_t)C * TAPF * 4; }
__host__ __device__ constexpr size_t smem_bytes(int C) { return 2 * xr_bytes(C) + ds_bytes(C) + tap_bytes(C) + (size_t)2 * C * 4; }
}  // namespace splittc

template <int C> struct SplitTcImage {       // device memory, every member a multiple of 16 bytes
    float taps[C * splittc::TAPF];            // depthwise taps PRE-MULTIPLIED by e * g of their channel (DW_k is linear in x1 = e * g * x), biases plain
    float eg[2 * C];                          // e[C] | g[C]
    uint32_t bfrag[3][C / 8][C / 8][32];      // [branch][n-tile][register][lane]
    float pwb[3 * C];
    float p[4];
};

// Persistent CTAs (two per SM for C <= 24): the raw bf16 halo tile of the NEXT work item streams into the other staging buffer with
// cp.async (8-byte pieces, zero-filled outside the image = the convolutions' padding) while this one is computed; the first form staged
// with plain loads and spent 46 % of its stall samples waiting for them (profiles/r02_split_block_tc_ncu.md).
template <int C>
__global__ void __launch_bounds__(splittc::NTHREADS, (C <= 24 ? 2 : 1))
split_block_tc_kernel(const bf16 *__restrict__ x, bf16 *__restrict__ y, const SplitTcImage<C> *__restrict__ img, int N, int H, int W,
                      int tiles_x, int tiles_y, int ntiles) {
    using namespace splittc;
    constexpr int NT = C / 8;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint8_t *xr = smem_raw;                                        // [2][C][SH][SWP] bf16: g-less, e-less x, zero outside the image
    uint8_t *ds = smem_raw + 2 * xr_bytes(C);                      // [3][C][DS_PITCH]  relu(DW_k(x1)) as bf16
    float *taps = reinterpret_cast<float *>(ds + ds_bytes(C));     // [C][TAPF]
    float *eg = taps + C * TAPF;                                   // e[C] | g[C]
    float *st = reinterpret_cast<float *>(ds);                     // [C][ST_PITCH]  branch sum (aliases ds once every warp's MMAs have read it)
    static_assert((size_t)C * ST_PITCH * 4 <= ds_bytes(C), "st must fit into ds");
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long plane = (long long)H * W;

    auto origin = [&](int tile, int &x0, int &y0, int &n) {
        x0 = (tile % tiles_x) * TW, y0 = ((tile / tiles_x) % tiles_y) * TH, n = tile / (tiles_x * tiles_y);
    };
    // staging work split: a piece index (0..9) and a first (channel, row) pair per thread, 25 row-pairs apart per step -- (c, r) and the
    // addresses advance incrementally (the flat-index form spent ~50 instructions per piece on divisions)
    const int pf_h = tid % (SWP / 4), pf_rc0 = tid / (SWP / 4);
    constexpr int PF_STEP = 25, PF_ITERS = (C * SH + PF_STEP - 1) / PF_STEP;
    auto prefetch = [&](int tile, int buf) {   // item = (channel, halo row, 4-pixel piece); W % 4 == 0: a piece is entirely inside or outside
        if (tid >= PF_STEP * (SWP / 4)) return;
        int x0, y0, n;
        origin(tile, x0, y0, n);
        const int gx = x0 - 4 + 4 * pf_h;
        const bool inx = gx >= 0 && gx < W;
        int c = pf_rc0 / SH, r = pf_rc0 % SH;
        uint32_t dst = smem_u32(xr + buf * xr_bytes(C)) + (uint32_t)(pf_rc0 * XR_ROW + 8 * pf_h);
        const bf16 *src = x + ((long long)n * C + c) * plane + (long long)(y0 - HALO + r) * W + gx;
#pragma unroll 2
        for (int k = 0; k < PF_ITERS; ++k) {
            if (c < C) {
                const int gy = y0 - HALO + r;
                const bool in = inx && gy >= 0 && gy < H;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(dst), "l"(in ? src : x), "r"(in ? 8 : 0));
            }
            dst += PF_STEP * XR_ROW;
            r += PF_STEP - SH, c += 1, src += plane + (long long)(PF_STEP - SH) * W;      // 25 rows on = one channel + 11 rows
            if (r >= SH) r -= SH, c += 1, src += plane - (long long)SH * W;
        }
    };
    int tile = blockIdx.x;
    if (tile < ntiles) prefetch(tile, 0);
    cp_async_commit();
    for (int i = tid; i < C * TAPF / 4 + 2 * C / 4; i += NTHREADS)      // taps | e | g are contiguous in the image and in shared memory
        reinterpret_cast<float4 *>(taps)[i] = __ldg(reinterpret_cast<const float4 *>(img->taps) + i);

    for (int it = 0; tile < ntiles; ++it, tile += gridDim.x) {
        const int cur = it & 1;
        int x0, y0, n;
        origin(tile, x0, y0, n);
        cp_async_wait<0>();
        __syncthreads();     // this tile's staging buffer is complete; the previous tile's final stage is over (other buffer and ds are free)
        if (tile + (int)gridDim.x < ntiles) prefetch(tile + gridDim.x, cur ^ 1);
        cp_async_commit();
        const uint8_t *xc = xr + cur * xr_bytes(C);

        // ---- depthwise 3x3 + 5x5 + 7x7 in one pass: a warp takes the 32 items of one channel (filter taps are warp-uniform)
        {
            const int r = lane >> 2, g4 = lane & 3;
#pragma unroll 1
            for (int c = warp; c < C; c += NTHREADS / 32) {
                float a3[8], a5[8], a7[8];
                const float4 *tp = reinterpret_cast<const float4 *>(taps + c * TAPF);   // warp-uniform addresses: broadcast reads
                {
                    const float4 b = tp[27];
#pragma unroll
                    for (int j = 0; j < 8; ++j) a3[j] = b.x, a5[j] = b.y, a7[j] = b.z;
                }
                const uint8_t *xp = xc + (c * SH + r) * XR_ROW + 16 * g4;
#pragma unroll
                for (int ky = 0; ky < 7; ++ky) {
                    float v[16];
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const uint4 t = *reinterpret_cast<const uint4 *>(xp + ky * XR_ROW + 16 * q);
                        const uint32_t tw[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            v[8 * q + 2 * j] = __uint_as_float(tw[j] << 16);
                            v[8 * q + 2 * j + 1] = __uint_as_float(tw[j] & 0xffff0000u);
                        }
                    }
                    float w7[8], w5[8], w3[4];
                    {
                        const float4 p0 = tp[2 * ky], p1 = tp[2 * ky + 1];
                        w7[0] = p0.x, w7[1] = p0.y, w7[2] = p0.z, w7[3] = p0.w, w7[4] = p1.x, w7[5] = p1.y, w7[6] = p1.z, w7[7] = p1.w;
                        if (ky >= 1 && ky <= 5) {
                            const float4 q0 = tp[14 + 2 * (ky - 1)], q1 = tp[15 + 2 * (ky - 1)];
                            w5[0] = q0.x, w5[1] = q0.y, w5[2] = q0.z, w5[3] = q0.w, w5[4] = q1.x, w5[5] = q1.y, w5[6] = q1.z, w5[7] = q1.w;
                        }
                        if (ky >= 2 && ky <= 4) {
                            const float4 q = tp[24 + (ky - 2)];
                            w3[0] = q.x, w3[1] = q.y, w3[2] = q.z, w3[3] = q.w;
                        }
                    }
#pragma unroll
                    for (int kx = 0; kx < 7; ++kx) {           // dx = kx - 3: window column of output pixel j is j + 4 + dx = j + 1 + kx
                        const float t7 = w7[kx];
#pragma unroll
                        for (int j = 0; j < 8; ++j) a7[j] = fmaf(v[j + 1 + kx], t7, a7[j]);
                        if (ky >= 1 && ky <= 5 && kx >= 1 && kx <= 5) {
                            const float t5 = w5[kx - 1];
#pragma unroll
                            for (int j = 0; j < 8; ++j) a5[j] = fmaf(v[j + 1 + kx], t5, a5[j]);
                        }
                        if (ky >= 2 && ky <= 4 && kx >= 2 && kx <= 4) {
                            const float t3 = w3[kx - 2];
#pragma unroll
                            for (int j = 0; j < 8; ++j) a3[j] = fmaf(v[j + 1 + kx], t3, a3[j]);
                        }
                    }
                }
                uint8_t *dp = ds + c * DS_PITCH + (r * TW + 8 * g4) * 2;
                auto rp = [](float lo, float hi) {   // relu + round + pack in one instruction
                    uint32_t d;
                    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
                    return d;
                };
                auto put = [&](uint8_t *d, const float(&a)[8]) {
                    *reinterpret_cast<uint4 *>(d) = make_uint4(rp(a[0], a[1]), rp(a[2], a[3]), rp(a[4], a[5]), rp(a[6], a[7]));
                };
                put(dp, a3);
                put(dp + C * DS_PITCH, a5);
                put(dp + 2 * C * DS_PITCH, a7);
            }
        }
        __syncthreads();     // ds complete

        // ---- pointwise 1x1 of the three branches on mma.sync: warp w owns tile row w (32 pixels = two 16-row M-tiles)
        const int g = lane >> 2, t = lane & 3;
        float s[2][NT][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int q = 0; q < 4; ++q) s[mt][nt][q] = 0.f;
        {
            // ldmatrix.trans row address of this lane: stored row = channel k0 + 8 * (lane bit 4) + (lane & 7), 8 pixels from m0 + 8 * (lane bit 3)
            const uint32_t ds_u = smem_u32(ds) + (uint32_t)((((lane >> 4) & 1) * 8 + (lane & 7)) * DS_PITCH + (warp * TW + ((lane >> 3) & 1) * 8) * 2);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                uint32_t bf[NT][NT];
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                    for (int q = 0; q < NT; ++q) bf[nt][q] = __ldg(&img->bfrag[k][nt][q][lane]);
                const float pk = __ldg(&img->p[k]);
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    float acc[NT][4];
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) {
                        const float2 bb = __ldg(reinterpret_cast<const float2 *>(&img->pwb[k * C + 8 * nt + 2 * t]));
                        acc[nt][0] = bb.x, acc[nt][1] = bb.y, acc[nt][2] = bb.x, acc[nt][3] = bb.y;
                    }
                    const uint32_t ab = ds_u + (uint32_t)(k * C * DS_PITCH + mt * 32);
#pragma unroll
                    for (int ks = 0; ks < C / 16; ++ks) {       // k16 steps
                        uint32_t a0, a1, a2, a3;
                        asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                                     : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3) : "r"(ab + ks * 16 * DS_PITCH));
#pragma unroll
                        for (int nt = 0; nt < NT; ++nt) mma_16816(acc[nt], a0, a1, a2, a3, bf[nt][2 * ks], bf[nt][2 * ks + 1]);
                    }
                    if (C % 16 == 8) {                           // k8 tail (channels C-8 .. C-1): lanes 0..15 supply the addresses
                        uint32_t a0, a1;
                        const uint32_t at = smem_u32(ds) + (uint32_t)((C - 8 + (lane & 7)) * DS_PITCH + (warp * TW + ((lane >> 3) & 1) * 8) * 2 +
                                                                      k * C * DS_PITCH + mt * 32);
                        asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];\n" : "=r"(a0), "=r"(a1) : "r"(at));
#pragma unroll
                        for (int nt = 0; nt < NT; ++nt) mma_1688(acc[nt], a0, a1, bf[nt][NT - 1]);
                    }
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                        for (int q = 0; q < 4; ++q) s[mt][nt][q] = fmaf(fmaxf(acc[nt][q], 0.f), pk, s[mt][nt][q]);   // + x_ * pro[i]
                }
            }
        }
        __syncthreads();     // every warp has read its columns of ds: st may overwrite it
        // branch sum -> st[channel][pixel] (C fragment: rows g / g + 8, columns 2t / 2t + 1)
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                float *sp = st + (8 * nt + 2 * t) * ST_PITCH + warp * TW + 16 * mt + g;
                sp[0] = s[mt][nt][0];
                sp[ST_PITCH] = s[mt][nt][1];
                sp[8] = s[mt][nt][2];
                sp[ST_PITCH + 8] = s[mt][nt][3];
            }
        __syncthreads();

        // ---- y = x2 + e * (x2 + sum + x1): item = (channel, tile row, 8-pixel segment); x comes back from the staged tile, 16-byte stores
        for (int i = tid; i < C * TH * (TW / 8); i += NTHREADS) {
            const int sg = i % (TW / 8), r = (i / (TW / 8)) % TH, c = i / ((TW / 8) * TH);
            const int gy = y0 + r, gx = x0 + 8 * sg;
            if (gy >= H || gx >= W) continue;
            const uint8_t *xq = xc + (c * SH + r + HALO) * XR_ROW + (4 + 8 * sg) * 2;
            const uint2 ra = *reinterpret_cast<const uint2 *>(xq), rb = *reinterpret_cast<const uint2 *>(xq + 8);
            const uint32_t rw[4] = {ra.x, ra.y, rb.x, rb.y};
            const float *sp = st + c * ST_PITCH + r * TW + 8 * sg;
            const float4 sa = *reinterpret_cast<const float4 *>(sp), sb = *reinterpret_cast<const float4 *>(sp + 4);
            const float sv[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
            const float gm = eg[C + c], em = eg[c];
            uint4 out;
            uint32_t *ow = reinterpret_cast<uint32_t *>(&out);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 xv = unpack_bf16x2(rw[j]);
                const float xa = xv.x * gm, xb = xv.y * gm;
                const float x1a = xa * em, x1b = xb * em, x2a = xa - x1a, x2b = xb - x1b;
                ow[j] = pack_bf16x2(x2a + ((x2a + sv[2 * j]) + x1a) * em, x2b + ((x2b + sv[2 * j + 1]) + x1b) * em);
            }
            *reinterpret_cast<uint4 *>(y + (long long)n * C * plane + c * plane + (long long)gy * W + gx) = out;
        }
    }
    cp_async_wait<0>();
}

// ---- host side -------------------------------------------------------------------------------------------------------------------
static inline uint16_t bf16_rne(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);   // NaN
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}

template <int C>
static void pack_t(const float *f, std::vector<uint8_t> &out) {
    // f: the packed fp32 image of split_block.cu:  dw3 | dw5 | dw7 | dw_bias[3][C] | pw[3][C in][C out] | pw_bias[3][C] | e[C] | p[4] | g[C]
    out.assign(sizeof(SplitTcImage<C>), 0);
    SplitTcImage<C> *P = reinterpret_cast<SplitTcImage<C> *>(out.data());
    const float *dw3 = f, *dw5 = dw3 + C * 9, *dw7 = dw5 + C * 25, *dwb = dw7 + C * 49, *pw = dwb + 3 * C, *pwb = pw + 3 * C * C, *e = pwb + 3 * C,
                *p = e + C, *g = p + 4;
    for (int c = 0; c < C; ++c) {
        float *t = P->taps + c * splittc::TAPF;
        const float eg_c = e[c] * g[c];     // the kernel stages raw x: DW_k(e * g * x) = sum (w * e * g) x + b
        for (int ky = 0; ky < 7; ++ky)
            for (int kx = 0; kx < 7; ++kx) t[8 * ky + kx] = dw7[c * 49 + ky * 7 + kx] * eg_c;
        for (int ky = 0; ky < 5; ++ky)
            for (int kx = 0; kx < 5; ++kx) t[56 + 8 * ky + kx] = dw5[c * 25 + ky * 5 + kx] * eg_c;
        for (int ky = 0; ky < 3; ++ky)
            for (int kx = 0; kx < 3; ++kx) t[96 + 4 * ky + kx] = dw3[c * 9 + ky * 3 + kx] * eg_c;
        t[108] = dwb[c], t[109] = dwb[C + c], t[110] = dwb[2 * C + c];
        P->eg[c] = e[c], P->eg[C + c] = g[c];
    }
    memcpy(P->pwb, pwb, sizeof P->pwb);
    memcpy(P->p, p, 3 * sizeof(float));
    auto w = [&](int k, int in, int o) { return (uint32_t)bf16_rne(pw[((size_t)k * C + in) * C + o]); };
    for (int k = 0; k < 3; ++k)
        for (int nt = 0; nt < C / 8; ++nt)
            for (int lane = 0; lane < 32; ++lane) {
                const int gq = lane >> 2, t = lane & 3, o = 8 * nt + gq;   // B[k][n]: n = g -> output channel, k pairs 2t, 2t + 1 (+ 8)
                int q = 0;
                for (int ks = 0; ks < C / 16; ++ks) {
                    const int kb = 16 * ks;
                    P->bfrag[k][nt][q++][lane] = w(k, kb + 2 * t, o) | (w(k, kb + 2 * t + 1, o) << 16);
                    P->bfrag[k][nt][q++][lane] = w(k, kb + 2 * t + 8, o) | (w(k, kb + 2 * t + 9, o) << 16);
                }
                if (C % 16 == 8) P->bfrag[k][nt][q++][lane] = w(k, C - 8 + 2 * t, o) | (w(k, C - 8 + 2 * t + 1, o) << 16);
            }
}

void split_tc_pack(int C, const float *packed, std::vector<uint8_t> &out) {
    if (C == 8) pack_t<8>(packed, out);
    else if (C == 16) pack_t<16>(packed, out);
    else if (C == 24) pack_t<24>(packed, out);
    else if (C == 32) pack_t<32>(packed, out);
    else out.clear();
}

bool split_tc_eligible(int dtype, const void *x, const void *y, int W) {
    return dtype == kBF16 && W % 8 == 0 && ((uintptr_t)x % 16) == 0 && ((uintptr_t)y % 16) == 0;
}

template <int C>
static cudaError_t split_tc_t(const void *x, void *y, const uint8_t *image /* device */, int N, int H, int W, cudaStream_t st) {
    using namespace splittc;
    auto kern = split_block_tc_kernel<C>;
    static thread_local bool set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes(C));
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) set[dev] = true;
    }
    const int tx = ceil_div(W, TW), ty = ceil_div(H, TH), ntiles = tx * ty * N;
    const int slots = sm_count() * (C <= 24 ? 2 : 1);
    const int grid = ntiles < slots ? ntiles : slots;   // persistent CTAs walk the tile list with stride gridDim.x
    kern<<<grid, NTHREADS, smem_bytes(C), st>>>((const bf16 *)x, (bf16 *)y, reinterpret_cast<const SplitTcImage<C> *>(image), N, H, W, tx, ty, ntiles);
    return cudaGetLastError();
}

cudaError_t launch_split_block_tc(int C, const void *x, void *y, const uint8_t *image, int N, int H, int W, cudaStream_t st) {
    if (C == 8) return split_tc_t<8>(x, y, image, N, H, W, st);
    if (C == 16) return split_tc_t<16>(x, y, image, N, H, W, st);
    if (C == 24) return split_tc_t<24>(x, y, image, N, H, W, st);
    if (C == 32) return split_tc_t<32>(x, y, image, N, H, W, st);
    return cudaErrorInvalidValue;
}

}  // namespace b200sr
