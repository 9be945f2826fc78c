// wdsr_bf16.cuh -- bf16 tensor-core kernels of the WDSR-B forward (fused residual block, fused tail).
//
// Arithmetic contract (the "bf16 path" of the north star, gated at PSNR >= 50 dB against the fp32 reference):
//   operands (activations, weights) bf16, rounded ONCE (weights at load time after the fp32 weight-norm fold);
//   every accumulation, bias add, ReLU and the residual add in fp32 registers; the trunk is stored bf16 NHWC.
//
// Fused residual block  out = x + conv3x3( W2 * relu(W1 * x + b1) + b2 )          models/basic_wdsr_b.py:96-144
//   One CTA owns a TW x TH spatial tile.  The trunk tile + 1-pixel halo is staged once in shared memory, the
//   144-channel expand output never exists: it is produced 16 channels at a time in MMA accumulators, ReLU'd,
//   re-packed in registers as the A operand of the reduce GEMM (accumulator layout == A-fragment layout), and the
//   20(->24)-channel reduce output t2 is the only intermediate written to shared memory.  The 3x3 then runs as nine
//   shifted implicit-GEMM taps over t2 (ldmatrix takes per-row addresses, so a tap shift is free), adds bias and the
//   residual from the staged trunk and writes one trunk tensor per block.
//   t2 of out-of-image halo pixels is forced to zero: the reference zero-pads t2, not the trunk.
//
// Pruned widths (IN, M1, M2) from the NAS search are handled by padded specialisations: IN and M2 are padded to
// multiples of 8 (template CP, M2P in {8,16,24}), M1 to a multiple of 16 (runtime chunk count); padded filter
// rows/cols are zero so the padding is exact.
#pragma once
#include "common.cuh"

namespace b200sr {

// byte layout of one block's weight image (built on the host, copied verbatim to shared memory)
struct BlockBf16Layout {
    int s1, s2, s3;                 // row strides (elements) of W1s, W2s, W3s
    int w1, w2, w3, b1, b2, b3, total;  // byte offsets
    __host__ __device__ BlockBf16Layout(int CP, int M1P, int M2P) {
        s1 = ldm_stride(CP);
        s2 = ldm_stride(M1P);
        s3 = ldm_stride(M2P);
        w1 = 0;                          // [M1P][s1]   rows = expand channel, cols = trunk channel
        w2 = w1 + M1P * s1 * 2;          // [M2P][s2]   rows = reduce channel, cols = expand channel
        w3 = w2 + M2P * s2 * 2;          // [9][CP][s3] rows = out channel,    cols = reduce channel
        b1 = w3 + 9 * CP * s3 * 2;       // f32 [M1P]
        b2 = b1 + M1P * 4;               // f32 [M2P]
        b3 = b2 + M2P * 4;               // f32 [CP]
        total = round_up(b3 + CP * 4, 16);
    }
};

// A fragment(s) of a 16-row x K operand whose rows are given per lane (K = 8, 16 or 24).
// a[0..3] cover k 0..15 (m16n8k16 layout), a[4..5] cover k 16..23 (m16n8k8 layout); for K == 8 only a[0..1] (k8).
template <int K>
__device__ __forceinline__ void load_a_frags(uint32_t (&a)[6], uint32_t row_addr_k0, int lane) {
    // row_addr_k0: shared address of element [row = lane % 16][k = 0] for this lane
    if constexpr (K == 8) {
        ldmatrix_x2(a[0], a[1], row_addr_k0);
    } else {
        ldmatrix_x4(a[0], a[1], a[2], a[3], row_addr_k0 + (lane >> 4) * 16);
        if constexpr (K == 24) ldmatrix_x2(a[4], a[5], row_addr_k0 + 32);
    }
}

// B fragments for NT n-tiles (8 output columns each) x K, from a [n][stride] bf16 array at `base` (shared address of
// row n0, k0).  b[nt][0..1] = k16 fragment, b[nt][2] = k8 remainder (K==24) ; for K == 8 only b[nt][0].
template <int K, int NT>
__device__ __forceinline__ void load_b_frags(uint32_t (&b)[NT][3], uint32_t base, int stride_bytes, int lane) {
    const int r8 = lane & 7, id = lane >> 3;
    if constexpr (K == 8) {
#pragma unroll
        for (int nt = 0; nt + 3 < NT; nt += 4)
            ldmatrix_x4(b[nt][0], b[nt + 1][0], b[nt + 2][0], b[nt + 3][0], base + ((nt + id) * 8 + r8) * stride_bytes);
        if constexpr (NT % 4 == 3) {
            ldmatrix_x2(b[NT - 3][0], b[NT - 2][0], base + ((NT - 3 + (id & 1)) * 8 + r8) * stride_bytes);
            ldmatrix_x1(b[NT - 1][0], base + ((NT - 1) * 8 + r8) * stride_bytes);
        } else if constexpr (NT % 4 == 2) {
            ldmatrix_x2(b[NT - 2][0], b[NT - 1][0], base + ((NT - 2 + (id & 1)) * 8 + r8) * stride_bytes);
        } else if constexpr (NT % 4 == 1) {
            ldmatrix_x1(b[NT - 1][0], base + ((NT - 1) * 8 + r8) * stride_bytes);
        }
    } else {
        // k 0..15: two matrices per n-tile -> x4 covers two n-tiles
#pragma unroll
        for (int nt = 0; nt + 1 < NT; nt += 2)
            ldmatrix_x4(b[nt][0], b[nt][1], b[nt + 1][0], b[nt + 1][1],
                        base + ((nt + (id >> 1)) * 8 + r8) * stride_bytes + (id & 1) * 16);
        if constexpr (NT % 2 == 1)
            ldmatrix_x2(b[NT - 1][0], b[NT - 1][1], base + ((NT - 1) * 8 + r8) * stride_bytes + (id & 1) * 16);
        if constexpr (K == 24) {
#pragma unroll
            for (int nt = 0; nt + 3 < NT; nt += 4)
                ldmatrix_x4(b[nt][2], b[nt + 1][2], b[nt + 2][2], b[nt + 3][2],
                            base + ((nt + id) * 8 + r8) * stride_bytes + 32);
            if constexpr (NT % 4 == 3) {
                ldmatrix_x2(b[NT - 3][2], b[NT - 2][2], base + ((NT - 3 + (id & 1)) * 8 + r8) * stride_bytes + 32);
                ldmatrix_x1(b[NT - 1][2], base + ((NT - 1) * 8 + r8) * stride_bytes + 32);
            } else if constexpr (NT % 4 == 2) {
                ldmatrix_x2(b[NT - 2][2], b[NT - 1][2], base + ((NT - 2 + (id & 1)) * 8 + r8) * stride_bytes + 32);
            } else if constexpr (NT % 4 == 1) {
                ldmatrix_x1(b[NT - 1][2], base + ((NT - 1) * 8 + r8) * stride_bytes + 32);
            }
        }
    }
}

// acc(16 x 8*NT) += A(16 x K) * B(K x 8*NT)
template <int K, int NT>
__device__ __forceinline__ void mma_tile(float (&acc)[NT][4], const uint32_t (&a)[6], const uint32_t (&b)[NT][3]) {
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        if constexpr (K == 8) {
            mma_1688(acc[nt], a[0], a[1], b[nt][0]);
        } else {
            mma_16816(acc[nt], a[0], a[1], a[2], a[3], b[nt][0], b[nt][1]);
            if constexpr (K == 24) mma_1688(acc[nt], a[4], a[5], b[nt][2]);
        }
    }
}

template <int CP, int M2P, int TW, int TH, int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32, 2)
wdsr_block_bf16_kernel(const bf16 *__restrict__ in, bf16 *__restrict__ out, const uint8_t *__restrict__ wimg, int M1P, int N,
                       int H, int W, int tiles_x, int tiles_y, int ntiles) {
    constexpr int HW_ = TW + 2, HH_ = TH + 2, HP = HW_ * HH_;
    constexpr int XS = ldm_stride(CP), TS = ldm_stride(M2P);  // pixel strides (elements)
    constexpr int NT1 = CP / 8;                               // n-tiles of the trunk (3x3 output)
    constexpr int NT2 = M2P / 8;                              // n-tiles of the reduce output
    constexpr int MB = 2;                                     // m16 tiles per warp iteration in phase 1
    constexpr int R = 4;                                      // output rows per warp strip in phase 3
    static_assert(TW % 16 == 0 && TH % R == 0, "tile shape");
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const BlockBf16Layout L(CP, M1P, M2P);
    uint8_t *wsm = smem_raw;
    bf16 *xs = reinterpret_cast<bf16 *>(smem_raw + L.total);  // [HP][XS]
    bf16 *t2s = xs + HP * XS;                                 // [HP][TS]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    constexpr int NTHREADS = NWARPS * 32;

    // weights: one verbatim copy per CTA (persistent over tiles)
    for (int i = tid; i < L.total / 16; i += NTHREADS) cp_async16(wsm + i * 16, wimg + i * 16, 16);
    cp_async_commit();

    const uint32_t w1s = smem_u32(wsm + L.w1), w2s = smem_u32(wsm + L.w2), w3s = smem_u32(wsm + L.w3);
    const float *b1s = reinterpret_cast<const float *>(wsm + L.b1);
    const float *b2s = reinterpret_cast<const float *>(wsm + L.b2);
    const float *b3s = reinterpret_cast<const float *>(wsm + L.b3);
    const uint32_t xs_u = smem_u32(xs), t2s_u = smem_u32(t2s);

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, n = tile / (tiles_x * tiles_y);
        const int x0 = tx * TW - 1, y0 = ty * TH - 1;  // image coords of halo pixel (0,0)

        // ---- stage trunk tile + halo (zero-filled outside the image)
        for (int i = tid; i < HP * NT1; i += NTHREADS) {
            const int hp = i / NT1, q = i % NT1;
            const int gy = y0 + hp / HW_, gx = x0 + hp % HW_;
            const bool ok = gy >= 0 && gy < H && gx >= 0 && gx < W;
            const bf16 *src = ok ? in + (((long long)n * H + gy) * W + gx) * CP + q * 8 : in;
            cp_async16(xs + hp * XS + q * 8, src, ok ? 16 : 0);
        }
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();

        // ---- phase 1: t2 = W2 * relu(W1 * x + b1) + b2 for all halo pixels, MB m16-tiles per warp iteration
        constexpr int NMT = ceil_div(HP, 16);
        for (int mt0 = warp * MB; mt0 < NMT; mt0 += NWARPS * MB) {
            uint32_t xa[MB][6];
#pragma unroll
            for (int mb = 0; mb < MB; ++mb) {
                int p = (mt0 + mb) * 16 + (lane & 15);
                p = p < HP ? p : HP - 1;
                load_a_frags<CP>(xa[mb], xs_u + p * (XS * 2), lane);
            }
            float acc2[MB][NT2][4];
#pragma unroll
            for (int mb = 0; mb < MB; ++mb)
#pragma unroll
                for (int nt = 0; nt < NT2; ++nt) {
                    const float2 bv = *reinterpret_cast<const float2 *>(b2s + nt * 8 + 2 * t);
                    acc2[mb][nt][0] = bv.x, acc2[mb][nt][1] = bv.y, acc2[mb][nt][2] = bv.x, acc2[mb][nt][3] = bv.y;
                }
            for (int ch = 0; ch < M1P; ch += 16) {
                uint32_t bw1[2][3];
                load_b_frags<CP, 2>(bw1, w1s + ch * (L.s1 * 2), L.s1 * 2, lane);
                uint32_t bw2[NT2][3];
                load_b_frags<16, NT2>(bw2, w2s + ch * 2, L.s2 * 2, lane);
                const float2 bia0 = *reinterpret_cast<const float2 *>(b1s + ch + 2 * t);
                const float2 bia1 = *reinterpret_cast<const float2 *>(b1s + ch + 8 + 2 * t);
#pragma unroll
                for (int mb = 0; mb < MB; ++mb) {
                    float acc1[2][4] = {{bia0.x, bia0.y, bia0.x, bia0.y}, {bia1.x, bia1.y, bia1.x, bia1.y}};
                    mma_tile<CP, 2>(acc1, xa[mb], bw1);
                    uint32_t a2[6];
                    a2[0] = pack_bf16x2(fmaxf(acc1[0][0], 0.f), fmaxf(acc1[0][1], 0.f));
                    a2[1] = pack_bf16x2(fmaxf(acc1[0][2], 0.f), fmaxf(acc1[0][3], 0.f));
                    a2[2] = pack_bf16x2(fmaxf(acc1[1][0], 0.f), fmaxf(acc1[1][1], 0.f));
                    a2[3] = pack_bf16x2(fmaxf(acc1[1][2], 0.f), fmaxf(acc1[1][3], 0.f));
                    a2[4] = a2[5] = 0u;
                    mma_tile<16, NT2>(acc2[mb], a2, bw2);
                }
            }
            // t2 -> shared (bf16), zero where the halo pixel lies outside the image
#pragma unroll
            for (int mb = 0; mb < MB; ++mb)
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int p = (mt0 + mb) * 16 + g + half * 8;
                    if (p < HP) {
                        const int gy = y0 + p / HW_, gx = x0 + p % HW_;
                        const bool ok = gy >= 0 && gy < H && gx >= 0 && gx < W;
#pragma unroll
                        for (int nt = 0; nt < NT2; ++nt) {
                            const uint32_t v = ok ? pack_bf16x2(acc2[mb][nt][half * 2], acc2[mb][nt][half * 2 + 1]) : 0u;
                            *reinterpret_cast<uint32_t *>(t2s + p * TS + nt * 8 + 2 * t) = v;
                        }
                    }
                }
        }
        __syncthreads();

        // ---- phase 3: out = x + b3 + conv3x3(t2); warp strip = 16 px wide x R rows
        constexpr int NSTRIP = (TW / 16) * (TH / R);
        for (int strip = warp; strip < NSTRIP; strip += NWARPS) {
            const int xt = strip % (TW / 16), ys = (strip / (TW / 16)) * R;  // interior coords of the strip origin
            float acc[R][NT1][4];
#pragma unroll
            for (int j = 0; j < R; ++j)
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int hp = (ys + j + 1) * HW_ + xt * 16 + 1 + g + half * 8;
#pragma unroll
                    for (int nt = 0; nt < NT1; ++nt) {
                        const float2 bv = *reinterpret_cast<const float2 *>(b3s + nt * 8 + 2 * t);
                        const float2 rv = unpack_bf16x2(*reinterpret_cast<const uint32_t *>(xs + hp * XS + nt * 8 + 2 * t));
                        acc[j][nt][half * 2] = bv.x + rv.x;
                        acc[j][nt][half * 2 + 1] = bv.y + rv.y;
                    }
                }
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                uint32_t bw[3][NT1][3];  // the three taps (dy = 0..2) of this dx
#pragma unroll
                for (int dy = 0; dy < 3; ++dy)
                    load_b_frags<M2P, NT1>(bw[dy], w3s + (dy * 3 + dx) * (CP * L.s3 * 2), L.s3 * 2, lane);
#pragma unroll
                for (int r = 0; r < R + 2; ++r) {  // input (halo) row ys + r feeds output rows r - dy
                    uint32_t a[6];
                    const int hp = (ys + r) * HW_ + xt * 16 + dx + (lane & 15);
                    load_a_frags<M2P>(a, t2s_u + hp * (TS * 2), lane);
#pragma unroll
                    for (int dy = 0; dy < 3; ++dy) {
                        const int j = r - dy;
                        if (j >= 0 && j < R) mma_tile<M2P, NT1>(acc[j], a, bw[dy]);
                    }
                }
            }
            // results -> xs in place (each thread overwrites exactly the residual elements it read)
#pragma unroll
            for (int j = 0; j < R; ++j)
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int hp = (ys + j + 1) * HW_ + xt * 16 + 1 + g + half * 8;
#pragma unroll
                    for (int nt = 0; nt < NT1; ++nt)
                        *reinterpret_cast<uint32_t *>(xs + hp * XS + nt * 8 + 2 * t) =
                            pack_bf16x2(acc[j][nt][half * 2], acc[j][nt][half * 2 + 1]);
                }
        }
        __syncthreads();

        // ---- coalesced 16-byte stores of the interior
        for (int i = tid; i < TW * TH * NT1; i += NTHREADS) {
            const int q = i % NT1, lp = i / NT1, lx = lp % TW, ly = lp / TW;
            const int gy = y0 + 1 + ly, gx = x0 + 1 + lx;
            if (gy < H && gx < W) {
                const uint4 v = *reinterpret_cast<const uint4 *>(xs + ((ly + 1) * HW_ + lx + 1) * XS + q * 8);
                *reinterpret_cast<uint4 *>(out + (((long long)n * H + gy) * W + gx) * CP + q * 8) = v;
            }
        }
        __syncthreads();
    }
    cp_async_wait<0>();
}

template <int CP, int M2P, int TW, int TH>
inline size_t wdsr_block_bf16_smem(int M1P) {
    BlockBf16Layout L(CP, M1P, M2P);
    return (size_t)L.total + (size_t)(TW + 2) * (TH + 2) * (ldm_stride(CP) + ldm_stride(M2P)) * 2;
}

// ------------------------------------------------------------------------------------------------------------------
// fused tail, bf16 tensor cores:
//   y = PixelShuffle_s( conv3x3(trunk, Wt) + conv5x5(x - mean, Ws) + (bt + bs) ) + out_add      basic_wdsr_b.py:90-92
//   tail: implicit GEMM, 9 taps x CP.   skip: x-mean is staged as NHWC4 bf16 (channel 3 = 0) so that the 5 taps of one
//   filter row are 20 contiguous values: K per filter row = 24 (6 pixels x 4; the 6th pixel and 4th channel carry
//   zero weights), 5 filter rows.  The PixelShuffle and the mean add happen in the store epilogue: for even s each
//   lane owns two horizontally adjacent output pixels, and 8 lanes x 2 = one full 32/64-byte segment of an output row.
// weight image: Wt[9][NOP][st] bf16 | Ws[NOP][120] bf16 | bias[NOP] f32     (NOP = 3*s*s padded to 8)
// ------------------------------------------------------------------------------------------------------------------
struct TailBf16Layout {
    int st, wt, ws, bias, total;
    __host__ __device__ TailBf16Layout(int CP, int NOP) {
        st = ldm_stride(CP);
        wt = 0;
        ws = wt + 9 * NOP * st * 2;
        bias = ws + NOP * 120 * 2;
        total = round_up(bias + NOP * 4, 16);
    }
};

template <typename TIN, typename TOUT, int CP, int S, int TW, int TH, int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32, 2)
wdsr_tail_bf16_kernel(const bf16 *__restrict__ trunk, const TIN *__restrict__ x, TOUT *__restrict__ y,
                      const uint8_t *__restrict__ wimg, int N, int H, int W, int tiles_x, int tiles_y, int ntiles, float mean,
                      float out_add) {
    constexpr int NO = 3 * S * S, NOP = round_up(NO, 8), NT = NOP / 8;
    constexpr int HW1 = TW + 2, HP1 = HW1 * (TH + 2);
    constexpr int XW = TW + 6, XH = TH + 4;  // x tile: 2 left, 3(+1 pad) right, 2 up/down ; 4 bf16 per pixel
    constexpr int XS = ldm_stride(CP);
    constexpr int R = 2;
    constexpr int NT1 = CP / 8;
    constexpr int NTHREADS = NWARPS * 32;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const TailBf16Layout L(CP, NOP);
    uint8_t *wsm = smem_raw;
    bf16 *ts = reinterpret_cast<bf16 *>(smem_raw + L.total);  // [HP1][XS]
    bf16 *x4s = ts + HP1 * XS;                                // [XH][XW][4]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;

    for (int i = tid; i < L.total / 16; i += NTHREADS) cp_async16(wsm + i * 16, wimg + i * 16, 16);
    cp_async_commit();
    const uint32_t wts = smem_u32(wsm + L.wt), wss = smem_u32(wsm + L.ws);
    const float *bias = reinterpret_cast<const float *>(wsm + L.bias);
    const uint32_t ts_u = smem_u32(ts);
    const int OH = S * H, OW = S * W;

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, n = tile / (tiles_x * tiles_y);
        const int x0 = tx * TW, y0 = ty * TH;
        for (int i = tid; i < HP1 * NT1; i += NTHREADS) {
            const int hp = i / NT1, q = i % NT1;
            const int gy = y0 - 1 + hp / HW1, gx = x0 - 1 + hp % HW1;
            const bool ok = gy >= 0 && gy < H && gx >= 0 && gx < W;
            const bf16 *src = ok ? trunk + (((long long)n * H + gy) * W + gx) * CP + q * 8 : trunk;
            cp_async16(ts + hp * XS + q * 8, src, ok ? 16 : 0);
        }
        cp_async_commit();
        for (int i = tid; i < XH * XW; i += NTHREADS) {
            const int gy = y0 - 2 + i / XW, gx = x0 - 2 + i % XW;
            float v0 = 0.f, v1 = 0.f, v2 = 0.f;
            if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
                const long long o = (((long long)n * 3) * H + gy) * W + gx;
                v0 = to_f32<TIN>(x[o]) - mean;
                v1 = to_f32<TIN>(x[o + (long long)H * W]) - mean;
                v2 = to_f32<TIN>(x[o + 2ll * H * W]) - mean;
            }
            uint2 pk;
            pk.x = pack_bf16x2(v0, v1);
            pk.y = pack_bf16x2(v2, 0.f);
            *reinterpret_cast<uint2 *>(x4s + i * 4) = pk;
        }
        cp_async_wait<0>();
        __syncthreads();

        constexpr int NSTRIP = (TW / 16) * (TH / R);
        for (int strip = warp; strip < NSTRIP; strip += NWARPS) {
            const int xt = strip % (TW / 16), ys = (strip / (TW / 16)) * R;
            float acc[R][NT][4];
#pragma unroll
            for (int j = 0; j < R; ++j)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    const float2 bv = *reinterpret_cast<const float2 *>(bias + nt * 8 + 2 * t);
                    acc[j][nt][0] = bv.x, acc[j][nt][1] = bv.y, acc[j][nt][2] = bv.x, acc[j][nt][3] = bv.y;
                }
            // tail 3x3 over the trunk
#pragma unroll 1
            for (int tap = 0; tap < 9; ++tap) {
                uint32_t bw[NT][3];
                load_b_frags<CP, NT>(bw, wts + tap * (NOP * L.st * 2), L.st * 2, lane);
#pragma unroll
                for (int j = 0; j < R; ++j) {
                    uint32_t a[6];
                    const int hp = (ys + j + tap / 3) * HW1 + xt * 16 + tap % 3 + (lane & 15);
                    load_a_frags<CP>(a, ts_u + hp * (XS * 2), lane);
                    mma_tile<CP, NT>(acc[j], a, bw);
                }
            }
            // skip 5x5 over x - mean: one K=24 step per filter row
#pragma unroll 1
            for (int ky = 0; ky < 5; ++ky) {
                uint32_t bw[NT][3];
                load_b_frags<24, NT>(bw, wss + ky * 48, 240, lane);
#pragma unroll
                for (int j = 0; j < R; ++j) {
                    // A[row = pixel, k = (dx*4 + c)], dx = 0..5 starting 2 pixels left of the output pixel
                    const bf16 *rowp = x4s + ((ys + j + ky) * XW + xt * 16) * 4;
                    uint32_t a[6];
                    a[0] = *reinterpret_cast<const uint32_t *>(rowp + g * 4 + 2 * t);
                    a[1] = *reinterpret_cast<const uint32_t *>(rowp + (g + 8) * 4 + 2 * t);
                    a[2] = *reinterpret_cast<const uint32_t *>(rowp + g * 4 + 8 + 2 * t);
                    a[3] = *reinterpret_cast<const uint32_t *>(rowp + (g + 8) * 4 + 8 + 2 * t);
                    a[4] = *reinterpret_cast<const uint32_t *>(rowp + g * 4 + 16 + 2 * t);
                    a[5] = *reinterpret_cast<const uint32_t *>(rowp + (g + 8) * 4 + 16 + 2 * t);
                    mma_tile<24, NT>(acc[j], a, bw);
                }
            }
            // PixelShuffle store epilogue
#pragma unroll
            for (int j = 0; j < R; ++j) {
                const int gy = y0 + ys + j;
                if (gy >= H) continue;
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    const int ch = nt * 8 + 2 * t;
                    if (ch >= NO) continue;
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        const int gx = x0 + xt * 16 + g + half * 8;
                        if (gx >= W) continue;
                        const float v0 = acc[j][nt][half * 2] + out_add, v1 = acc[j][nt][half * 2 + 1] + out_add;
                        const int c = ch / (S * S), rem = ch % (S * S), i = rem / S, jj = rem % S;
                        TOUT *o = y + (((long long)n * 3 + c) * OH + (S * gy + i)) * OW + S * gx + jj;
                        if constexpr (S % 2 == 0) {
                            if constexpr (sizeof(TOUT) == 4) {
                                *reinterpret_cast<float2 *>(o) = make_float2(v0, v1);
                            } else {
                                *reinterpret_cast<uint32_t *>(o) = pack_bf16x2(v0, v1);
                            }
                        } else {
                            o[0] = from_f32<TOUT>(v0);
                            const int ch1 = ch + 1;
                            if (ch1 < NO) {
                                const int c1 = ch1 / (S * S), rem1 = ch1 % (S * S);
                                y[(((long long)n * 3 + c1) * OH + (S * gy + rem1 / S)) * OW + S * gx + rem1 % S] =
                                    from_f32<TOUT>(v1);
                            }
                        }
                    }
                }
            }
        }
        __syncthreads();
    }
    cp_async_wait<0>();
}

template <int CP, int S, int TW, int TH>
inline size_t wdsr_tail_bf16_smem() {
    constexpr int NOP = round_up(3 * S * S, 8);
    TailBf16Layout L(CP, NOP);
    return (size_t)L.total + (size_t)(TW + 2) * (TH + 2) * ldm_stride(CP) * 2 + (size_t)(TW + 6) * (TH + 4) * 8;
}

}  // namespace b200sr
