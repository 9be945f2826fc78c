// wdsr_tc5.cuh -- fused WDSR-B residual block on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM).
//
//   out = x + b3 + conv3x3( W2 * relu(W1 * x + b1) + b2 )                     models/basic_wdsr_b.py:96-144
//
// One CTA owns a 32 x 16 spatial tile (+1 px halo = 34 x 18 = 612 pixels).  Pixels are the M dimension of three
// chained UMMA GEMMs (M = 128 pixels per instruction):
//   G1  D1[128 x M1] = X[128 x 24] * W1^T + b1     2 instr (K = 16 + 16: chunks {c0,c1} and {c2,ONE}); the bias rides on a
//                                                  constant-one channel (b1 split in two bf16 terms), so no K padding
//   E1  relu, round to bf16, written back IN PLACE over D1 as the packed A operand of G2 (A-from-TMEM)
//   G2  D2[128 x 32] = relu(T1)[128 x M1] * W2^T   M1/16 instr, A from TMEM
//   E2  + b2, zero for out-of-image pixels (the reference zero-pads t2, not the trunk), bf16 -> three x-shifted copies
//       of t2 in shared memory (each copy stores exactly the 32 columns one horizontal tap needs, so a tap shift is a
//       constant address offset and M-tiles of the 3x3 are contiguous)
//   G3  D3[128 x 32] = sum over 9 taps x 3 chunks  14 instr: the 27 (tap,chunk) slices are paired two per K=16
//                                                  instruction through the free LBO stride (no zero padding of K)
//   E3  + b3 + residual (fp32), bf16, 48-byte NHWC pixel stores
// Operand layout in shared memory is the SWIZZLE_NONE K-major "interleaved" form: [8-pixel group][chunk][8 px][16 B].
//
// This file holds the sequential reference form of the kernel (one role set, CTA-wide barriers between stages); it is
// the correctness vehicle for the descriptors / TMEM protocol.  The pipelined warp-specialised form builds on it.
#pragma once
#include "common.cuh"
#include "tc5.cuh"
#include "wdsr_tc5_layout.cuh"

// Optional phase timers for tools/tc5_probe.cu (compiled out of the library).
#ifdef B200SR_TC5_PROF
__device__ unsigned long long g_tc5_prof[64];
#define TC5_T0() const long long prof_t0__ = clock64()
#define TC5_ADD(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_tc5_prof[i] += (unsigned long long)(clock64() - prof_t0__); } while (0)
#else
#define TC5_T0() do {} while (0)
#define TC5_ADD(i) do {} while (0)
#endif

namespace b200sr {

__global__ void __launch_bounds__(128, 1)
wdsr_block_tc5_seq_kernel(const bf16 *__restrict__ in, bf16 *__restrict__ out, const uint8_t *__restrict__ wimg, int M1P, int N,
                          int H, int W, int tiles_x, int tiles_y, int ntiles) {
    using namespace tc5cfg;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const BlockTc5Layout L(M1P);
    uint8_t *ctrl = smem_raw;                 // [0,8) mbarrier, [16,20) tmem base
    uint8_t *xs = smem_raw + CTRL_BYTES;      // XS_BYTES
    uint8_t *t2 = xs + XS_BYTES;              // T2_BYTES
    uint8_t *wsm = t2 + T2_BYTES;             // L.total
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t bar = smem_u32(ctrl);
    const uint32_t xs_u = smem_u32(xs), t2_u = smem_u32(t2), w_u = smem_u32(wsm);

    // ---- one-time setup
    if (tid == 0) {
        tc5::mbar_init(bar, 1);
        tc5::mbar_init_fence();
    }
    __syncwarp();
    if (warp == 0) tc5::tmem_alloc(smem_u32(ctrl + 16), TMEM_COLS);
    for (int i = tid; i < L.total / 16; i += 128) cp_async16(wsm + i * 16, wimg + i * 16, 16);
    cp_async_commit();
    for (int i = tid; i < XS_BYTES / 16; i += 128) {  // zero, except the constant-one chunk: channels 0,1 = 1.0
        const bool one = ((i * 16) % XS_GROUP) >= 384;
        *reinterpret_cast<uint4 *>(xs + i * 16) = make_uint4(one ? 0x3F803F80u : 0u, 0u, 0u, 0u);
    }
    for (int i = tid; i < 16; i += 128) *reinterpret_cast<uint4 *>(t2 + 3 * T2_COPY + i * 16) = make_uint4(0u, 0u, 0u, 0u);
    cp_async_wait<0>();
    tc5::fence_proxy_async();
    tc5::fence_before_sync();
    __syncthreads();
    tc5::fence_after_sync();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t *>(ctrl + 16);
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;  // this warp's TMEM lane quadrant
    const float *b2s = reinterpret_cast<const float *>(wsm + L.b2);
    const float *b3s = reinterpret_cast<const float *>(wsm + L.b3);
    const uint32_t idesc1 = tc5::idesc_bf16_f32(128, M1P), idesc32 = tc5::idesc_bf16_f32(128, 32);
    uint32_t phase = 0;

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, n = tile / (tiles_x * tiles_y);
        const int x0 = tx * TW - 1, y0 = ty * TH - 1;

        // ---- stage the trunk tile + halo in the interleaved operand layout
        { TC5_T0();
        for (int i = tid; i < HP * 3; i += 128) {
            const int p = i / 3, q = i - 3 * p;
            const int gy = y0 + p / HW_, gx = x0 + p % HW_;
            const bool ok = gy >= 0 && gy < H && gx >= 0 && gx < W;
            const bf16 *src = ok ? in + (((long long)n * H + gy) * W + gx) * 24 + q * 8 : in;
            cp_async16(xs + (p >> 3) * XS_GROUP + q * 128 + (p & 7) * 16, src, ok ? 16 : 0);
        }
        cp_async_commit();
        cp_async_wait<0>();
        tc5::fence_proxy_async();
        __syncthreads();
        TC5_ADD(0); }

        for (int m = 0; m < NMT; ++m) {
            { TC5_T0();
            if (tid == 0) {  // G1
                tc5::fence_after_sync();
                const uint32_t a = xs_u + m * 16 * XS_GROUP;
                tc5::mma_ss(tmem + D1_COL, tc5::smem_desc(a, 128, XS_GROUP), tc5::smem_desc(w_u + L.w1, 128, 512), idesc1, false);
                tc5::mma_ss(tmem + D1_COL, tc5::smem_desc(a + 256, 128, XS_GROUP), tc5::smem_desc(w_u + L.w1 + 256, 128, 512),
                            idesc1, true);
                tc5::commit(bar);
            }
            tc5::mbar_wait(bar, phase);
            phase ^= 1;
            tc5::fence_after_sync();
            TC5_ADD(1); }
            { TC5_T0();
            // E1: relu -> bf16, in place
            for (int k = 0; k < M1P / 16; ++k) {
                uint32_t v[16], pk[8];
                tc5::tmem_ld16(tmem + lane_base + D1_COL + 16 * k, v);
                tc5::tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    pk[j] = pack_bf16x2(fmaxf(__uint_as_float(v[2 * j]), 0.f), fmaxf(__uint_as_float(v[2 * j + 1]), 0.f));
                tc5::tmem_st8(tmem + lane_base + D1_COL + 8 * k, pk);
            }
            tc5::tmem_wait_st();
            tc5::fence_before_sync();
            __syncthreads();
            TC5_ADD(2); }
            { TC5_T0();
            if (tid == 0) {  // G2: A from TMEM
                tc5::fence_after_sync();
                for (int j = 0; j < M1P / 16; ++j)
                    tc5::mma_ts(tmem + D2_COL, tmem + D1_COL + 8 * j, tc5::smem_desc(w_u + L.w2 + j * 256, 128, L.sbo2), idesc32,
                                j > 0);
                tc5::commit(bar);
            }
            tc5::mbar_wait(bar, phase);
            phase ^= 1;
            tc5::fence_after_sync();
            TC5_ADD(3); }
            { TC5_T0();
            {   // E2
                uint32_t v[32];
                tc5::tmem_ld16(tmem + lane_base + D2_COL, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
                tc5::tmem_ld16(tmem + lane_base + D2_COL + 16, *reinterpret_cast<uint32_t(*)[16]>(&v[16]));
                tc5::tmem_wait_ld();
                const int p = m * 128 + tid;
                if (p < HP) {
                    const int r = p / HW_, hx = p - r * HW_;
                    const int gy = y0 + r, gx = x0 + hx;
                    const bool ok = gy >= 0 && gy < H && gx >= 0 && gx < W;
                    uint4 c[3];
                    uint32_t *cw = reinterpret_cast<uint32_t *>(c);
#pragma unroll
                    for (int j = 0; j < 12; ++j)
                        cw[j] = ok ? pack_bf16x2(__uint_as_float(v[2 * j]) + b2s[2 * j], __uint_as_float(v[2 * j + 1]) + b2s[2 * j + 1])
                                   : 0u;
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        const int xi = hx - d;
                        if (xi >= 0 && xi < TW) {
                            uint8_t *dst = t2 + d * T2_COPY + r * T2_ROW + (xi >> 3) * T2_GROUP + (xi & 7) * 16;
#pragma unroll
                            for (int q = 0; q < 3; ++q) *reinterpret_cast<uint4 *>(dst + q * 128) = c[q];
                        }
                    }
                }
            }
            tc5::fence_before_sync();
            __syncthreads();
            TC5_ADD(4); }
        }
        tc5::fence_proxy_async();
        __syncthreads();

        for (int m3 = 0; m3 < TH / 4; ++m3) {
            { TC5_T0();
            if (tid == 0) {  // G3
                tc5::fence_after_sync();
                const uint32_t abase = t2_u + m3 * 4 * T2_ROW;
#pragma unroll 1
                for (int i = 0; i < 14; ++i) {
                    const int q0 = 2 * i, q1 = 2 * i + 1;
                    const uint32_t a0 = (q0 / 9) * T2_COPY + ((q0 / 3) % 3) * T2_ROW + (q0 % 3) * 128;
                    const uint32_t a1 = q1 < 27 ? (q1 / 9) * T2_COPY + ((q1 / 3) % 3) * T2_ROW + (q1 % 3) * 128 : a0 + 128;
                    tc5::mma_ss(tmem + D3_COL, tc5::smem_desc(abase + a0, a1 - a0, T2_GROUP),
                                tc5::smem_desc(w_u + L.w3 + i * 256, 128, 28 * 128), idesc32, i > 0);
                }
                tc5::commit(bar);
            }
            tc5::mbar_wait(bar, phase);
            phase ^= 1;
            tc5::fence_after_sync();
            TC5_ADD(5); }
            { TC5_T0();
            {   // E3
                uint32_t v[32];
                tc5::tmem_ld16(tmem + lane_base + D3_COL, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
                tc5::tmem_ld8(tmem + lane_base + D3_COL + 16, *reinterpret_cast<uint32_t(*)[8]>(&v[16]));
                tc5::tmem_wait_ld();
                const int ly = 4 * m3 + (tid >> 5), lx = tid & 31;
                const int gy = y0 + 1 + ly, gx = x0 + 1 + lx;
                const int p = (ly + 1) * HW_ + lx + 1;
                const uint8_t *res = xs + (p >> 3) * XS_GROUP + (p & 7) * 16;
                if (gy < H && gx < W) {
                    bf16 *o = out + (((long long)n * H + gy) * W + gx) * 24;
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        const uint4 rv = *reinterpret_cast<const uint4 *>(res + q * 128);
                        const uint32_t *rw = reinterpret_cast<const uint32_t *>(&rv);
                        uint4 ov;
                        uint32_t *ow = reinterpret_cast<uint32_t *>(&ov);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float2 r2 = unpack_bf16x2(rw[j]);
                            const int ch = q * 8 + 2 * j;
                            ow[j] = pack_bf16x2(__uint_as_float(v[ch]) + b3s[ch] + r2.x, __uint_as_float(v[ch + 1]) + b3s[ch + 1] + r2.y);
                        }
                        *reinterpret_cast<uint4 *>(o + q * 8) = ov;
                    }
                }
            }
            tc5::fence_before_sync();
            __syncthreads();
            TC5_ADD(6); }
        }
    }
    __syncthreads();
    if (warp == 0) tc5::tmem_free(tmem, TMEM_COLS);
}

inline size_t wdsr_block_tc5_seq_smem(int M1P) {
    using namespace tc5cfg;
    return (size_t)CTRL_BYTES + XS_BYTES + T2_BYTES + BlockTc5Layout(M1P).total + 1024;  // + slack for 1024-B alignment
}



}  // namespace b200sr
