// wdsr_rh.cuh -- row-streaming fused WDSR-B residual block, reduce 1x1 in registers  (models/basic_wdsr_b.py:96-144)
//
//   y = x + conv3x3( conv1x1_reduce( relu( conv1x1_expand(x) ) ) )            t2 (the reduce output) is zero-padded, NOT x
//
// Third form of the block.  What bounded the first two (wdsr_tc5p.cuh tiles, wdsr_rs.cuh rows) is not tensor-queue time but the
// G1 -> E1 -> G2 loop: the expand accumulator of 128 pixels is 144 TMEM columns, two of them fit next to the other accumulators, and
// one trip round the loop (MMA + commit -> warps wake -> tcgen05.ld -> cvt -> tcgen05.st -> arrive -> issuer wakes -> 9 more MMAs)
// is ~1.6-1.8 k clk: two buffers = one 128-pixel step per 0.85-0.95 k clk however the stages are tuned (profiles/r02_block_rs_ncu.md).
// Here the accumulator leaves the tensor-memory path as soon as it has been READ:
//
//   G1  (tcgen05, SS)  D1[128 lanes x 144] = X row . W1^T + b1                         2 MMAs per row, as in wdsr_rs.cuh
//   E1  (two warpgroups, even / odd rows)  tcgen05.ld.16x128b hands a warp the accumulator of 16 lanes in the register layout of the
//        warp-level MMA (up to a permutation of K that the packed W2 fragments follow): relu + bf16x2 pack of its registers IS an A
//        fragment of mma.sync.m16n8k16.  As soon as the loads have
//        retired the D1 buffer is released (the loop is now MMA + commit -> wake -> ld -> arrive -> issuer: ~1 k clk), and the REDUCE
//        1x1 (144 -> 20..24) runs on mma.sync out of registers -- 54 HMMA per warp and row, W2 held as 54 B-fragment registers -- next to
//        the tcgen05 stream (HMMA: 8.2 clk per instruction and sub-partition alone, 11 next to a saturated tcgen05 queue:
//        tools/hmma_bench.cu).  + b2, zero outside the image, bf16 -> the t2 row in shared memory.  No G2 on the tensor queue, no D2, no
//        tcgen05.st, no E2.
//   G3  (tcgen05, SS)  the 3x3 with the three vertical taps stacked in N (N = 96 into three consecutive slots of a 7-slot TMEM ring),
//        4..5 MMAs per row, exactly wdsr_rs.cuh's (same operand image, same A-slice table).
//   E3  OUT slot + b3 + residual -> planar-8 trunk; re-zeroes the slot.
//
// Lane stream, strips, work split and X ring: wdsr_rs.cuh.  The trunk rows are copied by a warp of their own (four lanes of the strip
// per thread, twelve 16-byte cp.async per row, LB rows per batch): fence.proxy.async is a MEMBAR that waits for the thread's copies in
// flight (~a memory latency), so it must not sit in a warp that has anything else to do.  (A fifth warpgroup for it would shrink the
// register pool: setmaxnreg redistributes what the CTA got at launch, 640 threads x 96 < 512 x 128.)
//
// TMEM (512 columns): D1[2] x 144 | OUT[7] x 32.     Threads: 512 = WG0 (three issuer warps + the loader warp) + E1 x 2 + E3.
#pragma once
#include "common.cuh"
#include "tc5.cuh"
#include "wdsr_rs.cuh"
#include "wdsr_rs_layout.cuh"

namespace b200sr {

namespace rh {
using rs::NX;
using rs::SPAN;
using rs::T2PLANE;
using rs::T2SLOT;
using rs::X_BYTES;
using rs::X_ONE;
using rs::XPLANE;
using rs::XSLOT;
constexpr int NTHREADS = 512;
constexpr int TMEM_COLS = 512;
constexpr int NT = 7;                         // t2 ring == OUT ring
constexpr int T2_BYTES = NT * T2SLOT;
constexpr int LB = 8;                         // loader: rows per batch (one memory round trip + proxy fence per batch)
__host__ __device__ constexpr int d1_col(int e) { return e * 144; }
__host__ __device__ constexpr int out_col(int k) { return 288 + k * 32; }
//   X_READY[slot] (32)  = the loader's copies of a trunk row have landed     X_EMPTY[slot] (128) = E3 has consumed the row's residual
//   D1_FREE[e] (128)  = the E1 warpgroup of parity e has read D1[e]: G1(s+2) may be issued
//   G3_READY[b] (256) = E1 wrote t2 row s (b = s % NT)          +  E3 of row s+1-NT re-zeroed OUT slot (s+1) % NT
//   STEP_DONE[b] (1)  = commit after G3(s): row s-1 is complete (E3) and t2 slot b may be overwritten (E1 of step s+NT)
enum Bar { X_READY = 0, X_EMPTY = NX, D1_FULL = 2 * NX, D1_FREE = D1_FULL + 2, G3_READY = D1_FREE + 2, STEP_DONE = G3_READY + NT, NBARS = STEP_DONE + NT };
constexpr int CTRL_BYTES = 512 + NX * 128;    // mbarriers + tmem base pointer at byte 496 | OK[NX][128]: per X-ring row and lane, 1 = a pixel of the image (written by the loader)
static_assert(NBARS * 8 <= 480, "control block");
__host__ __device__ inline size_t smem_bytes(int M1P) { return (size_t)CTRL_BYTES + X_BYTES + T2_BYTES + (size_t)BlockRsLayout(M1P).total; }
}  // namespace rh

// NC2 = 8-channel chunks of t2 = n-tiles of the reduce (3 dense; 2 / 1 for pruned M2 <= 16 / <= 8); PACK: wdsr_rs.cuh.
// NK = K = 16 steps of the reduce (M1P / 16) as a compile-time constant (9 = the dense block: no predication in the ld / HMMA sequence),
//      0 = read from M1P at run time (pruned expand widths).
template <int NC2, bool PACK, int NK>
__global__ void __launch_bounds__(rh::NTHREADS, 1)
wdsr_block_rh_kernel(const bf16 *__restrict__ in, bf16 *__restrict__ out, const uint8_t *__restrict__ wimg, int M1P, int N, int H, int W,
                     int total_rows) {
    using namespace rh;
    using rs::Steps;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const BlockRsLayout L(M1P);
    uint8_t *ctrl = smem_raw;
    uint8_t *xs = smem_raw + CTRL_BYTES;   // NX x XSLOT + constant-one plane  (CTRL_BYTES is a multiple of 512: operand alignment kept)
    uint8_t *t2 = xs + X_BYTES;            // NT x T2SLOT
    uint8_t *wsm = t2 + T2_BYTES;          // L.total
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bars = smem_u32(ctrl);
    auto bar = [&](int b) { return bars + 8u * (uint32_t)b; };
    const uint32_t xs_u = smem_u32(xs), t2_u = smem_u32(t2), w_u = smem_u32(wsm);

    const int g0 = (int)(((long long)blockIdx.x * total_rows) / gridDim.x), g1 = (int)(((long long)(blockIdx.x + 1) * total_rows) / gridDim.x);
    const int T = rs::count_steps(g0, g1, H);
    RS_DECL();

    // ---- one-time setup (may overlap the previous kernel's tail: programmatic stream serialization) ----
    tc5::pdl_launch_dependents();
    if (tid == 0) {
        for (int b = 0; b < NX; ++b) {
            tc5::mbar_init(bar(X_READY + b), 32);
            tc5::mbar_init(bar(X_EMPTY + b), 128);
        }
        for (int e = 0; e < 2; ++e) {
            tc5::mbar_init(bar(D1_FULL + e), 1);
            tc5::mbar_init(bar(D1_FREE + e), 128);
        }
        for (int k = 0; k < NT; ++k) {
            tc5::mbar_init(bar(G3_READY + k), 256);
            tc5::mbar_init(bar(STEP_DONE + k), 1);
        }
        tc5::mbar_init_fence();
    }
    if (warp == 0) tc5::tmem_alloc(smem_u32(ctrl + 496), TMEM_COLS);
    for (int i = tid; i < L.total / 16; i += NTHREADS) cp_async16(wsm + i * 16, wimg + i * 16, 16);
    cp_async_commit();
    for (int i = tid; i < X_ONE / 16; i += NTHREADS) *reinterpret_cast<uint4 *>(xs + i * 16) = make_uint4(0u, 0u, 0u, 0u);
    for (int i = tid; i < XPLANE / 16; i += NTHREADS) *reinterpret_cast<uint4 *>(xs + X_ONE + i * 16) = make_uint4(0x3F803F80u, 0u, 0u, 0u);
    for (int i = tid; i < T2_BYTES / 16; i += NTHREADS) *reinterpret_cast<uint4 *>(t2 + i * 16) = make_uint4(0u, 0u, 0u, 0u);
    cp_async_wait<0>();
    tc5::fence_proxy_async();
    tc5::fence_before_sync();
    __syncthreads();
    tc5::fence_after_sync();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t *>(ctrl + 496);

    const int wg = warp >> 2;
    if (wg == 0) {
        tc5::setmaxnreg_dec<56>();   // budgets (x 128 threads): WG0 56 + E1 2 x 192 + E3 72 = 4 x 128
        if (warp < 2) {
            // ============================== issuer A (even / odd rows): G1 (expand) ==============================
            // Two warps: a thread that has issued tcgen05.mma stalls on its next mbarrier probe until its MMAs have drained.
            const int e = warp;
            const bool leader = tc5::elect_one();
            const uint32_t idesc1 = tc5::idesc_bf16_f32(128, M1P);
            const uint64_t bw1a = tc5::smem_desc(w_u + L.w1, 128, 512), bw1b = tc5::smem_desc(w_u + L.w1 + 256, 128, 512);
            auto issue_g1 = [&](int s) {  // leader only
                const int slot = s % NX;
                const uint32_t base = xs_u + slot * XSLOT;
#ifndef RH_EXP_NOG1   // (timing experiment)
                tc5::mma_ss(tmem + d1_col(e), tc5::smem_desc(base, XPLANE, 128), bw1a, idesc1, false);                                // planes 0,1
                tc5::mma_ss(tmem + d1_col(e), tc5::smem_desc(base + 2 * XPLANE, X_ONE - slot * XSLOT - 2 * XPLANE, 128), bw1b, idesc1, true);  // plane 2, ONE
#endif
                tc5::commit(bar(D1_FULL + e));
            };
            if (e < T) {
                tc5::mbar_wait(bar(X_READY + e), 0);
                tc5::fence_after_sync();
                if (leader) issue_g1(e);
                __syncwarp();
            }
            for (int s = e; s + 2 < T; s += 2) {
                tc5::mbar_wait(bar(X_READY + (s + 2) % NX), ((s + 2) / NX) & 1);   // (rows are copied ~LB..NX ahead: normally long complete)
                RS_EVT(101);
                tc5::mbar_wait(bar(D1_FREE + e), (s >> 1) & 1);
                tc5::fence_after_sync();
                RS_EVT(100);
                if (leader) issue_g1(s + 2);   // D1[e] has been read into registers
                __syncwarp();
                RS_EVT(102);
            }
        } else if (warp == 2) {
            // ============================== issuer B: G3 (3x3, dy stacked in N) ==============================
            const bool leader = tc5::elect_one();
            const uint32_t idesc96 = tc5::idesc_bf16_f32(128, 96), idesc64 = tc5::idesc_bf16_f32(128, 64), idesc32 = tc5::idesc_bf16_f32(128, 32);
            const int *tab = reinterpret_cast<const int *>(wsm + L.tab);
            const int ng3 = tab[0];
            uint32_t alo[BlockRsLayout::MAXG3];
#pragma unroll
            for (int i = 0; i < BlockRsLayout::MAXG3; ++i)
                alo[i] = (uint32_t)tc5::smem_desc(t2_u + (uint32_t)tab[1 + i], (uint32_t)tab[1 + BlockRsLayout::MAXG3 + i], 128);
            const uint32_t ahi = (uint32_t)(tc5::smem_desc(0, 0, 128) >> 32);
            const uint64_t bw3 = tc5::smem_desc(w_u + L.w3, 128, L.sbo3);
            const uint32_t grp = (uint32_t)((4 * L.sbo3) >> 4);   // 32 B rows (one dy group) further into the B image
            for (int s = 0; s < T; ++s) {
                const int b = s % NT;
                tc5::mbar_wait(bar(G3_READY + b), (s / NT) & 1);
                tc5::fence_after_sync();
                RS_EVT(200);
                if (leader) {
                    const uint32_t aoff = (uint32_t)((b * T2SLOT) >> 4);
                    const int a = (s + NT - 1) % NT;   // OUT slot of row s-1; rows s and s+1 follow (mod NT)
                    auto batch = [&](int slot, uint64_t bdesc, uint32_t idesc) {
#pragma unroll
                        for (int i = 0; i < BlockRsLayout::MAXG3; ++i)
                            if (i < ng3) tc5::mma_ss(tmem + out_col(slot), ((uint64_t)ahi << 32) | (uint64_t)(alo[i] + aoff), bdesc + (uint64_t)(16 * i), idesc, true);
                    };
#ifdef RH_EXP_NOG3   // (timing experiment)
                    if (false) {
#else
                    if (a <= NT - 3) {
#endif
                        batch(a, bw3, idesc96);
#ifdef RH_EXP_NOG3
                    } else if (false) {
#else
                    } else if (a == NT - 2) {   // rows s-1, s in the last two slots; row s+1 in slot 0
#endif
                        batch(NT - 2, bw3, idesc64);
                        batch(0, bw3 + 2 * grp, idesc32);
#ifdef RH_EXP_NOG3
                    } else if (false) {
#else
                    } else {                    // row s-1 in the last slot (nothing there on the CTA's first step); rows s, s+1 in slots 0, 1
#endif
                        if (s > 0) batch(NT - 1, bw3, idesc32);
                        batch(0, bw3 + grp, idesc64);
                    }
                    tc5::commit(bar(STEP_DONE + b));
                }
                __syncwarp();
                RS_EVT(201);
            }
            if (T > 0) tc5::mbar_wait(bar(STEP_DONE + (T - 1) % NT), ((T - 1) / NT) & 1);   // every G3 retired
        } else {
            // ============================== loader (warp 3): trunk rows -> X ring ==============================
            Steps itl(g0, g1, H);
            bool lfresh = true;
            const uint8_t *src[4] = {nullptr, nullptr, nullptr, nullptr};   // strip lanes lane, lane + 32, lane + 64, lane + 96
            tc5::pdl_wait();   // the previous kernel's trunk is complete and visible (everything downstream, the output stores included, follows)
            for (int r0 = 0; r0 < T; r0 += LB) {
#pragma unroll 1
                for (int r = r0; r < r0 + LB && r < T; ++r) {
                    if (lfresh) {
#pragma unroll
                        for (int q4 = 0; q4 < 4; ++q4) {
                            const long long px = rs::lane_pixel(itl.strip, lane + 32 * q4, N, H, W);
                            src[q4] = px >= 0 ? reinterpret_cast<const uint8_t *>(in) + px * 16 : nullptr;
                        }
                    }
                    const int slot = r % NX;
                    if (r >= NX) tc5::mbar_wait(bar(X_EMPTY + slot), ((r / NX) - 1) & 1);
                    const bool iny = itl.in_image();
                    const long long yoff = (long long)itl.y * W * 16;
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4) {
                        const bool okp = src[q4] && iny;
                        // flags of strip lane l = 32 q4 + lane: byte [wq = q4][g = l & 7][2 mt + h = (l >> 3) & 3], see E1
                        ctrl[512 + slot * 128 + 32 * q4 + 4 * (lane & 7) + (lane >> 3)] = okp ? 1 : 0;
                        if (okp) {
                            const uint8_t *sp = src[q4] + yoff;
                            uint8_t *dp = xs + slot * XSLOT + (lane + 32 * q4) * 16;
#pragma unroll
                            for (int q = 0; q < 3; ++q) cp_async16(dp + q * XPLANE, sp + (long long)q * H * W * 16, 16);
                        }
                    }
                    if (r + 1 < T) lfresh = itl.advance();
                }
                RS_EVT(600);
                cp_async_commit();
                cp_async_wait<0>();
                RS_EVT(601);
                tc5::fence_proxy_async();
                for (int r = r0; r < r0 + LB && r < T; ++r) tc5::mbar_arrive(bar(X_READY + r % NX));
                RS_EVT(602);
            }
        }
    } else if (wg <= 2) {
        // ============================== E1 (WG1: even rows, WG2: odd rows) ==============================
        tc5::setmaxnreg_inc<192>();
        const int e = wg - 1;
        const int wq = warp & 3, g = lane >> 2, j = lane & 3;
        const int nk = NK ? NK : (M1P >> 4);
        const int once = total_rows > 0 ? 1 : 2;   // == 1
        const float *b2s = reinterpret_cast<const float *>(wsm + L.b2);
        // reduce filter: B fragments, resident in registers
        uint32_t bw[9][NC2][2];
#pragma unroll
        for (int k = 0; k < 9; ++k)
#pragma unroll
            for (int nt = 0; nt < NC2; ++nt) {
                uint2 t = make_uint2(0u, 0u);
                if (k < nk) t = *reinterpret_cast<const uint2 *>(wsm + L.w2f + ((k * 3 + nt) * 32 + lane) * 8);
                bw[k][nt][0] = t.x, bw[k][nt][1] = t.y;
            }

        // OK table: the four flags of a thread's pixels (m-tile mt, row half h) are one 32-bit word: byte (32 wq + 4 g + 2 mt + h) of the row's entry
        const uint8_t *okt = ctrl + 512 + 32 * wq + 4 * g;
        for (int s = e; s < T; s += 2) {
            float2 bb[NC2];   // b2 (read before the wait: off the critical path; the accumulators start from it)
#pragma unroll
            for (int nt = 0; nt < NC2; ++nt) bb[nt] = *reinterpret_cast<const float2 *>(b2s + 8 * nt + 2 * j);
            const int b = s % NT;
            tc5::mbar_wait(bar(D1_FULL + e), (s >> 1) & 1);
            tc5::fence_after_sync();
            RS_EVT(300);
            // ---- D1 -> registers: relu + pack = A fragments of the reduce (K order of the 16x128b shape, see tc5.cuh)
            uint32_t A[2][9][4];
            uint32_t okw = 0;
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                const uint32_t ta = tmem + ((uint32_t)(32 * wq + 16 * mt) << 16) + d1_col(e);
                uint32_t raw[12][8];
#pragma unroll
                for (int c = 0; c < 3; ++c) {   // 64 columns (4 K steps) per instruction where they exist, else 32 and / or 16
                    const int n = nk - 4 * c;   // K steps left
                    if (n >= 4) tc5::tmem_ld_16x128b_x16(ta + 64 * c, &raw[4 * c][0]);
                    else {
                        if (n >= 2) tc5::tmem_ld_16x128b_x8(ta + 64 * c, &raw[4 * c][0]);
                        if (n == 3) tc5::tmem_ld_16x128b_x4(ta + 64 * c + 32, &raw[4 * c + 2][0]);
                        if (n == 1) tc5::tmem_ld_16x128b_x4(ta + 64 * c, &raw[4 * c][0]);
                    }
                }
                if (mt == 1) {   // in the shadow of the loads: the row's flags, and "G3 of step s-NT has read this t2 slot" (long true)
                    okw = *reinterpret_cast<const uint32_t *>(okt + (s % NX) * 128);   // (visible: loader's X_READY arrive -> issuer -> G1 commit -> D1_FULL)
                    if (s >= NT) tc5::mbar_wait(bar(STEP_DONE + b), ((s / NT) - 1) & 1);
                }
                tc5::tmem_wait_ld();
                if (mt == 1) {
                    tc5::fence_before_sync();
                    tc5::mbar_arrive_relaxed(bar(D1_FREE + e));   // (the loads have retired: nothing to publish)
                    RS_EVT(301);
                }
#pragma unroll
                for (int k = 0; k < 9; ++k)
                    if (k < nk) {
                        A[mt][k][0] = tc5::relu_pack_bf16x2(raw[k][0], raw[k][2]);
                        A[mt][k][1] = tc5::relu_pack_bf16x2(raw[k][1], raw[k][3]);
                        A[mt][k][2] = tc5::relu_pack_bf16x2(raw[k][4], raw[k][6]);
                        A[mt][k][3] = tc5::relu_pack_bf16x2(raw[k][5], raw[k][7]);
                    }
            }
            RS_EVT_DEP(305, __uint_as_float(A[1][nk - 1][3] ^ A[0][0][0]));
            // ---- reduce 1x1 on mma.sync (six accumulator chains per warp; a second set for the odd K steps changed nothing: the phase is bound
            //      by the HMMA issue rate with fresh operand registers, 13.5-16 clk per instruction and sub-partition, tools/hmma_bench.cu)
            float acc[2][NC2][4];
#pragma unroll
            for (int nt = 0; nt < NC2; ++nt)
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) acc[mt][nt][0] = acc[mt][nt][2] = bb[nt].x, acc[mt][nt][1] = acc[mt][nt][3] = bb[nt].y;
            // (a loop of ONE iteration whose trip count ptxas cannot see: without it the HMMAs are scheduled in between the tcgen05.ld's, ahead
            //  of the D1_FREE arrive -- the critical G1 -> ld -> arrive -> G1 loop would wait for ~45 HMMA issue slots)
#pragma unroll 1
            for (int rep = 0; rep < once; ++rep) {
#pragma unroll
                for (int k = 0; k < 9; ++k)
                    if (k < nk) {
#pragma unroll
                        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                            for (int nt = 0; nt < NC2; ++nt) tc5::hmma_16816(acc[mt][nt], A[mt][k], bw[k][nt]);
                    }
            }
            RS_EVT_DEP(302, acc[1][NC2 - 1][3] + acc[0][0][0]);
            // ---- zero outside the image, bf16 -> t2 row (chunk-planar, lane l at entry l + 1)
            uint8_t *dst0 = t2 + b * T2SLOT + 16 + 4 * j;
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const bool ok = ((okw >> (8 * (2 * mt + h))) & 0xffu) != 0;
                    uint8_t *dst = dst0 + (32 * wq + 16 * mt + 8 * h + g) * 16;
#pragma unroll
                    for (int nt = 0; nt < NC2; ++nt) {
                        const uint32_t v = ok ? pack_bf16x2(acc[mt][nt][2 * h], acc[mt][nt][2 * h + 1]) : 0u;
                        if (PACK && nt == 2) {   // channels 16..19 twice: low half of the lane's own entry, high half of the entry on its left
                            if (j < 2) {
                                *reinterpret_cast<uint32_t *>(dst + 2 * T2PLANE) = v;
                                *reinterpret_cast<uint32_t *>(dst + 2 * T2PLANE - 16 + 8) = v;
                            }
                        } else {
                            *reinterpret_cast<uint32_t *>(dst + nt * T2PLANE) = v;
                        }
                    }
                }
            RS_EVT(307);
            tc5::fence_proxy_async();
            tc5::mbar_arrive_relaxed(bar(G3_READY + b));   // (fence.proxy.async above is a CTA-scope MEMBAR: the stores are performed)
            RS_EVT(303);
        }
    } else {
        // ============================== E3 (WG3) ==============================
        tc5::setmaxnreg_dec<72>();
        const int row = (warp & 3) * 32 + lane;  // lane of the strip == TMEM lane
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const float *b3s = reinterpret_cast<const float *>(wsm + L.b3);
        {
            const uint32_t z[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
#pragma unroll
            for (int k = 0; k < NT * 4; ++k) tc5::tmem_st8(tmem + lane_base + out_col(0) + 8 * k, z);
            tc5::tmem_wait_st();
            tc5::fence_before_sync();
#pragma unroll
            for (int k = 0; k < NT - 1; ++k) tc5::mbar_arrive(bar(G3_READY + k));  // stand-ins: steps 0..NT-2 find their new slot zeroed
        }
        Steps it(g0, g1, H);
        bool fresh = true;
        long long px = -1;
        for (int r = 0; r + 1 < T; ++r) {
            if (r > 0) fresh = it.advance();
            if (fresh) px = rs::lane_pixel(it.strip, row, N, H, W);
            tc5::mbar_wait(bar(STEP_DONE + (r + 1) % NT), ((r + 1) / NT) & 1);
            tc5::fence_after_sync();
            RS_EVT(500);
            uint32_t v[32];
            const uint32_t oc = tmem + lane_base + out_col(r % NT);
            tc5::tmem_ld32(oc, v);
            const int xslot = r % NX;
            // (the copied row is visible: loader arrive -> D1_FREE -> issuer -> ... -> STEP_DONE, acquired above)
            const uint8_t *res = xs + xslot * XSLOT + row * 16;
            uint4 rv[3];
#pragma unroll
            for (int q = 0; q < 3; ++q) rv[q] = *reinterpret_cast<const uint4 *>(res + q * XPLANE);
            tc5::tmem_wait_ld();
            {
                const uint32_t z[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
#pragma unroll
                for (int k = 0; k < 3; ++k) tc5::tmem_st8(oc + 8 * k, z);   // columns 24..31 only ever accumulate zero weights
            }
            tc5::tmem_wait_st();
            tc5::fence_before_sync();
            tc5::mbar_arrive_relaxed(bar(G3_READY + (r + NT - 1) % NT));   // slot r % NT is zero again: step r+NT-1 may start row r+NT in it
            tc5::mbar_arrive_relaxed(bar(X_EMPTY + xslot));                 // residual values are in registers
            RS_EVT(501);
            if (it.stored() && px >= 0 && row >= 1 && row <= SPAN) {
                bf16 *o = out + (px + (long long)it.y * W) * 8;   // planar-8 trunk: plane q is H*W*8 elements further
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    const uint32_t *rw = reinterpret_cast<const uint32_t *>(&rv[q]);
                    uint4 ov;
                    uint32_t *ow = reinterpret_cast<uint32_t *>(&ov);
#pragma unroll
                    for (int j2 = 0; j2 < 2; ++j2) {
                        const float4 bb = *reinterpret_cast<const float4 *>(b3s + q * 8 + 4 * j2);  // broadcast read
                        const float2 ra = unpack_bf16x2(rw[2 * j2]), rb = unpack_bf16x2(rw[2 * j2 + 1]);
                        const int ch = q * 8 + 4 * j2;
                        ow[2 * j2] = pack_bf16x2(__uint_as_float(v[ch]) + bb.x + ra.x, __uint_as_float(v[ch + 1]) + bb.y + ra.y);
                        ow[2 * j2 + 1] = pack_bf16x2(__uint_as_float(v[ch + 2]) + bb.z + rb.x, __uint_as_float(v[ch + 3]) + bb.w + rb.y);
                    }
                    *reinterpret_cast<uint4 *>(o + (long long)q * H * W * 8) = ov;
                }
            }
            RS_EVT(502);
        }
    }
    RS_FLUSH();
    tc5::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc5::tmem_free(tmem, TMEM_COLS);
}

}  // namespace b200sr
