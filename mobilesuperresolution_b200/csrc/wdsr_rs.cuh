// wdsr_rs.cuh -- row-streaming form of the tcgen05 fused WDSR-B residual block  (models/basic_wdsr_b.py:96-144)
//
//   y = x + conv3x3( conv1x1_reduce( relu( conv1x1_expand(x) ) ) )            t2 (the reduce output) is zero-padded, NOT x
//
// Why a second form.  The tile form (wdsr_tc5p.cuh) contracts the 3x3 over K = 9 taps x 24 channels with N = 32: 14 MMAs per
// 128 pixels whose A operand (4 KB of shared memory each) buys only 16 clk of tensor work -- 56 A-fetch-bound instructions per
// tile, 2.3 k of the tile's 3.8 k clk of tensor-queue time.  Here an M-tile is ONE ROW of 128 consecutive pixels ("lanes") and a
// CTA marches down the rows of its strip:
//
//   * the three vertical taps of the 3x3 are stacked in N:  D[lane, (dy, cout)] += t2row[lane + dx - 1, :] . W3[dy][dx]   (N = 96,
//     K = 3 dx x 20..24 channels = 4..5 MMAs per row instead of 14 per 128 pixels).  Row s of t2 feeds output rows s-1 (dy = 2),
//     s (dy = 1) and s+1 (dy = 0); those three accumulators are three CONSECUTIVE 32-column slots of a 5-slot TMEM ring, so one
//     instruction covers them and the sum over dy happens inside the tensor core (always-accumulate; the epilogue re-zeroes a slot
//     after it has read it).  Lane l of every row is the same image column, so all three partial sums are lane aligned.
//     Where the ring wraps (2 steps of 5) the instruction is split into an N = 64 and an N = 32 one.
//   * a horizontal tap is a 16-byte offset of the A operand's start address (t2 rows are chunk-planar [8-channel chunk][lane][16 B]);
//     lanes 0 and 127 of a strip are halo lanes: they hold t2 for their neighbours, their own 3x3 result is never stored.
//   * 20 reduce channels: the 4 channels of the third chunk are stored twice ([lane l | lane l+1] in one 16-byte entry), so two
//     horizontal taps share one K = 8 half: K = 3 x 16 + 8 + 8 = 64 -> 4 MMAs per row ("PACK"; 21..24 channels take 5).
//   * the lanes of a strip are a window of 128 slots of the LANE STREAM: the rows of all images laid end to end, W + 2 slots per
//     image (x = -1 .. W; the two halo slots stay zero = the conv's zero padding), strips overlap by 2 slots.  Narrow images
//     (cfg2: 96 px) therefore still fill 96 % of the 128 lanes.  No TMA tensor map: rows are copied with plain 16-byte cp.async,
//     one per thread of three of the epilogue warpgroups (thread = lane of the strip, warpgroup = plane).
//   * work split: the nstrips x H output rows are dealt to the CTAs as equal contiguous ranges; a range is 1..2 "units" (strip,
//     rows ya..yb-1) and costs 2 extra t2 rows per unit.
//
// Tensor-queue time per row of 126 output pixels: G1 2 x 69 + G2 9 x 18 + G3 4 x 56 (x 1.6 on the two wrap steps of five) ~ 575 clk
// = 4.6 clk/px; the tile form needs 7.4 clk/px (3.8 k clk per 512 px).
//
// Roles (1024 threads, eight warpgroups):
//   WG0 warps 0/1  issuer A of the even / odd steps: G2(s), G1(s+2)      warp 2  issuer B: G3(s)      warp 3  idle
//   WG1..WG3    also load the trunk rows into the X ring (NX slots): one 16-byte cp.async per thread and row, 8 rows ahead
//   WG1..WG4    E1         relu(D1) -> packed bf16 A operand of G2, in place; each warpgroup owns whole K = 16 steps of G2
//                          (expand channels 0..31 / 32..63 / 64..95 / 96..143).  A TMEM round trip (ld -> cvt -> st -> wait) is
//                          ~400 clk of latency however few columns a thread moves (tools/rs_e1_bench.cu: 64 + 80 columns on
//                          two warpgroups 770 clk, four warpgroups 440), and E1 sits on the G1 -> E1 -> G2 loop of the two D1 buffers
//   WG5 / WG6   E2         even / odd steps: D2 + b2 -> bf16 t2 row in shared memory (zero outside the image)
//   WG7         E3         OUT slot + b3 + residual -> planar-8 trunk; re-zero the slot
// TMEM (512 columns): D1[2] x 144 | D2[2] x 32 | OUT[5] x 32.
#pragma once
#include <type_traits>

#include "common.cuh"
#include "tc5.cuh"
#include "wdsr_rs_layout.cuh"

#ifdef B200SR_RS_PROF
// developer probe (tools/rs_probe.cu): lane 0 of every warp of CTA 0 logs (event id << 48 | clock64)
__device__ unsigned long long g_rs_evt[32][4096];
__device__ int g_rs_evtn[32];
#define RS_DECL() int evn__ = 0
#ifndef B200SR_RS_EVT_LO
#define B200SR_RS_EVT_LO 0
#define B200SR_RS_EVT_HI 1000
#endif
#define RS_EVT(id) do { if ((id) >= B200SR_RS_EVT_LO && (id) < B200SR_RS_EVT_HI && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && evn__ < 4096) { g_rs_evt[threadIdx.x >> 5][evn__++] = ((unsigned long long)(id) << 48) | ((unsigned long long)clock64() & 0xFFFFFFFFFFFFull); } } while (0)
#define RS_FLUSH() do { if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) g_rs_evtn[threadIdx.x >> 5] = evn__; } while (0)
// the clock is read under a predicate computed from `dep` (a float the event must come after): ptxas cannot schedule the read ahead of it
#define RS_EVT_DEP(id, dep) do { if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && evn__ < 4096) { unsigned long long t__; \
    asm volatile("{\n\t.reg .pred p;\n\tsetp.neu.f32 p, %1, 0f7F812345;\n\t@p mov.u64 %0, %%clock64;\n\t@!p mov.u64 %0, 0;\n\t}" : "=l"(t__) : "f"(dep)); \
    g_rs_evt[threadIdx.x >> 5][evn__++] = ((unsigned long long)(id) << 48) | (t__ & 0xFFFFFFFFFFFFull); } } while (0)
#else
#define RS_EVT_DEP(id, dep) do {} while (0)
#define RS_DECL() do {} while (0)
#define RS_EVT(id) do {} while (0)
#define RS_FLUSH() do {} while (0)
#endif

#ifdef B200SR_RS_TIMERS
// developer probe (tools/rs_probe.cu -DB200SR_RS_TIMERS): per-warp cycle accumulators kept in registers, flushed once by lane 0 of CTA 0
__device__ unsigned long long g_rs_tm[32][8];
#define RS_TM_DECL() long long tm__[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long last__ = clock64()
#define RS_MARK(i) do { const long long n__ = clock64(); tm__[i] += n__ - last__; last__ = n__; } while (0)
#define RS_TM_FLUSH() do { if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) { for (int i__ = 0; i__ < 8; ++i__) g_rs_tm[threadIdx.x >> 5][i__] = (unsigned long long)tm__[i__]; } } while (0)
#else
#define RS_TM_DECL() do {} while (0)
#define RS_MARK(i) do {} while (0)
#define RS_TM_FLUSH() do {} while (0)
#endif

// the two waits on the G1 -> E1 -> G2 loop (developer switch: -DB200SR_RS_SPIN spins on the non-blocking probe instead of suspending)
#ifdef B200SR_RS_SPIN
#define RS_CRIT_WAIT(b, p) tc5::mbar_spin(b, p)
#else
#define RS_CRIT_WAIT(b, p) tc5::mbar_wait(b, p)
#endif

namespace b200sr {

namespace rs {
// The CTA's share of the nstrips x H output rows, walked as t2 rows ("steps"): a unit (strip, ya..yb-1) is yb - ya + 2 steps.
struct Steps {
    int g, g1, H, strip, ya, yb, y;
    __device__ Steps(int g0, int g1_, int H_) : g(g0), g1(g1_), H(H_) { unit(); }
    __device__ void unit() {
        strip = g / H;
        ya = g - strip * H;
        yb = ya + (g1 - g) < H ? ya + (g1 - g) : H;
        y = ya - 1;
    }
    __device__ bool advance() {   // true when a new unit began (the strip may have changed)
        if (y < yb) { ++y; return false; }
        g += yb - ya;
        unit();
        return true;
    }
    __device__ bool stored() const { return y >= ya && y < yb; }
    __device__ bool in_image() const { return y >= 0 && y < H; }
};
__device__ inline int count_steps(int g0, int g1, int H) {
    int t = 0;
    for (int g = g0; g < g1;) {
        const int ya = g % H, n = (H - ya) < (g1 - g) ? (H - ya) : (g1 - g);
        t += n + 2, g += n;
    }
    return t;
}
// lane -> pixel of the strip: plane-0 / row-0 pixel index of the lane's image column in the planar-8 trunk, or -1 (halo slot, past the end)
__device__ inline long long lane_pixel(int strip, int lane, int N, int H, int W) {
    const int slot = strip * SPAN + lane;
    if (slot >= N * (W + 2)) return -1;
    const int n = slot / (W + 2), xs = slot - n * (W + 2);
    if (xs < 1 || xs > W) return -1;
    return (long long)n * 3 * H * W + (xs - 1);
}
}  // namespace rs

// NC2 = 8-channel chunks of t2 (3 dense; 2 / 1 for pruned M2 <= 16 / <= 8); PACK = the third chunk holds <= 4 channels, stored
// twice so that two horizontal taps share one K = 8 half (NC2 == 3 only).  The matching A-slice table comes with the weight image.
template <int NC2, bool PACK>
__global__ void __launch_bounds__(rs::NTHREADS, 1)
wdsr_block_rs_kernel(const bf16 *__restrict__ in, bf16 *__restrict__ out, const uint8_t *__restrict__ wimg, int M1P, int N, int H, int W,
                     int total_rows) {
    using namespace rs;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const BlockRsLayout L(M1P);
    uint8_t *ctrl = smem_raw;
    uint8_t *xs = smem_raw + CTRL_BYTES;   // NX x XSLOT + constant-one plane
    uint8_t *t2 = xs + X_BYTES;            // NT x T2SLOT
    uint8_t *wsm = t2 + T2_BYTES;          // L.total
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bars = smem_u32(ctrl);
    auto bar = [&](int b) { return bars + 8u * (uint32_t)b; };
    const uint32_t xs_u = smem_u32(xs), t2_u = smem_u32(t2), w_u = smem_u32(wsm);

    const int g0 = (int)(((long long)blockIdx.x * total_rows) / gridDim.x), g1 = (int)(((long long)(blockIdx.x + 1) * total_rows) / gridDim.x);
    const int T = count_steps(g0, g1, H);
    RS_DECL();
    RS_TM_DECL();

    // ---- one-time setup (may overlap the previous kernel's tail: programmatic stream serialization) ----
    tc5::pdl_launch_dependents();
    if (tid == 0) {
        tc5::mbar_init(bar(X_START), 3 * 128);       // rows 0..2 have landed (the loader threads, once)
        for (int b = 0; b < NX; ++b) tc5::mbar_init(bar(X_EMPTY + b), 128);   // the 128 threads of E3 after the row's residual was consumed
        for (int e = 0; e < 2; ++e) {
            tc5::mbar_init(bar(D1_FULL + e), 1);
            tc5::mbar_init(bar(G2_READY + e), 4 * 128 + 128);   // E1 warpgroups (A2 written, trunk row s+2 landed) + E2 (D2 drained)
            tc5::mbar_init(bar(D2_FULL + e), 1);
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
    // X ring and t2 ring start as zeros: halo lanes are never written by the row copies, and the zero-weight halves of the 3x3
    // slices must read finite values.  (Done before any copy is in flight: the producer starts after the barrier below.)
    for (int i = tid; i < (X_ONE + 0) / 16; i += NTHREADS) *reinterpret_cast<uint4 *>(xs + i * 16) = make_uint4(0u, 0u, 0u, 0u);
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
        tc5::setmaxnreg_dec<56>();   // budgets (x 128 threads): WG0 56 + E1 4 x 64 + E2 2 x 64 + E3 72 = 8 x 64
        // Three issuer warps.
        //   warps 0 / 1  issuer A of the even / odd steps: G2(s), then G1(s+2).  Everything that touches D1[e] / D2[e] comes from ONE
        //                thread, so "G1(s+2) overwrites D1[e] after G2(s) has read it" is plain issue order.  Two of them because a
        //                thread that has issued tcgen05.mma stalls on its next mbarrier probe until its queued MMAs have drained:
        //                while one drains, the other -- idle since the step before last -- takes the next step.
        //   warp 2       issuer B: G3(s) of every step, in order.  ONE thread on purpose: rows s-1, s, s+1 accumulate into the same
        //                TMEM slot, and only a single issue stream fixes the order of those fp32 additions (two alternating 3x3
        //                issuers were measured no faster and made the low bits depend on which thread's batch reached the queue first).
        // (tools/rs_umma_bench.cu: the three streams need ~630 clk of tensor-queue time per step however they are issued.)
        const int e = warp & 1;
        if (warp < 2) {
            // ============================== issuer A: G1 (expand) and G2 (reduce) ==============================
            const bool leader = tc5::elect_one();
            const uint32_t idesc1 = tc5::idesc_bf16_f32(128, M1P), idesc32 = tc5::idesc_bf16_f32(128, 32);
            const uint64_t bw1a = tc5::smem_desc(w_u + L.w1, 128, 512), bw1b = tc5::smem_desc(w_u + L.w1 + 256, 128, 512);
            const uint64_t bw2 = tc5::smem_desc(w_u + L.w2, 128, L.sbo2);
            const int nk2 = M1P / 16;
            auto issue_g1 = [&](int s) {  // leader only
                const int slot = s % NX;
                const uint32_t base = xs_u + slot * XSLOT;
                tc5::mma_ss(tmem + d1_col(e), tc5::smem_desc(base, XPLANE, 128), bw1a, idesc1, false);                                // planes 0,1
                tc5::mma_ss(tmem + d1_col(e), tc5::smem_desc(base + 2 * XPLANE, X_ONE - slot * XSLOT - 2 * XPLANE, 128), bw1b, idesc1, true);  // plane 2, ONE
                tc5::commit(bar(D1_FULL + e));
            };
            if (e < T) {
                tc5::mbar_wait(bar(X_START), 0);
                tc5::fence_after_sync();
                if (leader) issue_g1(e);
                __syncwarp();
            }
            RS_MARK(7);
            for (int s = e; s < T; s += 2) {
                RS_CRIT_WAIT(bar(G2_READY + e), (s >> 1) & 1);
                RS_MARK(0);
                RS_EVT(100);
                tc5::fence_after_sync();
                if (leader) {
                    const uint32_t d2 = tmem + d2_col(e), a2 = tmem + d1_col(e);
                    // A2 columns of K step j: each E1 warpgroup packs into the start of its own column range (see E1)
                    tc5::mma_ts(d2, a2, bw2, idesc32, false);
#pragma unroll
                    for (int j = 1; j < 9; ++j)
                        if (j < nk2) tc5::mma_ts(d2, a2 + a2_col(j), bw2 + (uint64_t)(16 * j), idesc32, true);
                    tc5::commit(bar(D2_FULL + e));
                    if (s + 2 < T) issue_g1(s + 2);   // D1[e] is free once G2(s) has read it: same thread, in order
                }
                __syncwarp();
                RS_MARK(2);
                RS_EVT(102);
            }
            const int last = ((T - 1 - e) & ~1) + e;   // this issuer's last step
            if (e < T) tc5::mbar_wait(bar(D2_FULL + e), (last >> 1) & 1);   // every G1 / G2 of this thread retired
        } else if (warp == 2) {
            // ============================== issuer B: G3 (3x3, dy stacked in N) ==============================
            const bool leader = tc5::elect_one();
            const uint32_t idesc96 = tc5::idesc_bf16_f32(128, 96), idesc64 = tc5::idesc_bf16_f32(128, 64), idesc32 = tc5::idesc_bf16_f32(128, 32);
            const int *tab = reinterpret_cast<const int *>(wsm + L.tab);
            const int ng3 = tab[0];
            // A descriptors: only the low words differ (start address | LBO << 16); the high word (SBO = 128 B, version 1) is shared
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
                    if (a <= NT - 3) {
#pragma unroll
                        for (int i = 0; i < BlockRsLayout::MAXG3; ++i)
                            if (i < ng3) tc5::mma_ss(tmem + out_col(a), ((uint64_t)ahi << 32) | (uint64_t)(alo[i] + aoff), bw3 + (uint64_t)(16 * i), idesc96, true);
                    } else if (a == NT - 2) {   // rows s-1, s in the last two slots; row s+1 in slot 0
#pragma unroll
                        for (int i = 0; i < BlockRsLayout::MAXG3; ++i)
                            if (i < ng3) tc5::mma_ss(tmem + out_col(NT - 2), ((uint64_t)ahi << 32) | (uint64_t)(alo[i] + aoff), bw3 + (uint64_t)(16 * i), idesc64, true);
#pragma unroll
                        for (int i = 0; i < BlockRsLayout::MAXG3; ++i)
                            if (i < ng3) tc5::mma_ss(tmem + out_col(0), ((uint64_t)ahi << 32) | (uint64_t)(alo[i] + aoff), bw3 + 2 * grp + (uint64_t)(16 * i), idesc32, true);
                    } else {                    // row s-1 in the last slot (nothing there on the CTA's first step); rows s, s+1 in slots 0, 1
                        if (s > 0) {
#pragma unroll
                            for (int i = 0; i < BlockRsLayout::MAXG3; ++i)
                                if (i < ng3) tc5::mma_ss(tmem + out_col(NT - 1), ((uint64_t)ahi << 32) | (uint64_t)(alo[i] + aoff), bw3 + (uint64_t)(16 * i), idesc32, true);
                        }
#pragma unroll
                        for (int i = 0; i < BlockRsLayout::MAXG3; ++i)
                            if (i < ng3) tc5::mma_ss(tmem + out_col(0), ((uint64_t)ahi << 32) | (uint64_t)(alo[i] + aoff), bw3 + grp + (uint64_t)(16 * i), idesc64, true);
                    }
                    tc5::commit(bar(STEP_DONE + b));
                }
                __syncwarp();
                RS_EVT(201);
            }
            if (T > 0) tc5::mbar_wait(bar(STEP_DONE + (T - 1) % NT), ((T - 1) / NT) & 1);   // every G3 retired
        }
    } else {
        // ============================== epilogue warpgroups ==============================
        const int row = (warp & 3) * 32 + lane;  // lane of the strip == TMEM lane
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const float *b2s = reinterpret_cast<const float *>(wsm + L.b2);
        const float *b3s = reinterpret_cast<const float *>(wsm + L.b3);

        if (wg <= 4) {
            // ---- E1 (WG1..WG4): relu(D1) -> bf16 A2, packed in place at the start of the warpgroup's own column range.
            //      Warpgroup q owns expand channels 32q .. 32q+31 (q < 3) / 96 .. M1P-1 (q = 3): whole K = 16 steps of G2.
            //      Each thread reads ALL of its columns before it writes.
            const int q = wg - 1;
            const int ncol = q < 3 ? (M1P - 32 * q >= 32 ? 32 : M1P - 32 * q > 0 ? 16 : 0) : (M1P > 96 ? M1P - 96 : 0);
            auto cvt = [&](uint32_t d1, auto nc) {
                constexpr int NC = decltype(nc)::value;   // 16, 32 or 48 fp32 columns -> NC / 2 packed ones
                uint32_t v[NC];
                if constexpr (NC >= 32) tc5::tmem_ld32(d1, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
                if constexpr (NC % 32 == 16) tc5::tmem_ld16(d1 + (NC - 16), *reinterpret_cast<uint32_t(*)[16]>(&v[NC - 16]));
                tc5::tmem_wait_ld();
                RS_MARK(1);
                RS_EVT(302);
#pragma unroll
                for (int j = 0; j < NC / 2; ++j) v[j] = tc5::relu_pack_bf16x2(v[2 * j], v[2 * j + 1]);
                if constexpr (NC >= 32) tc5::tmem_st16(d1, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
                if constexpr (NC % 32 == 16) tc5::tmem_st8(d1 + (NC - 16) / 2, *reinterpret_cast<uint32_t(*)[8]>(&v[(NC - 16) / 2]));
                tc5::tmem_wait_st();
            };
            // ---- trunk rows -> X ring, by the same threads: warpgroup q < 3 copies plane q, thread = lane of the strip: one 16-byte
            //      cp.async per row and thread, LA rows ahead, issued in the slack AFTER the thread's E1 arrive.  A row is handed to the
            //      tensor core through that very arrive: before arriving for step s the thread has waited for its copy of row s+2
            //      (cp.async.wait_group -- issued LA - 2 steps earlier, so it never blocks in steady state) and executed
            //      fence.proxy.async (generic-proxy write -> async-proxy operand fetch); G1(s+2) is issued behind G2(s).  E3 reads the
            //      residual of row r after STEP_DONE(r+1), far down the same causality chain: no per-row barrier at all.
            //      What was tried first: a producer warp with cp.async.bulk (~110-140 clk PER COPY, 6..9 copies per cfg2 row), the
            //      same warp with 12 cp.async per lane (~480 clk per row), cp.async.mbarrier.arrive.noinc (blocks ~a memory latency):
            //      tools/rs_prod_bench.cu, profiles/r02_block_rs_timeline.md.
            constexpr int LA = 8;     // < NX - (E3's lag behind G1, ~4 steps)
            Steps itl(g0, g1, H);
            const uint8_t *src = nullptr;
            bool lfresh = true;
            auto load_row = [&](int r) {   // row r of this CTA's walk (itl is at row r); one commit group per call
                if (r < T) {
                    if (lfresh) {
                        const long long px = lane_pixel(itl.strip, row, N, H, W);
                        src = px >= 0 ? reinterpret_cast<const uint8_t *>(in) + (px + (long long)q * H * W) * 16 : nullptr;
                    }
                    const int slot = r % NX;
                    if (r >= NX) tc5::mbar_wait(bar(X_EMPTY + slot), ((r / NX) - 1) & 1);
                    RS_MARK(6);
                    if (src && itl.in_image()) cp_async16(xs + slot * XSLOT + q * XPLANE + row * 16, src + (long long)itl.y * W * 16, 16);
                    if (r + 1 < T) lfresh = itl.advance();
                }
                cp_async_commit();
            };
            if (q < 3) {
                tc5::pdl_wait();   // the previous kernel's trunk is complete and visible
#pragma unroll 1
                for (int r = 0; r < LA; ++r) load_row(r);
                cp_async_wait<LA - 3>();     // rows 0..2 have landed
                tc5::fence_proxy_async();
                tc5::mbar_arrive(bar(X_START));
            }
            RS_MARK(7);
            for (int s = 0; s < T; ++s) {
                const int eb = s & 1;
                RS_CRIT_WAIT(bar(D1_FULL + eb), (s >> 1) & 1);
                tc5::fence_after_sync();
                RS_MARK(0);
                RS_EVT(300);
                const uint32_t d1 = tmem + lane_base + d1_col(eb) + (q < 3 ? 32 * q : 96);
#ifdef B200SR_EXP_E1SHORT
                if (true) { uint32_t v8[8]; tc5::tmem_ld8(d1, v8); tc5::tmem_wait_ld(); if (v8[0] == 0x7fc12345u) tc5::tmem_st8(d1, v8); } else   // (timing experiment)
#endif
                if (ncol == 32) cvt(d1, std::integral_constant<int, 32>{});
                else if (ncol == 48) cvt(d1, std::integral_constant<int, 48>{});
                else if (ncol == 16) cvt(d1, std::integral_constant<int, 16>{});
                tc5::fence_before_sync();
                if (q < 3) tc5::mbar_arrive(bar(G2_READY + eb));           // release: publishes this thread's copy of row s+2 as well
                else tc5::mbar_arrive_relaxed(bar(G2_READY + eb));
                RS_MARK(2);
                RS_EVT(301);
#ifndef B200SR_EXP_NOLOAD   // (timing experiment: no trunk loads at all)
                if (q < 3) {
                    cp_async_wait<LA - 4>();     // row s+3 has landed: the next arrive hands it over
                    RS_MARK(4);
                    tc5::fence_proxy_async();    // BEFORE the new copy is issued: the fence waits for the thread's copies in flight (~700 clk behind a fresh one)
                    RS_MARK(5);
                    load_row(s + LA);
                }
#endif
                RS_MARK(3);
                RS_EVT(303);
            }
        } else if (wg <= 6) {
            // ---- E2 (WG5: even steps, WG6: odd steps): D2 + b2 -> bf16 t2 row in shared memory, zero outside the image
            const int e = (wg - 5) & 1;
            tc5::mbar_arrive(bar(G2_READY + e));  // stand-in for "the previous E2 on this buffer drained D2[e]"
            Steps it(g0, g1, H);
            bool fresh = true, lane_ok = false;
            if (e == 1 && T > 1) fresh = it.advance() || fresh;
            for (int s = e; s < T; s += 2) {
                if (fresh) lane_ok = lane_pixel(it.strip, row, N, H, W) >= 0;
                const bool ok = lane_ok && it.in_image();
                RS_MARK(7);
                tc5::mbar_wait(bar(D2_FULL + e), (s >> 1) & 1);
                tc5::fence_after_sync();
                RS_MARK(0);
                RS_EVT(400);
                uint32_t v[32];
                tc5::tmem_ld32(tmem + lane_base + d2_col(e), v);
                tc5::tmem_wait_ld();
                tc5::fence_before_sync();
                tc5::mbar_arrive(bar(G2_READY + e));  // D2[e] drained: counts towards G2 of step s+2
                RS_MARK(1);
                RS_EVT(401);
                uint4 c[NC2];
                uint32_t *cw = reinterpret_cast<uint32_t *>(c);
#pragma unroll
                for (int j4 = 0; j4 < 2 * NC2; ++j4) {
                    const float4 bb = *reinterpret_cast<const float4 *>(b2s + 4 * j4);  // broadcast read
                    cw[2 * j4] = ok ? pack_bf16x2(__uint_as_float(v[4 * j4]) + bb.x, __uint_as_float(v[4 * j4 + 1]) + bb.y) : 0u;
                    cw[2 * j4 + 1] = ok ? pack_bf16x2(__uint_as_float(v[4 * j4 + 2]) + bb.z, __uint_as_float(v[4 * j4 + 3]) + bb.w) : 0u;
                }
                const int b = s % NT;
                RS_MARK(2);
                if (s >= NT) tc5::mbar_wait(bar(STEP_DONE + b), ((s / NT) - 1) & 1);   // G3 of step s-5 has read this t2 slot
                RS_MARK(3);
                RS_EVT(402);
                uint8_t *dst = t2 + b * T2SLOT + (row + 1) * 16;
#ifdef B200SR_EXP_NOE2
                if (cw[0] == 0x7fc12345u)   // (timing experiment)
#endif
#pragma unroll
                for (int q = 0; q < (PACK ? 2 : NC2); ++q) *reinterpret_cast<uint4 *>(dst + q * T2PLANE) = c[q];
                if (PACK) {   // channels 16..19 twice: low half of the lane's own entry, high half of the entry on its left
                    const uint2 p4 = make_uint2(cw[8], cw[9]);
                    *reinterpret_cast<uint2 *>(dst + 2 * T2PLANE) = p4;
                    *reinterpret_cast<uint2 *>(dst + 2 * T2PLANE - 16 + 8) = p4;
                }
                tc5::fence_proxy_async();
                tc5::mbar_arrive(bar(G3_READY + b));
                RS_MARK(4);
                RS_EVT(403);
                if (s + 2 < T) {
                    fresh = it.advance();
                    fresh = it.advance() || fresh;
                }
            }
        } else {
            tc5::setmaxnreg_inc<72>();
            // ---- E3 (WG7): OUT slot + b3 + residual -> planar-8 trunk; the slot is re-zeroed for the row that uses it next
            {
                const uint32_t z[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
#pragma unroll
                for (int k = 0; k < NT * 4; ++k) tc5::tmem_st8(tmem + lane_base + out_col(0) + 8 * k, z);
                tc5::tmem_wait_st();
                tc5::fence_before_sync();
#pragma unroll
                for (int k = 0; k < NT - 1; ++k) tc5::mbar_arrive(bar(G3_READY + k));  // stand-ins: steps 0..3 find their new slot zeroed
            }
            Steps it(g0, g1, H);
            bool fresh = true;
            long long px = -1;
            for (int r = 0; r + 1 < T; ++r) {
                if (r > 0) fresh = it.advance();
                if (fresh) px = lane_pixel(it.strip, row, N, H, W);
                RS_MARK(7);
                tc5::mbar_wait(bar(STEP_DONE + (r + 1) % NT), ((r + 1) / NT) & 1);
                tc5::fence_after_sync();
                RS_MARK(0);
                RS_EVT(500);
                uint32_t v[32];
                const uint32_t oc = tmem + lane_base + out_col(r % NT);
                tc5::tmem_ld32(oc, v);
                const int xslot = r % NX;
                // (the copied row is visible: loader arrive -> G2_READY -> issuer -> commit -> STEP_DONE, acquired above)
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
                tc5::mbar_arrive_relaxed(bar(G3_READY + (r + NT - 1) % NT));   // slot r % 5 is zero again: step r+4 may start row r+5 in it
                tc5::mbar_arrive_relaxed(bar(X_EMPTY + xslot));                 // residual values are in registers
                RS_MARK(1);
                RS_EVT(501);
#ifdef B200SR_EXP_NOSTORE
                if (it.stored() && px >= 0 && row >= 1 && row <= SPAN && v[0] == 0x7fc12345u) {   // (timing experiment)
#else
                if (it.stored() && px >= 0 && row >= 1 && row <= SPAN) {
#endif
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
                RS_MARK(2);
                RS_EVT(502);
            }
        }
    }
    RS_FLUSH();
    RS_TM_FLUSH();
    tc5::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc5::tmem_free(tmem, TMEM_COLS);
}

}  // namespace b200sr
