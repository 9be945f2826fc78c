// wdsr_tc5q.cuh -- tile form of the tcgen05 fused WDSR-B residual block with the expand accumulator DECOUPLED from the reduce operand
// (derived from wdsr_tc5p.cuh: same tiles, same operand images, same E2 / G3 / E3; what follows lists only what differs).
//
// wdsr_tc5p.cuh keeps two 144-column fp32 expand accumulators D1[2] and writes relu(t1) back IN PLACE as the bf16 A operand of G2: a buffer
// is busy from G1's issue until G2 has read it -- G1 + commit -> E1 wake + tcgen05.ld + cvt + tcgen05.st + arrive -> issuer wake + G2,
// ~2 k clk with the queueing behind the 3x3 stream -- and two of them make one 128-pixel M-tile per ~1.08 k clk however the stages are
// tuned (the hypothesis this file tests; DESIGN.md 4.1b records that neither one nor two staging slots per half beat the in-place form).  Here:
//   * the expand is issued as two N-halves (channels 0..63 / 64..M1P-1) into ONE 64- and ONE 80-column fp32 staging area D1H[a], D1H[b];
//     a half is released as soon as its E1 warpgroup's tcgen05.ld has RETIRED (G1 + commit -> wake + ld + arrive -> issuer: ~0.75 k clk),
//   * relu(t1) goes to separate bf16 buffers T1[2] x 72 columns (the A operand of G2; free again when the G2 that read it has retired =
//     D2_FULL of that M-tile), so the long part of the old loop (st + hop + G2) no longer holds fp32 columns,
//   * G1 and G2 have an issuer warp each (warp 1 / warp 3): neither stream's barrier wait blocks the other.
// TMEM (512 columns): D1H[a] 0..63 | D1H[b] 64..143 | T1[2] x 72 at 144 | D2[2] x 32 at 288 | D3[4] x 32 at 352.
#pragma once
#include "wdsr_tc5p.cuh"

namespace b200sr {
namespace tc5v4 {
using namespace tc5cfg;
constexpr int NTHREADS = 768;
constexpr int TMEM_COLS = 512;
constexpr int XS_PLANE = NMT * 128 * 16;       // 10,240 B: 640 pixels x 16 B
constexpr int XS_NBUF = 3;                    // TMA runs two tiles ahead of the MMA stream
constexpr int XS_BUF = 3 * XS_PLANE;           // 30,720 B of tile data per buffer; the constant-one plane is shared
constexpr int XS_ONE = XS_NBUF * XS_BUF;       // byte offset of the constant-one plane
constexpr int XS_BYTES_ALL = XS_ONE + XS_PLANE;
constexpr int TMA_BYTES = 3 * HP * 16;         // 29,376 B per tile
// TMEM column map.  NS = staging slots per expand half:
//   NS = 1 (any M1P <= 144):  D1H[a] 0..63 | D1H[b] 64..143 | T1[2] x 72 at 144 | D2[2] x 32 at 288 | D3[4] x 32 at 352
//   NS = 2 (M1P <= 128, the pruned widths): D1H[slot] x M1P (a at +0, b at +64) | T1[2] x M1P/2 | D2[2] x 32 | D3[2] x 32  = 3 M1P + 128 <= 512:
//           two M-tiles of each half in flight through the short G1 -> ld -> release trip, paid for with two of the four 3x3 accumulators
template <int NS> struct Cols {
    int m1p;
    __device__ int d1h(int half, int slot) const { return NS == 1 ? half * 64 : slot * m1p + half * 64; }   // fp32 staging: a = channels 0..63, b = 64..M1P-1
    __device__ int t1(int e) const { return NS == 1 ? 144 + e * 72 : 2 * m1p + e * (m1p / 2); }            // packed bf16 relu(t1): K step j of G2 at column 8 j
    __device__ int d2(int e) const { return (NS == 1 ? 288 : 3 * m1p) + e * 32; }
    __device__ int d3(int k) const { return NS == 1 ? 352 + k * 32 : 3 * m1p + 64 + (k & 1) * 32; }
};
// Everything a group of MMAs needs is folded into ONE barrier per issuer step:
//   D1H_FREE[half] (128) = the E1 warpgroup of that half has read D1H[half] of the current M-tile: the next G1 half may be issued
//   G2_READY[e] (384)    = both E1 warpgroups wrote their K steps of T1[e]  +  E2 of the previous M-tile on this buffer drained D2[e]
//   G3_READY[k] (384)    = E2 of M-tiles k and k+1 wrote their t2 rows  +  E3 of the previous tile drained D3[k]
//   D2_FULL[e] (1)       = commit after G2: D2[e] is complete AND T1[e] has been read (E1 waits for it before overwriting T1[e])
enum Bar { XS_FULL = 0 /*3*/, XS_EMPTY = 3 /*3*/, G2_READY = 6, D2_FULL = 8, G3_READY = 10 /*4*/, T2R_FREE = 14 /*4*/, D3_FULL = 18 /*4*/,
           D1H_FULL = 22 /*half * 2 + slot*/, D1H_FREE = 26 /*half * 2 + slot*/, NBARS = 30 };
constexpr int CTRL_BYTES = 256;  // 30 mbarriers (240 B) + tmem base pointer at byte 240
constexpr size_t smem_bytes(int M1P) { return (size_t)tc5v4::CTRL_BYTES + XS_BYTES_ALL + T2_BYTES + (size_t)BlockTc5Layout(M1P).total; }
}  // namespace tc5v4

// NC2 = 8-channel chunks of t2 the block really has (3 dense; 2 / 1 for pruned M2 <= 16 / <= 8): the host packs only those
// (tap, chunk) slices of w3 (b200sr.cu) and the 3x3 issues 14 / 9 / 5 MMAs.  A template parameter, not a run-time value: the
// dense instantiation is then exactly the code that was tuned (a run-time switch cost it 3-5 %).
template <int NC2, int NS>
__global__ void __launch_bounds__(tc5v4::NTHREADS, 1)
wdsr_block_tc5q_kernel(const __grid_constant__ CUtensorMap tmap_in, const bf16 *__restrict__ in, bf16 *__restrict__ out,
                       const uint8_t *__restrict__ wimg, int M1P, int N, int H, int W, int tiles_x, int tiles_y, int ntiles) {
    using namespace tc5v4;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const BlockTc5Layout L(M1P);
    const tc5v4::Cols<NS> col{M1P};
    uint8_t *ctrl = smem_raw;
    uint8_t *xs = smem_raw + tc5v4::CTRL_BYTES;  // XS_NBUF x XS_BUF + constant-one plane
    uint8_t *t2 = xs + XS_BYTES_ALL;      // T2_BYTES
    uint8_t *wsm = t2 + T2_BYTES;         // L.total
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bars = smem_u32(ctrl);
    auto bar = [&](int b) { return bars + 8u * (uint32_t)b; };
    const uint32_t xs_u = smem_u32(xs), t2_u = smem_u32(t2), w_u = smem_u32(wsm);
    V3_DECL();
#ifdef B200SR_TC5_PROF
    const long long kstart__ = clock64();
    if (threadIdx.x == 0) { g_tc5p_cta[blockIdx.x][0] = gtimer__(); g_tc5p_cta[blockIdx.x][2] = smid__(); }
#endif

    const int nmine = (int)blockIdx.x < ntiles ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    auto tile_origin = [&](int it, int &x0, int &y0, int &n) {
        const int tile = blockIdx.x + it * gridDim.x;
        x0 = (tile % tiles_x) * TW - 1;
        y0 = ((tile / tiles_x) % tiles_y) * TH - 1;
        n = tile / (tiles_x * tiles_y);
    };
    auto tma_tile = [&](int it) {  // one lane: the three plane loads of tile iteration `it` into XS[it % XS_NBUF]
        int x0, y0, n;
        tile_origin(it, x0, y0, n);
        const int xb = it % XS_NBUF;
        tc5::mbar_arrive_expect_tx(bar(XS_FULL + xb), TMA_BYTES);
#pragma unroll
        for (int c = 0; c < 3; ++c) tc5::tma_load_plane(xs_u + xb * XS_BUF + c * XS_PLANE, &tmap_in, bar(XS_FULL + xb), x0, y0, c, n);
    };
    const int npre = nmine < XS_NBUF ? nmine : XS_NBUF;  // tiles whose loads are issued from the prologue

    // ---- one-time setup.  Launched with programmatic stream serialization: everything up to griddepcontrol.wait (barrier
    //      init, TMEM allocation, the weight image -- a constant --, shared-memory constants) may overlap the previous block
    //      kernel's tail; only the trunk loads (and, through them, every store) depend on it.
    tc5::pdl_launch_dependents();
    if (tid == 0) {
        for (int b = 0; b < XS_NBUF; ++b) {
            tc5::mbar_init(bar(XS_FULL + b), 1);
            tc5::mbar_init(bar(XS_EMPTY + b), 129);  // commit after the last G1 + the 128 threads of WG5 after the tile's last E3
        }
        for (int e = 0; e < 2; ++e) {
            for (int sl = 0; sl < 2; ++sl) {
                tc5::mbar_init(bar(D1H_FULL + 2 * e + sl), 1);
                tc5::mbar_init(bar(D1H_FREE + 2 * e + sl), 128);
            }
            tc5::mbar_init(bar(G2_READY + e), 384);
            tc5::mbar_init(bar(D2_FULL + e), 1);
        }
        for (int k = 0; k < 4; ++k) {
            tc5::mbar_init(bar(G3_READY + k), 384);
            tc5::mbar_init(bar(T2R_FREE + k), 1);
            tc5::mbar_init(bar(D3_FULL + k), 1);
        }
        tc5::mbar_init_fence();
        tc5::tma_prefetch_desc(&tmap_in);
        tc5::pdl_wait();                                   // the previous kernel's trunk is complete and visible
        for (int it = 0; it < npre; ++it) tma_tile(it);    // first loads in flight while the rest of the CTA sets up
    }
    __syncwarp();
    if (warp == 0) tc5::tmem_alloc(smem_u32(ctrl + 240), tc5v4::TMEM_COLS);
    for (int i = tid; i < L.total / 16; i += NTHREADS) cp_async16(wsm + i * 16, wimg + i * 16, 16);
    cp_async_commit();
    // shared-memory constants; never touch bytes a TMA box lands on (pixel rows 0..611 of the tile planes): the loads are in flight
    for (int i = tid; i < 3 * XS_NBUF * (NMT * 128 - HP); i += NTHREADS) {  // pad pixel rows 612..639 of every tile plane stay zero
        const int pl = i / (NMT * 128 - HP), r = HP + i % (NMT * 128 - HP);
        *reinterpret_cast<uint4 *>(xs + pl * XS_PLANE + r * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
    for (int i = tid; i < XS_PLANE / 16; i += NTHREADS)                       // constant-one plane: 1.0 in channels 0,1
        *reinterpret_cast<uint4 *>(xs + XS_ONE + i * 16) = make_uint4(0x3F803F80u, 0u, 0u, 0u);
    for (int i = tid; i < 16; i += NTHREADS) *reinterpret_cast<uint4 *>(t2 + 3 * T2_COPY + i * 16) = make_uint4(0u, 0u, 0u, 0u);
    if (NC2 < 3)   // a pruned block never writes the absent chunks, and the zero-weight dummy half of its last 3x3 instruction reads
                   // one chunk past the last slice: t2 must start as zeros (the dense block pays nothing)
        for (int i = tid; i < T2_BYTES / 16; i += NTHREADS) *reinterpret_cast<uint4 *>(t2 + i * 16) = make_uint4(0u, 0u, 0u, 0u);
    cp_async_wait<0>();
    tc5::fence_proxy_async();
    tc5::fence_before_sync();
    __syncthreads();
    tc5::fence_after_sync();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t *>(ctrl + 240);
#ifdef B200SR_TC5_PROF
    prof__[6] = (unsigned long long)(clock64() - kstart__);
#endif

    const int wg = warp >> 2;  // warpgroup 0..5
    // (each setmaxnreg sits at the top of the branch it governs, so that ptxas sees it dominate that role's code)
    if (wg == 0) {
#ifdef B200SR_TC5_PROF
      tc5::setmaxnreg_dec<56>();   // the probe counters need registers; paid for by the E2 warpgroups (see below)
#else
      tc5::setmaxnreg_dec<48>();   // (one more issuer warp / per-layer descriptors: 40 registers spilled, and a spill in an issuer thread stalls the MMA stream)
#endif
      if (warp == 0) {
        // ============================== TMA producer ==============================
        if (lane == 0) {   // (the lane that issued the prologue loads)
            for (int it = npre; it < nmine; ++it) {
                tc5::mbar_wait(bar(XS_EMPTY + it % XS_NBUF), ((it / XS_NBUF) & 1) ^ 1);
                tma_tile(it);
            }
        }
        __syncwarp();
      } else if (warp == 1) {
        // ============================== MMA issuer A: the expand halves (G1) ==============================
        const bool leader = tc5::elect_one();
        const int NA = M1P < 64 ? M1P : 64, NB = M1P - NA;
        const uint32_t idesc_a = tc5::idesc_bf16_f32(128, NA), idesc_b = tc5::idesc_bf16_f32(128, NB > 0 ? NB : 16);
        // B operand rows = expand channels, 8-channel groups 512 B apart: half b starts 8 groups in
        const uint64_t bw1a = tc5::smem_desc(w_u + L.w1, 128, 512), bw1b = tc5::smem_desc(w_u + L.w1 + 256, 128, 512);
        const uint64_t half_b = (uint64_t)((8 * 512) >> 4);
        const uint64_t ax0 = tc5::smem_desc(xs_u, XS_PLANE, 128);  // planes paired through LBO
        auto issue_g1 = [&](int xb, int m, int half, int slot) {  // leader only
            const uint32_t off = xb * XS_BUF + m * 2048;
            const uint32_t d = tmem + col.d1h(half, slot);
            const uint64_t hb = half ? half_b : 0;
            const uint32_t idesc = half ? idesc_b : idesc_a;
            tc5::mma_ss(d, ax0 + (uint64_t)(off >> 4), bw1a + hb, idesc, false);  // planes 0,1
            // plane 2 paired with the shared constant-one plane: LBO = their distance
            tc5::mma_ss(d, tc5::smem_desc(xs_u + off + 2 * XS_PLANE, XS_ONE - xb * XS_BUF - 2 * XS_PLANE, 128), bw1b + hb, idesc, true);
            tc5::commit(bar(D1H_FULL + 2 * half + slot));
        };
        V3_T0();
        uint32_t g = 0;   // M-tiles issued so far: D1H_FREE[half] completes once per M-tile
        for (int it = 0; it < nmine; ++it) {
            const int xb = it % XS_NBUF;
            V3_WAIT(2, bar(XS_FULL + xb), (it / XS_NBUF) & 1);
            for (int m = 0; m < NMT; ++m, ++g) {
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    if (half == 1 && NB == 0) break;
                    // slot g % NS was last used by M-tile g - NS: its E1 warpgroup must have read it (phase (g - NS) / NS of that slot)
                    const int slot = NS == 1 ? 0 : (int)(g & 1);
                    if (g >= (uint32_t)NS) V3_WAIT(0, bar(D1H_FREE + 2 * half + slot), ((g - NS) / NS) & 1);
                    tc5::fence_after_sync();
                    V3_EVT(100 + m);
                    if (leader) {
                        issue_g1(xb, m, half, slot);
                        if (m == NMT - 1 && (half == 1 || NB == 0)) tc5::commit(bar(XS_EMPTY + xb));  // all G1 reads of XS[xb] retired
                    }
                    __syncwarp();
                }
            }
        }
        V3_ADD(5);
        if (nmine > 0) tc5::mbar_wait(bar(D1H_FULL + 2 * (NB > 0 ? 1 : 0) + (NS == 1 ? 0 : (int)((g - 1) & 1))), ((g - 1) / NS) & 1);  // the last G1 and all before it retired
      } else if (warp == 3) {
        // ============================== MMA issuer C: the reduce (G2) ==============================
        const bool leader = tc5::elect_one();
        const uint32_t idesc32 = tc5::idesc_bf16_f32(128, 32);
        const uint64_t bw2 = tc5::smem_desc(w_u + L.w2, 128, L.sbo2);
        const int nk2 = M1P / 16;
        uint32_t n_g2[2] = {0, 0};
        for (int it = 0; it < nmine; ++it)
            for (int m = 0; m < NMT; ++m) {
                const int e = m & 1;
                tc5::mbar_wait(bar(G2_READY + e), n_g2[e] & 1);
                ++n_g2[e];
                tc5::fence_after_sync();
                V3_EVT(110 + m);
                if (leader) {
                    const uint32_t d2 = tmem + col.d2(e), a2 = tmem + col.t1(e);
                    tc5::mma_ts(d2, a2, bw2, idesc32, false);
#pragma unroll 4
                    for (int j = 1; j < nk2; ++j) tc5::mma_ts(d2, a2 + 8 * j, bw2 + (uint64_t)(16 * j), idesc32, true);
                    tc5::commit(bar(D2_FULL + e));
                }
                __syncwarp();
            }
        if (nmine > 0) tc5::mbar_wait(bar(D2_FULL + 0), (n_g2[0] - 1) & 1);  // the last G2 (M-tile 4, buffer 0) and all before it retired
      } else if (warp == 2) {
        // ============================== MMA issuer B: the 3x3 stream ==============================
        const bool leader = tc5::elect_one();
        const uint32_t idesc32 = tc5::idesc_bf16_f32(128, 32);
        const uint64_t bw3 = tc5::smem_desc(w_u + L.w3, 128, 28 * 128);
        const uint64_t at0 = tc5::smem_desc(t2_u, 0, T2_GROUP);    // LBO added per instruction
        auto issue_g3_nc = [&](int k, auto ncc) {  // leader only; NC = chunks of t2 (3 dense; 2 or 1 for pruned M2 <= 16 / <= 8)
            constexpr int NC = decltype(ncc)::value, NS = 9 * NC, NM = (NS + 1) / 2;
            const uint64_t abase = at0 + (uint64_t)((k * 4 * T2_ROW) >> 4);
            const uint32_t d3 = tmem + col.d3(k);
#ifdef B200SR_EXP_G3SHORT
            constexpr int NG3 = 2;   // (timing experiment: results are wrong)
#else
            constexpr int NG3 = NM;
#endif
#pragma unroll
            for (int i = 0; i < NG3; ++i) {   // slice q = (dx * 3 + dy) * NC + chunk, two slices per K = 16 instruction through LBO
                const int q0 = 2 * i, q1 = 2 * i + 1;
                const int a0 = (q0 / (3 * NC)) * T2_COPY + ((q0 / NC) % 3) * T2_ROW + (q0 % NC) * 128;
                const int a1 = q1 < NS ? (q1 / (3 * NC)) * T2_COPY + ((q1 / NC) % 3) * T2_ROW + (q1 % NC) * 128 : a0 + 128;
                tc5::mma_ss(d3, abase + (uint64_t)(a0 >> 4) + ((uint64_t)((a1 - a0) >> 4) << 16), bw3 + (uint64_t)(16 * i), idesc32,
                            i > 0);
            }
            tc5::commit(bar(D3_FULL + k));
            if (k == 3) tc5::commit(bar(T2R_FREE + 3));  // (in-order) every 3x3 MMA of this tile has retired
        };
        auto issue_g3 = [&](int k) { issue_g3_nc(k, std::integral_constant<int, NC2>{}); };
        // wait (whole warp) for what G3(k) of tile `t` needs, then issue it
        auto do_g3 = [&](int t, int k) {
            V3_WAIT(3, bar(G3_READY + k), t & 1);
            tc5::fence_after_sync();
            V3_EVT(200 + k);
            if (leader) issue_g3(k);
            __syncwarp();
            V3_EVT(210 + k);
        };
        V3_T0();
        for (int it = 0; it < nmine; ++it)
#pragma unroll 1   // (unrolled over k, the descriptor temporaries of the shorter pruned batches spill at this warpgroup's 40 registers)
            for (int k = 0; k < 4; ++k) do_g3(it, k);
        V3_ADD(5);
        if (nmine > 0) tc5::mbar_wait(bar(T2R_FREE + 3), (nmine - 1) & 1);  // every G3 of this CTA has retired
      }
    } else {
        // ============================== epilogue warpgroups ==============================
        const int e = (wg - 3) & 1;              // WG3 / WG4: M-tile parity / buffer index this warpgroup serves
        const int row = (warp & 3) * 32 + lane;  // row of the M-tile == TMEM lane
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const float *b2s = reinterpret_cast<const float *>(wsm + L.b2);
        const float *b3s = reinterpret_cast<const float *>(wsm + L.b3);

        // ---- E1 of one expand half: fp32 staging -> registers (the half is released as soon as the loads have retired) -> relu -> bf16x2 ->
        //      this half's K steps of T1[eb].  WG1: channels 0..63 -> T1 columns 0..31; WG2: channels 64..M1P-1 -> columns 32...
        //      N32 / R16: 32-column loads and a 16-column remainder (compile-time shapes: the arrays stay in registers).
        auto e1_half = [&](int eb, int nth, int half, int slot, auto n32c, auto r16c) {
            constexpr int N32 = decltype(n32c)::value;
            constexpr bool R16 = decltype(r16c)::value != 0;
            const uint32_t d1 = tmem + lane_base + col.d1h(half, slot);
            const uint32_t t1 = tmem + lane_base + col.t1(eb) + (half ? 32 : 0);
            uint32_t va[32], vb[32], vc[16];
            if constexpr (N32 >= 1) tc5::tmem_ld32(d1, va);
            if constexpr (N32 >= 2) tc5::tmem_ld32(d1 + 32, vb);
            if constexpr (R16) tc5::tmem_ld16(d1 + 32 * N32, vc);
            tc5::tmem_wait_ld();
            tc5::fence_before_sync();
            tc5::mbar_arrive_relaxed(bar(D1H_FREE + 2 * half + slot));   // the staging slot may take a later M-tile's half
            V3_EVT(302);
            if constexpr (N32 >= 1) {
#pragma unroll
                for (int j = 0; j < 16; ++j) va[j] = tc5::relu_pack_bf16x2(va[2 * j], va[2 * j + 1]);
            }
            if constexpr (N32 >= 2) {
#pragma unroll
                for (int j = 0; j < 16; ++j) vb[j] = tc5::relu_pack_bf16x2(vb[2 * j], vb[2 * j + 1]);
            }
            if constexpr (R16) {
#pragma unroll
                for (int j = 0; j < 8; ++j) vc[j] = tc5::relu_pack_bf16x2(vc[2 * j], vc[2 * j + 1]);
            }
            // T1[eb] was the A operand of the G2 two M-tiles ago on this buffer: its commit (D2_FULL) says it has been read
            if (nth > 0) {
                tc5::mbar_wait(bar(D2_FULL + eb), (nth - 1) & 1);
                tc5::fence_after_sync();
            }
            if constexpr (N32 >= 1) tc5::tmem_st16(t1, *reinterpret_cast<uint32_t(*)[16]>(&va[0]));
            if constexpr (N32 >= 2) tc5::tmem_st16(t1 + 16, *reinterpret_cast<uint32_t(*)[16]>(&vb[0]));
            if constexpr (R16) tc5::tmem_st8(t1 + 16 * N32, *reinterpret_cast<uint32_t(*)[8]>(&vc[0]));
        };
        auto e1 = [&](int eb, int nth, int slot) {
            using std::integral_constant;
            tc5::fence_after_sync();
            V3_T0();
            V3_EVT(300);
            const int half = wg - 1;
            const int ncol = half == 0 ? (M1P < 64 ? M1P : 64) : (M1P > 64 ? M1P - 64 : 0);
            switch (ncol) {
                case 80: e1_half(eb, nth, half, slot, integral_constant<int, 2>{}, integral_constant<int, 1>{}); break;
                case 64: e1_half(eb, nth, half, slot, integral_constant<int, 2>{}, integral_constant<int, 0>{}); break;
                case 48: e1_half(eb, nth, half, slot, integral_constant<int, 1>{}, integral_constant<int, 1>{}); break;
                case 32: e1_half(eb, nth, half, slot, integral_constant<int, 1>{}, integral_constant<int, 0>{}); break;
                case 16: e1_half(eb, nth, half, slot, integral_constant<int, 0>{}, integral_constant<int, 1>{}); break;
                default:   // M1P <= 64: nothing in the upper half -- keep in step with the buffer's phases all the same
                    if (nth > 0) tc5::mbar_wait(bar(D2_FULL + eb), (nth - 1) & 1);
                    break;
            }
            tc5::tmem_wait_st();
            tc5::fence_before_sync();
            tc5::mbar_arrive_relaxed(bar(G2_READY + eb));  // this half's K steps of T1[eb] complete (wait::st); no release: E3's output stores may be in flight
            V3_ADD(5);
            V3_EVT(301);
        };
        // ---- E3: D3 + b3 + residual -> bf16 NHWC (3x3 M-tile k of tile iteration t)
        auto e3 = [&](int t, int k, int x0, int y0, int n) {
            const int xb = t % XS_NBUF;
            tc5::fence_after_sync();
            V3_T0();
            V3_EVT(500 + k);
            uint32_t v[32];
            tc5::tmem_ld32(tmem + lane_base + col.d3(k), v);
            const int ly = 4 * k + (row >> 5), lx = row & 31;
            const int gy = y0 + 1 + ly, gx = x0 + 1 + lx;
            const uint8_t *res = xs + xb * XS_BUF + ((ly + 1) * HW_ + lx + 1) * 16;
            uint4 rv[3];
#pragma unroll
            for (int q = 0; q < 3; ++q) rv[q] = *reinterpret_cast<const uint4 *>(res + q * XS_PLANE);
            V3_EVT(520 + k);
            tc5::tmem_wait_ld();
            V3_EVT(530 + k);
            tc5::fence_before_sync();
            // D3 buffer drained: counts towards the G3 that uses it next (NS = 1: G3(k) of the next tile; NS = 2, two buffers: G3(k + 2) of this
            // tile for k < 2, G3(k - 2) of the next one otherwise)
            tc5::mbar_arrive_relaxed(bar(G3_READY + (NS == 1 ? k : ((k + 2) & 3))));
            V3_EVT(540 + k);  // D3[k] drained (wait::ld): counts towards the next tile's G3(k)
#ifdef B200SR_EXP_NOSTORE
            if (gy < H && gx < W && v[0] == 0x7fc12345u) {   // (timing experiment)
#else
            if (gy < H && gx < W) {
#endif
                bf16 *o = out + (((long long)n * 3 * H + gy) * W + gx) * 8;   // planar-8 trunk: plane q is H*W*8 elements further
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
                    *reinterpret_cast<uint4 *>(o + (long long)q * H * W * 8) = ov;   // 32 lanes = 512 contiguous bytes
                }
            }
            V3_ADD(7);
            V3_EVT(510 + k);
        };
        // ---- E2: D2 + b2 -> bf16 -> three x-shifted copies of t2 (zero outside the image)
        auto e2 = [&](int m, uint32_t par, int x0, int y0) {
            V3_WAIT(1, bar(D2_FULL + e), par);
            tc5::fence_after_sync();
            V3_T0();
            V3_EVT(400 + m);
            uint32_t v[32];
            tc5::tmem_ld32(tmem + lane_base + col.d2(e), v);
            const int p = m * 128 + row;
            const int r = p / HW_, hx = p - r * HW_;
            const int gy = y0 + r, gx = x0 + hx;
            const bool ok = gy >= 0 && gy < H && gx >= 0 && gx < W;
            tc5::tmem_wait_ld();
            tc5::fence_before_sync();
            tc5::mbar_arrive(bar(G2_READY + e));  // D2[e] drained: counts towards the NEXT G2 on this buffer
#ifdef B200SR_EXP_NOE2
            if (false) {
#else
            if (p < HP) {
#endif
                uint4 c[NC2];   // only the chunks this block has (compile-time sized: a partly used array went to local memory)
                uint32_t *cw = reinterpret_cast<uint32_t *>(c);
#pragma unroll
                for (int j4 = 0; j4 < 2 * NC2; ++j4) {
                    const float4 bb = *reinterpret_cast<const float4 *>(b2s + 4 * j4);  // broadcast read
                    cw[2 * j4] = ok ? pack_bf16x2(__uint_as_float(v[4 * j4]) + bb.x, __uint_as_float(v[4 * j4 + 1]) + bb.y) : 0u;
                    cw[2 * j4 + 1] = ok ? pack_bf16x2(__uint_as_float(v[4 * j4 + 2]) + bb.z, __uint_as_float(v[4 * j4 + 3]) + bb.w) : 0u;
                }
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    const int xi = hx - d;
                    if (xi >= 0 && xi < TW) {
                        uint8_t *dst = t2 + d * T2_COPY + r * T2_ROW + (xi >> 3) * T2_GROUP + (xi & 7) * 16;
#pragma unroll
                        for (int q = 0; q < NC2; ++q) *reinterpret_cast<uint4 *>(dst + q * 128) = c[q];
                    }
                }
            }
            tc5::fence_proxy_async();
            if (m >= 1) tc5::mbar_arrive(bar(G3_READY + m - 1));
            if (m <= 3) tc5::mbar_arrive(bar(G3_READY + m));
            V3_ADD(6);
            V3_EVT(410 + m);
        };

        if (wg <= 2) {
            tc5::setmaxnreg_inc<104>();
            // WG1 / WG2: their expand half of every M-tile, in issue order.  M-tile i % 5 of tile i / 5 uses T1 / D2 buffer (i % 5) & 1;
            // a buffer serves 3 (even) or 2 (odd) M-tiles per tile; the staging areas D1H serve every M-tile (phase = i).
            const int hbar = (wg == 2 && M1P > 64) ? 1 : 0;   // (an absent upper half paces itself on the lower half's barrier)
            for (int i1 = 0; i1 < NMT * nmine; ++i1) {
                const int mm = i1 % NMT, eb = mm & 1, nth = (i1 / NMT) * (eb == 0 ? 3 : 2) + (mm >> 1);
                const int slot = NS == 1 ? 0 : (i1 & 1);
                V3_WAIT(0, bar(D1H_FULL + 2 * hbar + slot), (i1 / NS) & 1);
                e1(eb, nth, slot);
            }
        } else if (wg <= 4) {
#ifdef B200SR_TC5_PROF
            tc5::setmaxnreg_dec<72>();   // keep the sum of the warpgroup budgets at the 80 x 768 the CTA was launched with
#endif
            // WG3 / WG4: E2 of M-tiles m = e, e+2, ..
            tc5::mbar_arrive(bar(G2_READY + e));  // stand-in for "previous E2 drained D2[e]"
            uint32_t n_d2 = 0;
            for (int it = 0; it < nmine; ++it) {
                int x0, y0, n;
                tile_origin(it, x0, y0, n);
                for (int m = e; m < NMT; m += 2) {
                    // t2 rows may be overwritten once the previous tile's 3x3 MMAs that READ them have retired.  M-tile m covers halo
                    // rows ~3.8m .. 3.8m+3.8 and G3(k) reads halo rows 4k .. 4k+5, so the last reader is G3(min(m, 3)); commits retire
                    // in order, hence D3_FULL[min(m,3)] of tile it-1 (its phase `it` cannot complete before this very E2 has run, so
                    // the parity wait is unambiguous).
                    if (it > 0) V3_WAIT(2, bar(D3_FULL + (m < 3 ? m : 3)), (it - 1) & 1);
                    e2(m, n_d2 & 1, x0, y0);
                    ++n_d2;
                }
            }
        } else {
            tc5::setmaxnreg_dec<64>();
            // WG5: E3 of 3x3 M-tiles k = 0..3 of every tile
#pragma unroll
            for (int k = 0; k < (NS == 1 ? 4 : 2); ++k) tc5::mbar_arrive(bar(G3_READY + k));  // stand-ins: the first G3s find their D3 buffer free
            for (int it = 0; it < nmine; ++it) {
                int x0, y0, n;
                tile_origin(it, x0, y0, n);
                for (int k = 0; k < 4; ++k) {
                    V3_WAIT(3, bar(D3_FULL + k), it & 1);
                    if (k == 0) V3_WAIT(4, bar(XS_FULL + (it % XS_NBUF)), (it / XS_NBUF) & 1);  // acquire the TMA-written tile (residual)
                    e3(it, k, x0, y0, n);
                }
                tc5::mbar_arrive_relaxed(bar(XS_EMPTY + (it % XS_NBUF)));  // all four E3 done: the residual values were consumed
            }
        }
    }
#ifdef B200SR_TC5_PROF
    if (threadIdx.x == 0) g_tc5p_cta[blockIdx.x][1] = gtimer__();
    if (blockIdx.x == 0 && threadIdx.x == 0) { g_tc5p_prof[56] = (unsigned long long)(clock64() - kstart__); }
    if (blockIdx.x == 0 && threadIdx.x == 32) { g_tc5p_prof[57] = prof__[6]; }
#endif
#ifdef B200SR_TC5_PROF
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) g_tc5p_evtn[threadIdx.x >> 5] = evn__;
#endif
    if (warp == 1) V3_FLUSH(0);
    if (warp == 2) V3_FLUSH(40);
    if (warp == 4) V3_FLUSH(8);
    if (warp == 16) V3_FLUSH(24);
    if (warp == 20) V3_FLUSH(32);
    tc5::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc5::tmem_free(tmem, tc5v4::TMEM_COLS);
}

}  // namespace b200sr
