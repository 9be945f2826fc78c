// wdsr_tc5p.cuh -- production form of the tcgen05 fused WDSR-B residual block (see wdsr_tc5.cuh for the algebra and the
// sequential reference form that validated the descriptor / TMEM protocol).
//
//   Six warpgroups (768 threads); EVERY warp has exactly one in-order work queue and only ever blocks on the one mbarrier it
//   needs next (a failed probe of a pending mbarrier and a polling loop both cost hundreds of cycles of hand-off latency):
//   WG0  warp 0   TMA producer   3 x cp.async.bulk.tensor.5d per tile (one per 8-channel plane) into a triple-buffered,
//                                chunk-planar trunk tile; out-of-image halo is zero-filled by the TMA unit
//        warp 1   MMA issuer A   one elected lane issues the G1 / G2 stream (software-pipelined over M-tiles)
//        warp 2   MMA issuer B   one elected lane issues the 3x3 (G3) stream; two issuers so that neither stream's barrier
//                                wait blocks the other -- the tensor pipe interleaves them
//        warp 3   idle
//   WG1  warps 4-7    E1 (relu -> bf16 A operand), expand channels 0..63 of EVERY M-tile
//   WG2  warps 8-11   E1, expand channels 64..143 of every M-tile
//               (E1 sits on the G1 -> G2 critical loop: two warpgroups split its columns and do nothing else;
//                WG1 packs into columns 0..31 of D1, WG2 into columns 104..143 -- both in place inside their own half)
//   WG3  warps 12-15  E2 (t2 -> three shifted copies in shared memory) of even M-tiles
//   WG4  warps 16-19  E2 of odd M-tiles
//   WG5  warps 20-23  E3 (bias + residual + store) of every 3x3 M-tile -- the only warps with global stores in flight
//   Registers are re-balanced with setmaxnreg: WG0 40, E1 104, E2 80, E3 72 per thread; the budgets x 128 threads must sum to at
//   most the 80 x 768 registers the CTA is launched with, or setmaxnreg.inc waits forever.
// (a warp can only touch TMEM lanes 32*(warp%4)..+31, so every warpgroup is 4 consecutive warps.)
//
// Shared-memory operand layout (SWIZZLE_NONE, K-major): chunk-planar  XS[plane c][pixel p][16 B]  for the trunk tile
// (planes 0..2 = channels 8c..8c+7, plane 3 = the constant-one channel that carries b1), so one TMA box lands as one
// contiguous plane and a K=16 MMA pairs two planes through LBO = plane stride, SBO = 128 B.
// TMEM (512 columns): D1[2] x 144 (relu(t1) is written back in place as the bf16 A operand of G2), D2[2] x 32, D3[4] x 32.
#pragma once
#include <cuda.h>

#include <type_traits>

#include "common.cuh"
#include "tc5.cuh"
#include "wdsr_tc5_layout.cuh"

#ifdef B200SR_TC5_PROF
__device__ unsigned long long g_tc5p_prof[64];
__device__ unsigned long long g_tc5p_cta[1024][3];
__device__ unsigned long long g_tc5p_evt[24][2048];   // per warp: (id << 48) | clock
__device__ int g_tc5p_evtn[24];
__device__ __forceinline__ unsigned long long gtimer__() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ unsigned smid__() { unsigned r; asm volatile("mov.u32 %0, %smid;" : "=r"(r)); return r; }
// timers accumulate in registers (prof__[slot & 7]); each warp's lane 0 of CTA 0 flushes them once at kernel end
#define V3_WAIT(slot, b, par) do { const long long w0__ = clock64(); tc5::mbar_wait(b, par); prof__[(slot) & 7] += (unsigned long long)(clock64() - w0__); } while (0)
#define V3_T0() const long long v3t0__ = clock64()
#define V3_ADD(slot) do { prof__[(slot) & 7] += (unsigned long long)(clock64() - v3t0__); } while (0)
#define V3_DECL() unsigned long long prof__[8] = {0, 0, 0, 0, 0, 0, 0, 0}; int evn__ = 0
#define V3_EVT(id) do { if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && evn__ < 2048) { g_tc5p_evt[threadIdx.x >> 5][evn__++] = ((unsigned long long)(id) << 48) | ((unsigned long long)clock64() & 0xFFFFFFFFFFFFull); } } while (0)
#define V3_FLUSH(base) do { if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) { for (int i__ = 0; i__ < 8; ++i__) g_tc5p_prof[(base) + i__] = prof__[i__]; } } while (0)
#else
#define V3_WAIT(slot, b, par) tc5::mbar_wait(b, par)
#define V3_T0() do {} while (0)
#define V3_ADD(slot) do {} while (0)
#define V3_DECL() do {} while (0)
#define V3_EVT(id) do {} while (0)
#define V3_FLUSH(base) do {} while (0)
#endif

namespace b200sr {
namespace tc5v3 {
using namespace tc5cfg;
constexpr int NTHREADS = 768;
constexpr int TMEM_COLS = 512;
constexpr int XS_PLANE = NMT * 128 * 16;       // 10,240 B: 640 pixels x 16 B
constexpr int XS_NBUF = 3;                    // TMA runs two tiles ahead of the MMA stream
constexpr int XS_BUF = 3 * XS_PLANE;           // 30,720 B of tile data per buffer; the constant-one plane is shared
constexpr int XS_ONE = XS_NBUF * XS_BUF;       // byte offset of the constant-one plane
constexpr int XS_BYTES_ALL = XS_ONE + XS_PLANE;
constexpr int TMA_BYTES = 3 * HP * 16;         // 29,376 B per tile
__host__ __device__ constexpr int d1_col(int e) { return e * 144; }
__host__ __device__ constexpr int d2_col(int e) { return 288 + e * 32; }
__host__ __device__ constexpr int d3_col(int k) { return 352 + k * 32; }
// The MMA issuer is one thread and an mbarrier wait costs it ~100 clk even when already complete, so everything a group of
// MMAs needs is folded into ONE barrier:
//   G2_READY[e] (384) = both column halves of E1 wrote A2 into D1[e]  +  E2 of the previous M-tile on this buffer drained D2[e]
//   G3_READY[k] (384) = E2 of M-tiles k and k+1 wrote their t2 rows  +  E3 of the previous tile drained D3[k]
// (arrivals are per thread: 32 lanes arriving in one instruction measured no slower than an elected lane after __syncwarp)
enum Bar { XS_FULL = 0 /*3*/, XS_EMPTY = 3 /*3*/, D1_FULL = 6, G2_READY = 8, D2_FULL = 10, G3_READY = 12 /*4*/, T2R_FREE = 16 /*4*/,
           D3_FULL = 20 /*4*/, NBARS = 24 };
constexpr int CTRL_BYTES = 256;  // 22 mbarriers (176 B) + tmem base pointer at byte 240
constexpr size_t smem_bytes(int M1P) { return (size_t)tc5v3::CTRL_BYTES + XS_BYTES_ALL + T2_BYTES + (size_t)BlockTc5Layout(M1P).total; }
}  // namespace tc5v3

// NC2 = 4: three chunks whose last one holds <= 4 channels (M2 = 17..20, the dense block), "packed": the two copies of that chunk hold
//          [pixel | right neighbour] x 4 channels per 16-byte entry, so two horizontal taps share one K = 8 half and the 3x3 is 24 slices =
//          12 MMAs per M-tile instead of 27 slices = 14 (the operand image must be the packed one: b200sr.cu, pack_w3_tc5).
// NC2 = 8-channel chunks of t2 the block really has (3 dense; 2 / 1 for pruned M2 <= 16 / <= 8): the host packs only those
// (tap, chunk) slices of w3 (b200sr.cu) and the 3x3 issues 14 / 9 / 5 MMAs.  A template parameter, not a run-time value: the
// dense instantiation is then exactly the code that was tuned (a run-time switch cost it 3-5 %).
template <int NC2>
__global__ void __launch_bounds__(tc5v3::NTHREADS, 1)
wdsr_block_tc5p_kernel(const __grid_constant__ CUtensorMap tmap_in, const bf16 *__restrict__ in, bf16 *__restrict__ out,
                       const uint8_t *__restrict__ wimg, int M1P, int N, int H, int W, int tiles_x, int tiles_y, int ntiles) {
    using namespace tc5v3;
    constexpr bool PK = NC2 == 4;            // packed last chunk
    constexpr int NCH = PK ? 3 : NC2;        // chunks E2 produces
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const BlockTc5Layout L(M1P);
    uint8_t *ctrl = smem_raw;
    uint8_t *xs = smem_raw + tc5v3::CTRL_BYTES;  // XS_NBUF x XS_BUF + constant-one plane
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
            tc5::mbar_init(bar(D1_FULL + e), 1);
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
    if (warp == 0) tc5::tmem_alloc(smem_u32(ctrl + 240), tc5v3::TMEM_COLS);
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
    if (NCH < 3)   // a pruned block never writes the absent chunks, and the zero-weight dummy half of its last 3x3 instruction reads
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
      tc5::setmaxnreg_dec<40>();
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
        // ============================== MMA issuer ==============================
        const bool leader = tc5::elect_one();
        const uint32_t idesc1 = tc5::idesc_bf16_f32(128, M1P), idesc32 = tc5::idesc_bf16_f32(128, 32);
        const uint64_t bw1a = tc5::smem_desc(w_u + L.w1, 128, 512), bw1b = tc5::smem_desc(w_u + L.w1 + 256, 128, 512);
        const uint64_t bw2 = tc5::smem_desc(w_u + L.w2, 128, L.sbo2);
        const uint64_t ax0 = tc5::smem_desc(xs_u, XS_PLANE, 128);  // planes paired through LBO
        const int nk2 = M1P / 16;
        const int a2hi = M1P - (M1P - 64) / 2;   // first column of WG-B's packed output (104 for M1P = 144)
        auto issue_g1 = [&](int xb, int m) {  // leader only
            const uint32_t off = xb * XS_BUF + m * 2048;
            const int e = m & 1;
            tc5::mma_ss(tmem + d1_col(e), ax0 + (uint64_t)(off >> 4), bw1a, idesc1, false);  // planes 0,1
            // plane 2 paired with the shared constant-one plane: LBO = their distance
            tc5::mma_ss(tmem + d1_col(e), tc5::smem_desc(xs_u + off + 2 * XS_PLANE, XS_ONE - xb * XS_BUF - 2 * XS_PLANE, 128), bw1b, idesc1,
                        true);
            tc5::commit(bar(D1_FULL + e));
        };
        uint32_t n_g2[2] = {0, 0};
        V3_T0();
        for (int it = 0; it < nmine; ++it) {
            const int xb = it % XS_NBUF;
            if (it == 0) {
                tc5::mbar_wait(bar(XS_FULL + xb), 0);
                tc5::fence_after_sync();
                if (leader) {
                    issue_g1(xb, 0);
                    issue_g1(xb, 1);
                }
                __syncwarp();
            }
            for (int m = 0; m < NMT; ++m) {
                const int e = m & 1;
                V3_WAIT(0, bar(G2_READY + e), n_g2[e] & 1);
                ++n_g2[e];
                const bool next_g1 = (m >= NMT - 2) && (it + 1 < nmine);  // m = 3 -> G1'(1), m = 4 -> G1'(0)
                if (next_g1 && m == NMT - 2) V3_WAIT(2, bar(XS_FULL + ((it + 1) % XS_NBUF)), ((it + 1) / XS_NBUF) & 1);
                tc5::fence_after_sync();
                V3_EVT(100 + m);
                if (leader) {
                    const uint32_t d2 = tmem + d2_col(e), a2 = tmem + d1_col(e);
                    // A2 columns: k-steps 0..3 at columns 8j (written by WG-A), k-steps 4.. at 104 + 8(j-4) (WG-B); see e1()
                    tc5::mma_ts(d2, a2, bw2, idesc32, false);
#ifdef B200SR_EXP_G2SHORT
                    const int nk2x = 2;      // (timing experiment: results are wrong)
#else
                    const int nk2x = nk2;
#endif
                    // unroll sweep at cfg2 / 360p / 1080p (us per launch): 1: 27.7 / 15.4 / 86.3, 2: 25.9 / 14.6 / 81.4, 4: 25.8 / 14.5 / 80.1,
                    // 8: 25.2 / 14.2 / 78.8; fully unrolled behind a uniform guard: 31.5 / 16.9 / 108 (the descriptors then live in vector
                    // registers and every MMA pays R2UR moves)
#pragma unroll 8
                    for (int j = 1; j < nk2x; ++j) tc5::mma_ts(d2, a2 + (j < 4 ? 8 * j : a2hi + 8 * (j - 4)), bw2 + (uint64_t)(16 * j), idesc32, true);
                    tc5::commit(bar(D2_FULL + e));
                    if (m + 2 < NMT) {
                        issue_g1(xb, m + 2);
                        if (m + 2 == NMT - 1) tc5::commit(bar(XS_EMPTY + xb));  // all G1 reads of XS[xb] retired
                    } else if (next_g1) {
                        issue_g1((it + 1) % XS_NBUF, m == NMT - 2 ? 1 : 0);  // D1[1] is free after G2(3), D1[0] after G2(4)
                    }
                }
                __syncwarp();
                V3_EVT(110 + m);
            }
        }
        V3_ADD(5);
        if (nmine > 0) tc5::mbar_wait(bar(D2_FULL + 0), (n_g2[0] - 1) & 1);  // the last G2 (M-tile 4, buffer 0) and all before it retired
      } else if (warp == 2) {
        // ============================== MMA issuer B: the 3x3 stream ==============================
        const bool leader = tc5::elect_one();
        const uint32_t idesc32 = tc5::idesc_bf16_f32(128, 32);
        const uint64_t bw3 = tc5::smem_desc(w_u + L.w3, 128, 28 * 128);
        const uint64_t at0 = tc5::smem_desc(t2_u, 0, T2_GROUP);    // LBO added per instruction
        auto issue_g3_nc = [&](int k, auto ncc) {  // leader only; NC = chunks of t2 (3 dense; 2 or 1 for pruned M2 <= 16 / <= 8)
            constexpr int NCX = decltype(ncc)::value;
            constexpr bool PKX = NCX == 4;
            constexpr int NC = PKX ? 2 : NCX, NS = PKX ? 24 : 9 * NC, NM = (NS + 1) / 2;
            // slice address: q < 9 NC: (dx, dy, chunk) = (q / (3 NC), (q / NC) % 3, q % NC); packed: slices 18 + 2 dy + w = chunk 2 of copy 0 (w = 0:
            // taps dx 0 | 1) or of copy 2 (w = 1: tap dx 2 | zero weights) at row dy
            auto sl = [](int q) constexpr { return q < 9 * NC ? (q / (3 * NC)) * T2_COPY + ((q / NC) % 3) * T2_ROW + (q % NC) * 128
                                                            : ((q - 18) % 2 ? 2 : 0) * T2_COPY + ((q - 18) / 2) * T2_ROW + 2 * 128; };
            const uint64_t abase = at0 + (uint64_t)((k * 4 * T2_ROW) >> 4);
            const uint32_t d3 = tmem + d3_col(k);
#ifdef B200SR_EXP_G3SHORT
            constexpr int NG3 = 2;   // (timing experiment: results are wrong)
#elif defined(B200SR_EXP_DXN)
            // (timing experiment for DESIGN.md 9-1, results are wrong: the tensor-queue shape of the "horizontal taps in N" mapping -- 5 MMAs of
            //  N = 80 per M-tile into one of two 80-column accumulators at TMEM columns 352 / 432)
            constexpr int NG3 = 5;
            const uint32_t idesc80 = tc5::idesc_bf16_f32(128, 80);
            const uint32_t d3x = tmem + 352 + (k & 1) * 80;
#pragma unroll
            for (int i = 0; i < NG3; ++i)
                tc5::mma_ss(d3x, abase + (uint64_t)((i % 3) * T2_ROW >> 4) + ((uint64_t)(128 >> 4) << 16), tc5::smem_desc(w_u + L.w3, 128, 1024) + (uint64_t)(16 * i), idesc80, i > 0);
            tc5::commit(bar(D3_FULL + k));
            if (k == 3) tc5::commit(bar(T2R_FREE + 3));
            return;
#else
            constexpr int NG3 = NM;
#endif
#pragma unroll
            for (int i = 0; i < NG3; ++i) {   // slice q = (dx * 3 + dy) * NC + chunk, two slices per K = 16 instruction through LBO
                const int q0 = 2 * i, q1 = 2 * i + 1;
                const int a0 = sl(q0);
                const int a1 = q1 < NS ? sl(q1) : a0 + 128;
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

        // ---- E1: relu(D1) -> bf16 A2, packed in place.  Column split between the two warpgroups (wg 0: expand channels 0..63 ->
        //      columns 0..31; wg 1: channels 64..M1P-1 -> the top (M1P-64)/2 columns of D1).  Each thread reads ALL of its
        //      columns before it writes, so packing in place inside its own half is safe.
        auto e1 = [&](int eb) {
            tc5::fence_after_sync();
            V3_T0();
            V3_EVT(300);
            const uint32_t d1 = tmem + lane_base + d1_col(eb);
#ifdef B200SR_EXP_E1SHORT
            if (true) { uint32_t v8[8]; tc5::tmem_ld8(d1, v8); tc5::tmem_wait_ld(); if (v8[0] == 0x7fc12345u) tc5::tmem_st8(d1, v8); } else   // (timing experiment)
#endif
            if (wg == 1) {
                uint32_t va[32], vb[32];
                tc5::tmem_ld32(d1, va);
                tc5::tmem_ld32(d1 + 32, vb);
                tc5::tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 16; ++j) va[j] = tc5::relu_pack_bf16x2(va[2 * j], va[2 * j + 1]);
#pragma unroll
                for (int j = 0; j < 16; ++j) va[16 + j] = tc5::relu_pack_bf16x2(vb[2 * j], vb[2 * j + 1]);
                tc5::tmem_st16(d1, *reinterpret_cast<uint32_t(*)[16]>(&va[0]));
                tc5::tmem_st16(d1 + 16, *reinterpret_cast<uint32_t(*)[16]>(&va[16]));
            } else {
                // columns 64 .. M1P-1 (nhi = 0, 16, .. 80 of them) -> nhi/2 packed columns at the top of D1: [M1P - nhi/2, M1P).
                // All loads are issued before the one wait; the shape is a compile-time parameter of the body so that the register
                // arrays stay in registers (pruned widths used to take a 16-columns-at-a-time loop).
                const int nhi = M1P - 64, dst = M1P - nhi / 2;
                auto hi = [&](auto n32c, auto r16c) {
                    constexpr int N32 = decltype(n32c)::value;
                    constexpr bool R16 = decltype(r16c)::value != 0;
                    uint32_t va[32], vb[32], vc[16];
                    if constexpr (N32 >= 1) tc5::tmem_ld32(d1 + 64, va);
                    if constexpr (N32 >= 2) tc5::tmem_ld32(d1 + 96, vb);
                    if constexpr (R16) tc5::tmem_ld16(d1 + 64 + 32 * N32, vc);
                    tc5::tmem_wait_ld();
                    if constexpr (N32 >= 1) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) va[j] = tc5::relu_pack_bf16x2(va[2 * j], va[2 * j + 1]);
                        tc5::tmem_st16(d1 + dst, *reinterpret_cast<uint32_t(*)[16]>(&va[0]));
                    }
                    if constexpr (N32 >= 2) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) vb[j] = tc5::relu_pack_bf16x2(vb[2 * j], vb[2 * j + 1]);
                        tc5::tmem_st16(d1 + dst + 16, *reinterpret_cast<uint32_t(*)[16]>(&vb[0]));
                    }
                    if constexpr (R16) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) vc[j] = tc5::relu_pack_bf16x2(vc[2 * j], vc[2 * j + 1]);
                        tc5::tmem_st8(d1 + dst + 16 * N32, *reinterpret_cast<uint32_t(*)[8]>(&vc[0]));
                    }
                };
                using std::integral_constant;
                if (M1P == 144) {   // the dense width: kept as straight-line code in front of the dispatch
                    uint32_t va[32], vb[32], vc[16];
                    tc5::tmem_ld32(d1 + 64, va);
                    tc5::tmem_ld32(d1 + 96, vb);
                    tc5::tmem_ld16(d1 + 128, vc);
                    tc5::tmem_wait_ld();
#pragma unroll
                    for (int j = 0; j < 16; ++j) va[j] = tc5::relu_pack_bf16x2(va[2 * j], va[2 * j + 1]);
#pragma unroll
                    for (int j = 0; j < 16; ++j) va[16 + j] = tc5::relu_pack_bf16x2(vb[2 * j], vb[2 * j + 1]);
#pragma unroll
                    for (int j = 0; j < 8; ++j) vc[j] = tc5::relu_pack_bf16x2(vc[2 * j], vc[2 * j + 1]);
                    tc5::tmem_st16(d1 + dst, *reinterpret_cast<uint32_t(*)[16]>(&va[0]));
                    tc5::tmem_st16(d1 + dst + 16, *reinterpret_cast<uint32_t(*)[16]>(&va[16]));
                    tc5::tmem_st8(d1 + dst + 32, *reinterpret_cast<uint32_t(*)[8]>(&vc[0]));
                } else switch (nhi) {
                    case 80: hi(integral_constant<int, 2>{}, integral_constant<int, 1>{}); break;
                    case 64: hi(integral_constant<int, 2>{}, integral_constant<int, 0>{}); break;
                    case 48: hi(integral_constant<int, 1>{}, integral_constant<int, 1>{}); break;
                    case 32: hi(integral_constant<int, 1>{}, integral_constant<int, 0>{}); break;
                    case 16: hi(integral_constant<int, 0>{}, integral_constant<int, 1>{}); break;
                    default: break;   // M1P <= 64: nothing in the upper half
                }
            }
            tc5::tmem_wait_st();
            tc5::fence_before_sync();
            tc5::mbar_arrive_relaxed(bar(G2_READY + eb));  // A2 half complete (wait::st); no release: E3's output stores may be in flight
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
#if defined(B200SR_EXP_DXN) && !defined(B200SR_EXP_DXN_LIGHTE3) && !defined(B200SR_EXP_DXN_HALFE3)
            {   // (timing experiment: the heavier E3 of the mapping -- 72 accumulator columns and two shuffles per output value)
                uint32_t v1[32], v2[16];
                tc5::tmem_ld32(tmem + lane_base + 352 + (k & 1) * 80, v);
                tc5::tmem_ld32(tmem + lane_base + 352 + (k & 1) * 80 + 32, v1);
                tc5::tmem_ld16(tmem + lane_base + 352 + (k & 1) * 80 + 64, v2);
                tc5::tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 24; ++j)
                    v[j] = __float_as_uint(__uint_as_float(v[j]) + __shfl_down_sync(0xffffffffu, __uint_as_float(j < 8 ? v[24 + j] : v1[j - 8]), 1) +
                                           __shfl_down_sync(0xffffffffu, __uint_as_float(j < 8 ? v1[16 + j] : j < 16 ? v1[24 + j - 8] : v2[j - 16]), 2));
            }
#elif defined(B200SR_EXP_DXN) && defined(B200SR_EXP_DXN_HALFE3)
            {   // (proxy for "the combine split over two warpgroups": half of the extra columns and shuffles on this one)
                uint32_t v1[16];
                tc5::tmem_ld32(tmem + lane_base + 352 + (k & 1) * 80, v);
                tc5::tmem_ld16(tmem + lane_base + 352 + (k & 1) * 80 + 32, v1);
                tc5::tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 12; ++j)
                    v[j] = __float_as_uint(__uint_as_float(v[j]) + __shfl_down_sync(0xffffffffu, __uint_as_float(v[12 + j]), 1) +
                                           __shfl_down_sync(0xffffffffu, __uint_as_float(v1[j]), 2));
            }
#elif defined(B200SR_EXP_DXN)
            tc5::tmem_ld32(tmem + lane_base + 352 + (k & 1) * 80, v);   // (upper bound of the mapping: E3 as light as the shipped one)
#else
            tc5::tmem_ld32(tmem + lane_base + d3_col(k), v);
#endif
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
            tc5::mbar_arrive_relaxed(bar(G3_READY + k));
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
            tc5::tmem_ld32(tmem + lane_base + d2_col(e), v);
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
                uint4 c[NCH];   // only the chunks this block has (compile-time sized: a partly used array went to local memory)
                uint32_t *cw = reinterpret_cast<uint32_t *>(c);
#pragma unroll
                for (int j4 = 0; j4 < 2 * NCH; ++j4) {
                    const float4 bb = *reinterpret_cast<const float4 *>(b2s + 4 * j4);  // broadcast read
                    cw[2 * j4] = ok ? pack_bf16x2(__uint_as_float(v[4 * j4]) + bb.x, __uint_as_float(v[4 * j4 + 1]) + bb.y) : 0u;
                    cw[2 * j4 + 1] = ok ? pack_bf16x2(__uint_as_float(v[4 * j4 + 2]) + bb.z, __uint_as_float(v[4 * j4 + 3]) + bb.w) : 0u;
                }
#pragma unroll
#ifdef B200SR_EXP_DXN
                for (int d = 0; d < 1; ++d) {      // (timing experiment: ONE copy of t2)
#else
                for (int d = 0; d < 3; ++d) {
#endif
                    const int xi = hx - d;
                    if (xi >= 0 && xi < TW) {
                        uint8_t *dst = t2 + d * T2_COPY + r * T2_ROW + (xi >> 3) * T2_GROUP + (xi & 7) * 16;
#pragma unroll
                        for (int q = 0; q < (PK ? 2 : NCH); ++q) *reinterpret_cast<uint4 *>(dst + q * 128) = c[q];
                        if (PK && d == 0) *reinterpret_cast<uint2 *>(dst + 2 * 128) = make_uint2(cw[8], cw[9]);            // own 4 channels: low half, taps dx 0
                        if (PK && d == 1) *reinterpret_cast<uint2 *>(dst - 1 * T2_COPY + 2 * 128 + 8) = make_uint2(cw[8], cw[9]);   // copy 0, slot hx - 1: high half, tap dx 1
                        if (PK && d == 2) *reinterpret_cast<uint4 *>(dst + 2 * 128) = make_uint4(cw[8], cw[9], 0u, 0u);    // copy 2, slot hx - 2: tap dx 2 | zero weights
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
            // WG1 / WG2: their column half of E1 for every M-tile, in issue order.  M-tile i % 5 of tile i / 5 lives in buffer
            // (i % 5) & 1; a buffer completes 3 (even) or 2 (odd) G1s per tile.
            for (int i1 = 0; i1 < NMT * nmine; ++i1) {
                const int mm = i1 % NMT, eb = mm & 1, nth = (i1 / NMT) * (eb == 0 ? 3 : 2) + (mm >> 1);
                V3_WAIT(0, bar(D1_FULL + eb), nth & 1);
                e1(eb);
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
            tc5::setmaxnreg_dec<72>();
            // WG5: E3 of 3x3 M-tiles k = 0..3 of every tile
#pragma unroll
            for (int k = 0; k < 4; ++k) tc5::mbar_arrive(bar(G3_READY + k));  // stand-ins for "previous tile's E3 drained D3[k]"
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
    if (warp == 0) tc5::tmem_free(tmem, tc5v3::TMEM_COLS);
}

}  // namespace b200sr
