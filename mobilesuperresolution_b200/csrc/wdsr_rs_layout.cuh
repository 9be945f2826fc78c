// wdsr_rs_layout.cuh -- operand-image layout and ring constants of the row-streaming fused block (wdsr_rs.cuh), shared by the
// host packer (wdsr_rs_pack.h), the launcher and the kernel.
#pragma once
#include "common.cuh"

namespace b200sr {

struct BlockRsLayout {  // weight image, bytes (host builds it, the kernel copies it verbatim to shared memory)
    static constexpr int MAXG3 = 5;
    int w1, w2, w3, b2, b3, tab, w2f, total, sbo2, sbo3;
    __host__ __device__ BlockRsLayout(int M1P) {
        w1 = 0;                           // [M1P/8][c0,c1,c2,BIAS][8 rows][8]      rows = expand channel (same as BlockTc5Layout)
        w2 = w1 + (M1P / 8) * 512;        // [4][M1P/8 chunks][8 rows][8]           rows = reduce channel (32, <= 24 used)
        sbo2 = (M1P / 8) * 128;
        w3 = w2 + 4 * sbo2;               // [12][2 * MAXG3 chunks][8 rows][8]      rows = dy group (2,1,0) * 32 + out channel
        sbo3 = MAXG3 * 256;
        b2 = w3 + 12 * sbo3;              // f32[32]
        b3 = b2 + 128;                    // f32[32]
        tab = b3 + 128;                   // int32: ng3, a_off[MAXG3], a_lbo[MAXG3]   (A-operand slices of the 3x3, bytes)
        w2f = tab + 64;                   // [M1P/16 k-steps][3 n-tiles][32 lanes][2] u32: B fragments of mma.sync.m16n8k16 (wdsr_rh.cuh)
        total = w2f + (M1P / 16) * 3 * 256;
    }
};

namespace rs {
constexpr int NTHREADS = 1024;
constexpr int TMEM_COLS = 512;
constexpr int NX = 16;                        // X ring: trunk rows in flight (a row is needed again by E3 ~4 steps after G1)
constexpr int XPLANE = 128 * 16;              // one 8-channel plane of one row
constexpr int XSLOT = 3 * XPLANE;
constexpr int X_ONE = NX * XSLOT;             // byte offset of the constant-one plane (carries b1)
constexpr int X_BYTES = X_ONE + XPLANE;
constexpr int NT = 5;                         // t2 ring == OUT ring
constexpr int T2PLANE = 130 * 16;             // lane l at entry l + 1: entries 0 and 129 are only read by the halo lanes
constexpr int T2SLOT = 6272;                  // 3 planes + 32 zero bytes (read by the dummy half of the last slice)
constexpr int T2_BYTES = NT * T2SLOT;
constexpr int SPAN = 126;                     // output lanes per strip
__host__ __device__ constexpr int d1_col(int e) { return e * 144; }
// packed A2 columns of G2's K step j (16 expand channels) inside D1: E1 warpgroup q packs channels 32q.. at column 32q, the last one 96..
__host__ __device__ constexpr int a2_col(int j) { return j < 6 ? 32 * (j / 2) + 8 * (j % 2) : 96 + 8 * (j - 6); }
__host__ __device__ constexpr int d2_col(int e) { return 288 + e * 32; }
__host__ __device__ constexpr int out_col(int k) { return 352 + k * 32; }
//   G2_READY[e] (640) = the four E1 warpgroups wrote A2 into D1[e] (and their copies of trunk row s+2 have landed)  +  E2 of step s-2 drained D2[e]
//   G3_READY[b] (256) = E2 wrote t2 row s (b = s % 5)          +  E3 of row s-4 re-zeroed OUT slot (s+1) % 5
//   STEP_DONE[b] (1)  = commit after G3(s): row s-1 is complete (E3) and t2 slot b may be overwritten (E2 of step s+5)
enum Bar { X_START = 0, X_EMPTY = 1, D1_FULL = 1 + NX, G2_READY = D1_FULL + 2, D2_FULL = G2_READY + 2, G3_READY = D2_FULL + 2,
           STEP_DONE = G3_READY + NT, NBARS = STEP_DONE + NT };
constexpr int CTRL_BYTES = 512;               // mbarriers + tmem base pointer at byte 496
static_assert(NBARS * 8 <= 480, "control block");
__host__ __device__ inline size_t smem_bytes(int M1P) { return (size_t)CTRL_BYTES + X_BYTES + T2_BYTES + (size_t)BlockRsLayout(M1P).total; }

__host__ __device__ inline int num_strips(int N, int W) { return ((N * (W + 2) - 2) + SPAN - 1) / SPAN; }

}  // namespace rs

}  // namespace b200sr
