// wdsr_tc5_layout.cuh -- operand-image layout and tile constants shared by the host packer and the tcgen05 kernels.
#pragma once
#include "common.cuh"

namespace b200sr {

struct BlockTc5Layout {  // weight image, bytes (host builds it, kernel copies it verbatim to shared memory)
    int w1, w2, w3, b2, b3, total, sbo2;
    __host__ __device__ BlockTc5Layout(int M1P) {
        w1 = 0;                           // [M1P/8][c0,c1,c2,BIAS][8 rows][8]      rows = expand channel
        w2 = w1 + (M1P / 8) * 512;        // [4][M1P/8 chunks][8 rows][8]           rows = reduce channel (32, 20 used)
        sbo2 = (M1P / 8) * 128;
        w3 = w2 + 4 * sbo2;               // [4][28 chunks][8 rows][8]              rows = out channel (32, 24 used)
        b2 = w3 + 4 * 28 * 128;           // f32[32]
        b3 = b2 + 128;                    // f32[32]
        total = b3 + 128;
    }
};

namespace tc5cfg {
constexpr int TW = 32, TH = 16, HW_ = TW + 2, HH_ = TH + 2, HP = HW_ * HH_;  // 612 halo pixels
constexpr int NMT = (HP + 127) / 128;                                        // 5 M-tiles for G1/G2
constexpr int XS_GROUP = 512;                                                // c0,c1,c2,ONE
constexpr int XS_BYTES = NMT * 16 * XS_GROUP;                                // 40,960
constexpr int T2_GROUP = 384;                                                // c0,c1,c2
constexpr int T2_ROW = (TW / 8) * T2_GROUP;                                  // 1,536
constexpr int T2_COPY = HH_ * T2_ROW;                                        // 27,648
constexpr int T2_BYTES = 3 * T2_COPY + 256;                                  // + zero pad read by the dummy half of instr 13
constexpr int TMEM_COLS = 256, D1_COL = 0, D2_COL = 160, D3_COL = 192;
constexpr int CTRL_BYTES = 128;
}  // namespace tc5cfg

}  // namespace b200sr
