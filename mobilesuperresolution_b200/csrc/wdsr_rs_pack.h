// wdsr_rs_pack.h -- host-side packing of one block's folded filters into the operand image of the row-streaming tcgen05 block
// (wdsr_rs.cuh / wdsr_rs_layout.cuh).  Host only, no CUDA calls: b200sr.cu uploads the image; tests read it back through
// b200sr_debug_pack_block_rs and replay the kernel's data flow on the CPU.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

#include "wdsr_rs_layout.cuh"

namespace b200sr {

inline uint16_t rs_f2bf(float f) {  // round-to-nearest-even (finite values), as __float2bfloat16_rn
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}

inline int rs_nc2(int M2) { return M2 <= 8 ? 1 : M2 <= 16 ? 2 : 3; }
inline bool rs_pack20(int M2) { return M2 > 16 && M2 <= 20; }

// One K = 8 half of a 3x3 instruction: 8 (channel, dx) pairs; channel < 0 = zero weight.
struct RsHalf { int ch[8], dx[8]; };
struct RsSlices {
    int ng3 = 0;
    int a_off[BlockRsLayout::MAXG3] = {}, a_lbo[BlockRsLayout::MAXG3] = {};
    RsHalf half[BlockRsLayout::MAXG3][2];
};

// The A-operand slices of the 3x3 for M2 reduce channels.  A t2 row is chunk-planar: plane c (channels 8c..8c+7) starts at
// c * T2PLANE, lane l at entry l + 1; the slice "plane c at tap dx" starts dx * 16 bytes into the plane (row l reads lane l + dx - 1).
inline RsSlices rs_slices(int M2) {
    using namespace rs;
    RsSlices s;
    auto plain = [&](RsHalf &h, int c, int dx) { for (int j = 0; j < 8; ++j) h.ch[j] = 8 * c + j < M2 ? 8 * c + j : -1, h.dx[j] = dx; };
    auto none = [&](RsHalf &h) { for (int j = 0; j < 8; ++j) h.ch[j] = -1, h.dx[j] = 0; };
    auto add = [&](int off, int lbo) { s.a_off[s.ng3] = off, s.a_lbo[s.ng3] = lbo; return s.ng3++; };
    const int nc2 = rs_nc2(M2);
    if (nc2 == 1) {
        int i = add(0, 16);            // (c0, dx0) | (c0, dx1): the second half is the next lane of the same plane
        plain(s.half[i][0], 0, 0), plain(s.half[i][1], 0, 1);
        i = add(32, 16);               // (c0, dx2) | zero weights (reads lane + 2: finite)
        plain(s.half[i][0], 0, 2), none(s.half[i][1]);
        return s;
    }
    for (int dx = 0; dx < 3; ++dx) {   // (c0, dx) | (c1, dx): paired through the plane stride
        const int i = add(dx * 16, T2PLANE);
        plain(s.half[i][0], 0, dx), plain(s.half[i][1], 1, dx);
    }
    if (nc2 == 2) return s;
    if (rs_pack20(M2)) {
        // third plane, PACK form: entry l+1 = [channels 16..19 of lane l | channels 16..19 of lane l + 1]
        const int i = add(2 * T2PLANE, 32);   // entry at tap dx0 = (lane-1 | lane) = taps dx0, dx1;  32 bytes on = (lane+1 | lane+2) = tap dx2, nothing
        for (int j = 0; j < 8; ++j) {
            const int ch = 16 + (j & 3);
            s.half[i][0].ch[j] = ch < M2 ? ch : -1, s.half[i][0].dx[j] = j < 4 ? 0 : 1;
            s.half[i][1].ch[j] = (j < 4 && ch < M2) ? ch : -1, s.half[i][1].dx[j] = 2;
        }
        return s;
    }
    int i = add(2 * T2PLANE, 16);      // (c2, dx0) | (c2, dx1)
    plain(s.half[i][0], 2, 0), plain(s.half[i][1], 2, 1);
    i = add(2 * T2PLANE + 32, 16);     // (c2, dx2) | zero weights
    plain(s.half[i][0], 2, 2), none(s.half[i][1]);
    return s;
}

// w1 (M1, C), w2 (M2, M1), w3 (C, M2, 3, 3) folded fp32 filters in the reference's layouts; C <= 24, M1 <= M1P <= 144, M2 <= 24.
inline void pack_block_rs(std::vector<uint8_t> &img, int C, int M1, int M2, int M1P, const float *w1, const float *b1, const float *w2,
                          const float *b2, const float *w3, const float *b3) {
    const BlockRsLayout L(M1P);
    img.assign((size_t)L.total, 0);
    auto at = [&](int off) { return (uint16_t *)(img.data() + off); };
    for (int n = 0; n < M1; ++n) {
        for (int c = 0; c < C; ++c) at(L.w1 + (n / 8) * 512 + (c / 8) * 128 + (n % 8) * 16)[c % 8] = rs_f2bf(w1[(size_t)n * C + c]);
        const uint16_t hi = rs_f2bf(b1[n]);
        uint32_t hu = (uint32_t)hi << 16;
        float hf;
        memcpy(&hf, &hu, 4);
        uint16_t *bc = at(L.w1 + (n / 8) * 512 + 3 * 128 + (n % 8) * 16);
        bc[0] = hi, bc[1] = rs_f2bf(b1[n] - hf);   // b1 = hi + lo against the two constant-one channels
    }
    for (int j = 0; j < M2; ++j)
        for (int m = 0; m < M1; ++m) at(L.w2 + (j / 8) * L.sbo2 + (m / 8) * 128 + (j % 8) * 16)[m % 8] = rs_f2bf(w2[(size_t)j * M1 + m]);
    const RsSlices s = rs_slices(M2);
    for (int g = 0; g < 3; ++g) {          // dy group: columns 32g.. of the instruction = output row s-1+g = vertical tap dy = 2 - g
        const int dy = 2 - g;
        for (int o = 0; o < C; ++o) {
            const int n = g * 32 + o;
            for (int i = 0; i < s.ng3; ++i)
                for (int h = 0; h < 2; ++h)
                    for (int j = 0; j < 8; ++j) {
                        const int ch = s.half[i][h].ch[j], dx = s.half[i][h].dx[j];
                        if (ch >= 0) at(L.w3 + (n / 8) * L.sbo3 + (2 * i + h) * 128 + (n % 8) * 16)[j] = rs_f2bf(w3[((size_t)o * M2 + ch) * 9 + dy * 3 + dx]);
                    }
        }
    }
    // reduce filter as mma.sync B fragments (wdsr_rh.cuh) in the K order of tcgen05.ld.16x128b (tc5.cuh): thread (g = lane / 4,
    // j = lane % 4) of n-tile nt, k-step ks holds {w2[8nt + g][16ks + j], [16ks + 4 + j]} and {w2[8nt + g][16ks + 8 + j], [16ks + 12 + j]}
    for (int ks = 0; ks < M1P / 16; ++ks)
        for (int nt = 0; nt < 3; ++nt)
            for (int lane = 0; lane < 32; ++lane) {
                const int n = 8 * nt + lane / 4, k0 = 16 * ks + lane % 4;
                auto w = [&](int k) { return (n < M2 && k < M1) ? rs_f2bf(w2[(size_t)n * M1 + k]) : (uint16_t)0; };
                uint16_t *f = at(L.w2f + ((ks * 3 + nt) * 32 + lane) * 8);
                f[0] = w(k0), f[1] = w(k0 + 4), f[2] = w(k0 + 8), f[3] = w(k0 + 12);
            }
    float *pb2 = (float *)(img.data() + L.b2), *pb3 = (float *)(img.data() + L.b3);
    for (int j = 0; j < M2; ++j) pb2[j] = b2[j];
    for (int o = 0; o < C; ++o) pb3[o] = b3[o];
    int *tab = (int *)(img.data() + L.tab);
    tab[0] = s.ng3;
    for (int i = 0; i < BlockRsLayout::MAXG3; ++i) {
        const int k = i < s.ng3 ? i : 0;     // unused entries repeat slice 0 (never issued)
        tab[1 + i] = s.a_off[k], tab[1 + BlockRsLayout::MAXG3 + i] = s.a_lbo[k];
    }
}

}  // namespace b200sr
