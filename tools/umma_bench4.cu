// umma_bench4.cu -- how much do concurrent tcgen05.ld/st (epilogue-like) and shared-memory traffic slow the MMA stream of one tile?
// warp 0 issues the kernel's full-tile MMA mix; warps 4.. are "disturbers" that run one access pattern until the MMAs retire.
#include <cstdio>
#include "tc5.cuh"
#include "wdsr_tc5_layout.cuh"
using namespace b200sr;
using namespace b200sr::tc5cfg;
__device__ unsigned long long g_out[128];
__device__ unsigned g_sink;
constexpr int XS_PLANE = 10240, XS_BUF = 40960;

__global__ void __launch_bounds__(640, 1) bench(int reps) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *ctrl = smem, *xs = smem + 256, *t2 = xs + 2 * XS_BUF, *wsm = t2 + T2_BYTES;
    const BlockTc5Layout L(144);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar = smem_u32(ctrl);
    volatile int *flag = reinterpret_cast<volatile int *>(ctrl + 128);
    if (tid == 0) { tc5::mbar_init(bar, 1); tc5::mbar_init_fence(); *flag = 0; }
    __syncwarp();
    if (warp == 0) tc5::tmem_alloc(smem_u32(ctrl + 16), 512);
    for (int i = tid; i < (2 * XS_BUF + T2_BYTES + L.total) / 16; i += 640) *reinterpret_cast<uint4 *>(xs + i * 16) = make_uint4(0, 0, 0, 0);
    tc5::fence_proxy_async(); tc5::fence_before_sync(); __syncthreads(); tc5::fence_after_sync();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t *>(ctrl + 16);
    const uint32_t xs_u = smem_u32(xs), t2_u = smem_u32(t2), w_u = smem_u32(wsm);
    uint32_t phase = 0;
    const uint32_t idesc1 = tc5::idesc_bf16_f32(128, 144), idesc32 = tc5::idesc_bf16_f32(128, 32);
    const uint64_t bw1a = tc5::smem_desc(w_u + L.w1, 128, 512), bw1b = tc5::smem_desc(w_u + L.w1 + 256, 128, 512);
    const uint64_t bw2 = tc5::smem_desc(w_u + L.w2, 128, L.sbo2), bw3 = tc5::smem_desc(w_u + L.w3, 128, 28 * 128);
    const uint64_t ax0 = tc5::smem_desc(xs_u, XS_PLANE, 128), at0 = tc5::smem_desc(t2_u, 0, T2_GROUP);
    // mode = disturbance (0 none, 1 tmem ld x32, 2 E1-like ld 64 cols + st 32 cols, 3 st.shared 16 B x 9, 4 ld.shared 16 B x 3, 5 global stores)
    // nw = number of disturber warps (4, 8, 16)
    int slot = 0;
    for (int mma_on = 1; mma_on >= 0; --mma_on)
    for (int mode = 1; mode < 6; ++mode)
        for (int nw = 4; nw <= 16; nw *= 2) {
            long long t0 = 0;
            unsigned long long cnt = 0;
            __syncthreads();
            if (warp == 0) {
                tc5::fence_after_sync();
                t0 = clock64();
                if (!mma_on) { while (clock64() - t0 < 30000) {} }
                else if (tc5::elect_one()) {
                    for (int r = 0; r < reps; ++r)
                        for (int m = 0; m < 5; ++m) {
                            const uint32_t d2 = tmem + 288 + (m & 1) * 32, a2 = tmem + (m & 1) * 144;
                            tc5::mma_ts(d2, a2, bw2, idesc32, false);
#pragma unroll 4
                            for (int j = 1; j < 9; ++j) tc5::mma_ts(d2, a2 + 8 * j, bw2 + (uint64_t)(16 * j), idesc32, true);
                            const uint64_t a = ax0 + (uint64_t)((m * 2048) >> 4);
                            tc5::mma_ss(tmem + (m & 1) * 144, a, bw1a, idesc1, false);
                            tc5::mma_ss(tmem + (m & 1) * 144, a + (uint64_t)((2 * XS_PLANE) >> 4), bw1b, idesc1, true);
                            if (m >= 1) {
                                const int k = m - 1;
                                const uint64_t abase = at0 + (uint64_t)((k * 4 * T2_ROW) >> 4);
#pragma unroll
                                for (int i = 0; i < 14; ++i) {
                                    const int q0 = 2 * i, q1 = 2 * i + 1;
                                    const int a0 = (q0 / 9) * T2_COPY + ((q0 / 3) % 3) * T2_ROW + (q0 % 3) * 128;
                                    const int a1 = q1 < 27 ? (q1 / 9) * T2_COPY + ((q1 / 3) % 3) * T2_ROW + (q1 % 3) * 128 : a0 + 128;
                                    tc5::mma_ss(tmem + 352 + k * 32, abase + (uint64_t)(a0 >> 4) + ((uint64_t)((a1 - a0) >> 4) << 16), bw3 + (uint64_t)(16 * i), idesc32, i > 0);
                                }
                            }
                        }
                    tc5::commit(bar);
                }
                __syncwarp();
                if (mma_on) tc5::mbar_wait(bar, phase);
                if (lane == 0) { g_out[2 * slot] = (unsigned long long)(clock64() - t0); *flag = slot + 1; }
                __syncwarp();
            } else if (warp >= 4 && warp < 4 + nw && mode > 0) {
                // disturbers: TMEM columns 480..511 are not touched by the MMAs; lanes = own quarter
                const uint32_t tb = tmem + ((uint32_t)((warp & 3) * 32) << 16);
                unsigned acc = 0;
                uint8_t *sp = t2 + (warp - 4) * 4608 + lane * 16;   // inside T2 (zeros; MMAs only read it)
                while (*flag != slot + 1) {
                  for (int inner = 0; inner < 8; ++inner) {
                    if (mode == 1) {
                        uint32_t v[32]; tc5::tmem_ld32(tb + 448, v); tc5::tmem_wait_ld(); acc += v[3];
                    } else if (mode == 2) {
                        uint32_t va[32], vb[32];
                        tc5::tmem_ld32(tb + 448, va); tc5::tmem_ld32(tb + 480, vb); tc5::tmem_wait_ld();
#pragma unroll
                        for (int j = 0; j < 16; ++j) va[j] = tc5::relu_pack_bf16x2(va[2 * j], va[2 * j + 1]);
#pragma unroll
                        for (int j = 0; j < 16; ++j) va[16 + j] = tc5::relu_pack_bf16x2(vb[2 * j], vb[2 * j + 1]);
                        tc5::tmem_st16(tb + 448, *reinterpret_cast<uint32_t(*)[16]>(&va[0]));
                        tc5::tmem_st16(tb + 464, *reinterpret_cast<uint32_t(*)[16]>(&va[16]));
                        tc5::tmem_wait_st();
                    } else if (mode == 3) {
#pragma unroll
                        for (int q = 0; q < 9; ++q) asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" :: "r"(smem_u32(sp) + q * 512), "r"(0u) : "memory");
                    } else if (mode == 4) {
#pragma unroll
                        for (int q = 0; q < 3; ++q) { uint32_t a0, a1, a2, a3; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3) : "r"(xs_u + q * XS_PLANE + (tid & 511) * 16)); acc += a0 ^ a3; }
                    } else {
                        acc += clock();
                    }
                    ++cnt;
                  }
                }
                if (acc == 0x1234567) g_sink = acc;
                if (warp == 4 && lane == 0) g_out[2 * slot + 1] = cnt;
            }
            __syncthreads();
            if (mma_on) phase ^= 1;
            ++slot;
        }
    if (warp == 0) tc5::tmem_free(tmem, 512);
}

int main() {
    const int reps = 8;
    size_t smem = 256 + 2 * XS_BUF + T2_BYTES + BlockTc5Layout(144).total;
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    bench<<<1, 640, smem>>>(reps);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    unsigned long long out[128]; cudaMemcpyFromSymbol(out, g_out, sizeof out);
    const char *names[] = {"none", "tmem ld x32 loop", "E1-like ld64+cvt+st32 loop", "st.shared 9x16B loop", "ld.shared 3x16B loop", "nanosleep pollers"};
    int slot = 0;
    for (int mma_on = 1; mma_on >= 0; --mma_on)
    for (int mode = 1; mode < 6; ++mode)
        for (int nw = 4; nw <= 16; nw *= 2, ++slot)
            printf("%s %-28s %2d warps: MMA tile mix %6.0f clk/tile   disturber iters/tile %.1f (=> %.0f clk/iter)\n", mma_on ? "MMA on " : "MMA off", names[mode], nw, (double)out[2 * slot] / reps,
                   (double)out[2 * slot + 1] / reps, out[2 * slot + 1] ? (double)out[2 * slot] / out[2 * slot + 1] : 0.0);
    return 0;
}
