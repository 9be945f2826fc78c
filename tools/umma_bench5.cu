// umma_bench5.cu -- the G1/G2 stream and the G3 stream issued CONCURRENTLY from two warps (as wdsr_tc5p.cuh does) vs from one (developer probe).
#include <cstdio>
#include "tc5.cuh"
#include "wdsr_tc5_layout.cuh"
using namespace b200sr;
using namespace b200sr::tc5cfg;
__device__ unsigned long long g_out[64];
constexpr int XS_PLANE = 10240, XS_BUF = 40960;

__global__ void __launch_bounds__(256, 1) bench(int reps) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *ctrl = smem, *xs = smem + 256, *t2 = xs + 2 * XS_BUF, *wsm = t2 + T2_BYTES;
    const BlockTc5Layout L(144);
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t bar = smem_u32(ctrl);
    if (tid == 0) { tc5::mbar_init(bar, 2); for (int i = 8; i < 12; ++i) tc5::mbar_init(bar + 8 * i, 1); tc5::mbar_init_fence(); }
    __syncwarp();
    if (warp == 0) tc5::tmem_alloc(smem_u32(ctrl + 16), 512);
    for (int i = tid; i < (2 * XS_BUF + T2_BYTES + L.total) / 16; i += 256) *reinterpret_cast<uint4 *>(xs + i * 16) = make_uint4(0, 0, 0, 0);
    tc5::fence_proxy_async(); tc5::fence_before_sync(); __syncthreads(); tc5::fence_after_sync();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t *>(ctrl + 16);
    const uint32_t xs_u = smem_u32(xs), t2_u = smem_u32(t2), w_u = smem_u32(wsm);
    uint32_t phase = 0;
    const uint32_t idesc1 = tc5::idesc_bf16_f32(128, 144), idesc32 = tc5::idesc_bf16_f32(128, 32);
    const uint64_t bw1a = tc5::smem_desc(w_u + L.w1, 128, 512), bw1b = tc5::smem_desc(w_u + L.w1 + 256, 128, 512);
    const uint64_t bw2 = tc5::smem_desc(w_u + L.w2, 128, L.sbo2), bw3 = tc5::smem_desc(w_u + L.w3, 128, 28 * 128);
    const uint64_t ax0 = tc5::smem_desc(xs_u, XS_PLANE, 128), at0 = tc5::smem_desc(t2_u, 0, T2_GROUP);

    // mode 0: one warp issues everything (kernel order); mode 1: warp 0 = G1/G2, warp 5 = G3 (different SMSP); mode 2: warp 0 / warp 4 (same SMSP)
    for (int mode = 0; mode < 3; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            long long t0 = clock64(), t1 = 0;
            const int wb = mode == 1 ? 5 : 4;
            if (warp == 0) {
                tc5::fence_after_sync();
                if (tc5::elect_one()) {
                    for (int r = 0; r < reps; ++r)
                        for (int m = 0; m < 5; ++m) {
                            const uint32_t d2 = tmem + 288 + (m & 1) * 32, a2 = tmem + (m & 1) * 144;
                            tc5::mma_ts(d2, a2, bw2, idesc32, false);
#pragma unroll 4
                            for (int j = 1; j < 9; ++j) tc5::mma_ts(d2, a2 + 8 * j, bw2 + (uint64_t)(16 * j), idesc32, true);
                            const uint64_t a = ax0 + (uint64_t)((m * 2048) >> 4);
                            tc5::mma_ss(tmem + (m & 1) * 144, a, bw1a, idesc1, false);
                            tc5::mma_ss(tmem + (m & 1) * 144, a + (uint64_t)((2 * XS_PLANE) >> 4), bw1b, idesc1, true);
                            if (mode == 0 && m >= 1) {
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
                    if (mode == 0) tc5::commit(bar);
                }
                t1 = clock64();
                __syncwarp();
            } else if (warp == wb && mode > 0) {
                tc5::fence_after_sync();
                if (tc5::elect_one()) {
                    for (int r = 0; r < reps; ++r)
                        for (int k = 0; k < 4; ++k) {
                            const uint64_t abase = at0 + (uint64_t)((k * 4 * T2_ROW) >> 4);
#pragma unroll
                            for (int i = 0; i < 14; ++i) {
                                const int q0 = 2 * i, q1 = 2 * i + 1;
                                const int a0 = (q0 / 9) * T2_COPY + ((q0 / 3) % 3) * T2_ROW + (q0 % 3) * 128;
                                const int a1 = q1 < 27 ? (q1 / 9) * T2_COPY + ((q1 / 3) % 3) * T2_ROW + (q1 % 3) * 128 : a0 + 128;
                                tc5::mma_ss(tmem + 352 + k * 32, abase + (uint64_t)(a0 >> 4) + ((uint64_t)((a1 - a0) >> 4) << 16), bw3 + (uint64_t)(16 * i), idesc32, i > 0);
                            }
                        }
                    tc5::commit(bar);
                }
                t1 = clock64();
                __syncwarp();
            }
            if (warp != 0 || tid == 0) tc5::mbar_wait(bar, phase);
            phase ^= 1;
            if (tid == 0 && rep == 1) { g_out[3 * mode] = (unsigned long long)(t1 - t0); g_out[3 * mode + 1] = (unsigned long long)(clock64() - t0); }
            if (tid == wb * 32 && rep == 1) g_out[3 * mode + 2] = (unsigned long long)(t1 - t0);
            tc5::fence_before_sync();
            __syncthreads();
        }
    }
    if (warp == 0) tc5::tmem_free(tmem, 512);
}

int main() {
    const int reps = 8;
    size_t smem = 256 + 2 * XS_BUF + T2_BYTES + BlockTc5Layout(144).total;
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    bench<<<1, 256, smem>>>(reps);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    unsigned long long out[64]; cudaMemcpyFromSymbol(out, g_out, sizeof out);
    const char *names[] = {"one issuer (kernel order)", "two issuers, different SMSP", "two issuers, same SMSP"};
    for (int i = 0; i < 3; ++i)
        printf("%-32s per tile: issue A %6.0f  issue B %6.0f  complete %6.0f clk\n", names[i], (double)out[3 * i] / reps, (double)out[3 * i + 2] / reps, (double)out[3 * i + 1] / reps);
    return 0;
}
