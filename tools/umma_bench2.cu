// umma_bench2.cu -- issue cost vs completion cost of tcgen05.mma under different issue styles (developer probe).
#include <cstdio>
#include <vector>
#include "tc5.cuh"
using namespace b200sr;
__device__ unsigned long long g_out[64];

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

template <int STYLE, int N, int TS = 0>
__device__ void run(uint32_t tmem, uint32_t a_u, uint32_t b_u, uint32_t bar, uint32_t &phase, int slot, int R) {
    const int tid = threadIdx.x, warp = tid >> 5;
    constexpr uint32_t idesc = tc5::idesc_bf16_f32(128, N);
    for (int rep = 0; rep < 2; ++rep) {
        long long t0 = 0, t1 = 0;
        if (warp == 0) {
            const uint64_t ad = tc5::smem_desc(a_u, 128, 256), bd = tc5::smem_desc(b_u, 128, 256);
            tc5::fence_after_sync();
            t0 = clock64();
            if (STYLE == 0) {          // lane 0 only, rolled loop
                if (tid == 0) { for (int r = 0; r < R; ++r) tc5::mma_ss(tmem + 256, ad, bd, idesc, r > 0); tc5::commit(bar); }
            } else if (STYLE == 1) {   // converged warp, elect.sync per instruction
                for (int r = 0; r < R; ++r) if (elect_one()) tc5::mma_ss(tmem + 256, ad, bd, idesc, r > 0);
                if (elect_one()) tc5::commit(bar);
            } else if (STYLE == 2) {   // converged warp, one elect, unrolled x8
                if (elect_one()) {
                    for (int r = 0; r < R; r += 8) {
#pragma unroll
                        for (int u = 0; u < 8; ++u) tc5::mma_ss(tmem + 256, ad + 2 * u, bd, idesc, (r + u) > 0);
                    }
                    tc5::commit(bar);
                }
            } else {                   // like 2 but accumulate flag constant (no setp dependence)
                if (elect_one()) {
                    tc5::mma_ss(tmem + 256, ad, bd, idesc, false);
                    for (int r = 8; r < R; r += 8) {
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            if (TS == 1) tc5::mma_ts(tmem + 256, tmem + 8 * u, bd + 16 * u, idesc, true);
                            else if (TS == 2) { if (u & 1) tc5::mma_ts(tmem + 256, tmem + 8 * u, bd + 16 * u, idesc, true); else tc5::mma_ss(tmem + 320, ad + 16 * u, bd, idesc, true); }
                            else tc5::mma_ss(tmem + 256, ad + 16 * u, bd + 16 * u, idesc, true);
                        }
                    }
                    tc5::commit(bar);
                }
            }
            t1 = clock64();
            __syncwarp();
        }
        if (warp != 0 || tid == 0) tc5::mbar_wait(bar, phase);
        phase ^= 1;
        if (tid == 0 && rep == 1) { g_out[2 * slot] = (unsigned long long)(t1 - t0); g_out[2 * slot + 1] = (unsigned long long)(clock64() - t0); }
        tc5::fence_before_sync();
        __syncthreads();
    }
}

__global__ void __launch_bounds__(128, 1) bench(int R) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *ctrl = smem, *a = smem + 1024, *b = a + 65536;
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t bar = smem_u32(ctrl);
    if (tid == 0) { tc5::mbar_init(bar, 1); tc5::mbar_init_fence(); }
    __syncwarp();
    if (warp == 0) tc5::tmem_alloc(smem_u32(ctrl + 16), 512);
    for (int i = tid; i < 131072 / 16; i += 128) *reinterpret_cast<uint4 *>(a + i * 16) = make_uint4(0, 0, 0, 0);
    tc5::fence_proxy_async(); tc5::fence_before_sync(); __syncthreads(); tc5::fence_after_sync();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t *>(ctrl + 16);
    uint32_t phase = 0;
    const uint32_t a_u = smem_u32(a), b_u = smem_u32(b);
    run<0, 32>(tmem, a_u, b_u, bar, phase, 0, R);
    run<1, 32>(tmem, a_u, b_u, bar, phase, 1, R);
    run<2, 32>(tmem, a_u, b_u, bar, phase, 2, R);
    run<3, 32>(tmem, a_u, b_u, bar, phase, 3, R);
    run<3, 64>(tmem, a_u, b_u, bar, phase, 4, R);
    run<3, 128>(tmem, a_u, b_u, bar, phase, 5, R);
    run<3, 256>(tmem, a_u, b_u, bar, phase, 6, R);
    run<3, 144>(tmem, a_u, b_u, bar, phase, 7, R);
    run<3, 16>(tmem, a_u, b_u, bar, phase, 8, R);
    run<3, 32, 1>(tmem, a_u, b_u, bar, phase, 9, R);
    run<3, 16, 1>(tmem, a_u, b_u, bar, phase, 10, R);
    run<3, 64, 1>(tmem, a_u, b_u, bar, phase, 11, R);
    run<3, 32, 2>(tmem, a_u, b_u, bar, phase, 12, R);
    run<3, 256, 1>(tmem, a_u, b_u, bar, phase, 13, R);
    if (warp == 0) tc5::tmem_free(tmem, 512);
}

int main() {
    const int R = 64;
    size_t smem = 1024 + 131072;
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    bench<<<1, 128, smem>>>(R);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    unsigned long long out[64]; cudaMemcpyFromSymbol(out, g_out, sizeof out);
    const char *names[] = {"style0 lane0 rolled N=32", "style1 elect per instr N=32", "style2 one elect unroll8 N=32", "style3 const-acc N=32",
                           "style3 N=64", "style3 N=128", "style3 N=256", "style3 N=144", "style3 N=16", "TS N=32", "TS N=16", "TS N=64", "alternating SS/TS N=32", "TS N=256"};
    for (int i = 0; i < 14; ++i) printf("%-34s issue %7.1f clk/instr   complete %7.1f clk/instr\n", names[i], (double)out[2 * i] / R, (double)out[2 * i + 1] / R);
    return 0;
}
