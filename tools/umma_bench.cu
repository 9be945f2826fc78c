// umma_bench.cu -- per-instruction cost of tcgen05.mma as a function of operand layout (developer probe).
#include <cstdio>
#include <vector>
#include "tc5.cuh"
using namespace b200sr;

struct Cfg { int ts, N, sbo_a, lbo_a, sbo_b, lbo_b, layout, nacc, M; const char *name; };

__device__ unsigned long long g_out[64];

__device__ __forceinline__ uint64_t desc_l(uint32_t saddr, uint32_t lbo, uint32_t sbo, int layout) {
    return tc5::smem_desc(saddr, lbo, sbo) | ((uint64_t)layout << 61);
}

__global__ void __launch_bounds__(128, 1) bench(const Cfg *cfgs, int ncfg, int R) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *ctrl = smem;
    uint8_t *a = smem + 1024;            // 64 KB operand A area
    uint8_t *b = a + 65536;              // 64 KB operand B area
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t bar = smem_u32(ctrl);
    if (tid == 0) { tc5::mbar_init(bar, 1); tc5::mbar_init_fence(); }
    __syncwarp();
    if (warp == 0) tc5::tmem_alloc(smem_u32(ctrl + 16), 512);
    for (int i = tid; i < 131072 / 16; i += 128) *reinterpret_cast<uint4 *>(a + i * 16) = make_uint4(0, 0, 0, 0);
    tc5::fence_proxy_async();
    tc5::fence_before_sync();
    __syncthreads();
    tc5::fence_after_sync();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t *>(ctrl + 16);
    uint32_t phase = 0;
    for (int c = 0; c < ncfg; ++c) {
        const Cfg k = cfgs[c];
        for (int rep = 0; rep < 2; ++rep) {
            long long t0 = 0;
            if (tid == 0) {
                const uint32_t idesc = tc5::idesc_bf16_f32(k.M, k.N);
                const uint64_t ad = desc_l(smem_u32(a), k.lbo_a, k.sbo_a, k.layout), bd = desc_l(smem_u32(b), k.lbo_b, k.sbo_b, k.layout);
                tc5::fence_after_sync();
                t0 = clock64();
                for (int r = 0; r < R; ++r) {
                    const uint32_t d = tmem + 256 + (r % k.nacc) * 32 * ((k.N + 31) / 32) % 256;
                    if (k.ts) tc5::mma_ts(d, tmem + (r % 8) * 8, bd, idesc, r >= k.nacc);
                    else tc5::mma_ss(d, ad, bd, idesc, r >= k.nacc);
                }
                tc5::commit(bar);
            }
            __syncwarp();
            if (warp != 0 || tid == 0) tc5::mbar_wait(bar, phase);   // lanes 1..31 of the issuing warp must NOT spin beside lane 0
            phase ^= 1;
            if (tid == 0 && rep == 1) g_out[c] = (unsigned long long)(clock64() - t0);
            tc5::fence_before_sync();
            __syncthreads();
        }
    }
    if (warp == 0) tc5::tmem_free(tmem, 512);
}

int main() {
    std::vector<Cfg> cfgs = {
        {0, 32, 384, 128, 3584, 128, 0, 1, 128, "SS N=32  none  A:sbo384 lbo128 (G3-like)"},
        {0, 32, 384, 1280, 3584, 128, 0, 1, 128, "SS N=32  none  A:sbo384 lbo1280"},
        {0, 32, 512, 128, 3584, 128, 0, 1, 128, "SS N=32  none  A:sbo512 lbo128"},
        {0, 32, 256, 128, 256, 128, 0, 1, 128, "SS N=32  none  dense sbo256 lbo128"},
        {0, 32, 128, 2048, 128, 2048, 0, 1, 128, "SS N=32  none  A:sbo128 lbo2048 (chunk-major)"},
        {0, 32, 384, 128, 3584, 128, 0, 2, 128, "SS N=32  none  G3-like, 2 accumulators"},
        {0, 144, 512, 128, 512, 128, 0, 1, 128, "SS N=144 none  sbo512 (G1-like)"},
        {0, 256, 256, 128, 256, 128, 0, 1, 128, "SS N=256 none  dense"},
        {1, 32, 0, 0, 2304, 128, 0, 1, 128, "TS N=32  none  B:sbo2304 (G2-like)"},
        {1, 32, 0, 0, 256, 128, 0, 1, 128, "TS N=32  none  B dense"},
        {1, 256, 0, 0, 256, 128, 0, 1, 128, "TS N=256 none  B dense"},
        {0, 32, 1024, 16, 1024, 16, 2, 1, 128, "SS N=32  sw128 sbo1024"},
        {0, 144, 1024, 16, 1024, 16, 2, 1, 128, "SS N=144 sw128"},
        {0, 256, 1024, 16, 1024, 16, 2, 1, 128, "SS N=256 sw128"},
        {0, 32, 512, 16, 512, 16, 4, 1, 128, "SS N=32  sw64  sbo512"},
        {0, 32, 256, 16, 256, 16, 6, 1, 128, "SS N=32  sw32  sbo256"},
        {0, 64, 384, 128, 3584, 128, 0, 1, 128, "SS N=64  none  G3-like"},
        {0, 32, 384, 128, 3584, 128, 0, 1, 64, "SS N=32  none  M=64"},
    };
    Cfg *d; cudaMalloc(&d, cfgs.size() * sizeof(Cfg)); cudaMemcpy(d, cfgs.data(), cfgs.size() * sizeof(Cfg), cudaMemcpyHostToDevice);
    const int R = 64;
    size_t smem = 1024 + 131072;
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    bench<<<1, 128, smem>>>(d, (int)cfgs.size(), R);
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(e));
    unsigned long long out[64]; cudaMemcpyFromSymbol(out, g_out, sizeof out);
    for (size_t i = 0; i < cfgs.size(); ++i) printf("%-52s %6.1f clk/instr (total %llu for %d)\n", cfgs[i].name, (double)out[i] / R, out[i], R);
    return 0;
}
