// hmma_bench.cu -- developer probe: throughput of the legacy warp-level mma.sync (HMMA.16816 bf16 -> fp32) on one SM, alone and
// next to a tcgen05.mma stream.  Question behind it: can the 1x1 -> ReLU -> 1x1 pair of the WDSR block live in registers
// (no TMEM round trip between the two GEMMs) while the 3x3 stays on tcgen05?
#include <cstdio>
#include "tc5.cuh"
using namespace b200sr;
__device__ unsigned long long g_out[64];

__device__ __forceinline__ void hmma(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// mode bit 0: HMMA warps run; bit 1: the tcgen05 issuer runs.  nw = HMMA warps (<= 16); warp 16 is the tcgen05 issuer.
__global__ void __launch_bounds__(544, 1) bench(int mode, int nw, int iters, int ummas, int nacc, int slot) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar = smem_u32(smem);
    if (tid == 0) { tc5::mbar_init(bar, 1); tc5::mbar_init_fence(); }
    if (warp == 0) tc5::tmem_alloc(smem_u32(smem + 16), 512);
    for (int i = tid; i < 60 * 1024 / 16; i += blockDim.x) *reinterpret_cast<uint4 *>(smem + 1024 + i * 16) = make_uint4(0, 0, 0, 0);
    tc5::fence_proxy_async(); tc5::fence_before_sync(); __syncthreads(); tc5::fence_after_sync();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t *>(smem + 16);
    const long long t0 = clock64();
    long long t1 = t0;
    if (warp < nw && (mode & 1)) {
        float c[6][4];
        uint32_t a[2][4], b[3][2];
        for (int i = 0; i < 6; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
        for (int i = 0; i < 2; ++i) for (int j = 0; j < 4; ++j) a[i][j] = 0x3c003c00u + lane + i + j;
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 2; ++j) b[i][j] = 0x3c003c00u + lane * 3 + i + j;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int k = 0; k < 9; ++k) {
#pragma unroll
                for (int m = 0; m < 2; ++m)
#pragma unroll
                    for (int n = 0; n < 3; ++n)
                        if (m * 3 + n < nacc) hmma(c[m * 3 + n], a[m], b[n]);
            }
        }
        float s = 0.f;
        for (int i = 0; i < 6; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
        if (s == 12345.f) g_out[63] = 1;
        t1 = clock64();
    } else if (warp == 16 && (mode & 2)) {
        const uint32_t idesc = tc5::idesc_bf16_f32(128, 96);
        const uint64_t ad = tc5::smem_desc(smem_u32(smem + 1024), 2048, 128), bd = tc5::smem_desc(smem_u32(smem + 1024 + 16384), 128, 256);
        if (tc5::elect_one()) {
            for (int i = 0; i < ummas; ++i) tc5::mma_ss(tmem + (i & 3) * 96, ad, bd, idesc, true);
            tc5::commit(bar);
        }
        __syncwarp();
        tc5::mbar_wait(bar, 0);
        t1 = clock64();
    }
    if (lane == 0 && warp <= 16) g_out[slot * 20 + warp] = (unsigned long long)(t1 - t0);
    tc5::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc5::tmem_free(tmem, 512);
}

// the register footprint of wdsr_rh.cuh's E1: A[2][9][4] and B[9][3][2] all distinct registers (slot 2 of g_out)
__global__ void __launch_bounds__(512, 1) bench_full(int nw, int iters, int two_sets) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long t0 = clock64();
    long long t1 = t0;
    if (warp < nw) {
        float c[6][4], c1[6][4];
        uint32_t a[2][9][4], b[9][3][2];
        for (int i = 0; i < 6; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f, c1[i][j] = 0.f;
        for (int m = 0; m < 2; ++m) for (int k = 0; k < 9; ++k) for (int j = 0; j < 4; ++j) a[m][k][j] = 0x3c003c00u + lane + m * 7 + k * 3 + j;
        for (int k = 0; k < 9; ++k) for (int n = 0; n < 3; ++n) for (int j = 0; j < 2; ++j) b[k][n][j] = 0x3c003c00u + lane * 3 + k * 5 + n + j;
#pragma unroll 1
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int k = 0; k < 9; ++k) {
#pragma unroll
                for (int m = 0; m < 2; ++m)
#pragma unroll
                    for (int n = 0; n < 3; ++n) {
                        if (two_sets && (k & 1)) hmma(c1[m * 3 + n], a[m][k], b[k][n]);
                        else hmma(c[m * 3 + n], a[m][k], b[k][n]);
                    }
            }
        }
        float s = 0.f;
        for (int i = 0; i < 6; ++i) for (int j = 0; j < 4; ++j) s += c[i][j] + c1[i][j];
        if (s == 12345.f) g_out[63] = 1;
        t1 = clock64();
    }
    if (lane == 0 && warp < 16) g_out[40 + warp] = (unsigned long long)(t1 - t0);
}

int main() {
    const size_t smem = 64 * 1024;
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    unsigned long long out[64];
    const int iters = 200, ummas = 2000;
    auto run = [&](int mode, int nw, int nacc, const char *what) {
        for (int rep = 0; rep < 2; ++rep) bench<<<1, 544, smem>>>(mode, nw, iters, ummas, nacc, 0);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: %s\n", what, cudaGetErrorString(e)); return; }
        cudaMemcpyFromSymbol(out, g_out, sizeof out);
        unsigned long long mx = 0;
        for (int w = 0; w < nw; ++w) mx = out[w] > mx ? out[w] : mx;
        const double nh = (double)iters * 9 * nacc;   // HMMAs per warp
        printf("%-46s", what);
        if (mode & 1) printf(" HMMA: %8llu clk, %5.2f clk per HMMA per SMSP (%d warps -> %d per SMSP), %6.0f MAC/clk/SM", mx, (double)mx / (nh * ((nw + 3) / 4)), nw, (nw + 3) / 4,
                             nh * nw * 2048.0 / (double)mx);
        if (mode & 2) printf("   tcgen05: %8llu clk, %5.1f clk per SS N=96 MMA", out[16], (double)out[16] / ummas);
        printf("\n");
    };
    run(1, 4, 6, "HMMA alone, 4 warps x 6 acc");
    run(1, 8, 6, "HMMA alone, 8 warps x 6 acc");
    run(1, 16, 6, "HMMA alone, 16 warps x 6 acc");
    run(1, 16, 3, "HMMA alone, 16 warps x 3 acc");
    run(1, 4, 1, "HMMA alone, 4 warps x 1 acc (dependent chain)");
    run(2, 0, 6, "tcgen05 alone");
    run(3, 4, 6, "both, 4 warps");
    run(3, 8, 6, "both, 8 warps");
    run(3, 16, 6, "both, 16 warps");
    for (int two = 0; two < 2; ++two)
        for (int nw : {4, 8, 16}) {
            for (int rep = 0; rep < 2; ++rep) bench_full<<<1, 512>>>(nw, iters, two);
            printf("%s ", cudaGetErrorString(cudaDeviceSynchronize()));
            cudaMemcpyFromSymbol(out, g_out, sizeof out);
            unsigned long long mx = 0;
            for (int w = 0; w < nw; ++w) mx = out[40 + w] > mx ? out[40 + w] : mx;
            printf("full footprint (126 operand registers), %2d warps, %d accumulator set(s): %5.2f clk per HMMA per SMSP\n", nw, two + 1, (double)mx / (iters * 54.0 * ((nw + 3) / 4)));
        }
    // whole chip: power / clock effects are not visible from one SM -- repeat "both" on every SM
    for (int rep = 0; rep < 2; ++rep) bench<<<148, 544, smem>>>(3, 16, iters * 10, ummas * 10, 6, 1);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    cudaMemcpyFromSymbol(out, g_out, sizeof out);
    printf("148 SMs, both, x10 work: HMMA %llu clk (%.2f clk per HMMA per SMSP), tcgen05 %llu clk (%.1f per MMA)\n", out[20], (double)out[20] / (iters * 10.0 * 54 * 4),
           out[36], (double)out[36] / (ummas * 10.0));
    return 0;
}
