// tc5p_time.cu -- plain timing of the production tcgen05 block kernel at cfg2 shape (no probes): 16 launches back to back
// (ping-pong trunk buffers, as in the model), median over repetitions.  Build variants with -D flags to A/B experiments.
#include <cstdio>
#include <vector>
#include <algorithm>
#include "wdsr_tc5p.cuh"
#include "tma_map.h"
using namespace b200sr;
int main(int argc, char **argv) {
    const int N = argc > 1 ? atoi(argv[1]) : 64, H = argc > 2 ? atoi(argv[2]) : 96, W = argc > 3 ? atoi(argv[3]) : 96, M1P = 144;
    BlockTc5Layout L(M1P);
    std::vector<uint8_t> img(L.total, 0);
    reinterpret_cast<float *>(img.data() + L.b2)[31] = 3.f;   // dense block: all three 8-channel chunks of t2 (see b200sr.cu)
    uint8_t *dimg; bf16 *buf[2];
    cudaMalloc(&dimg, L.total); cudaMemcpy(dimg, img.data(), L.total, cudaMemcpyHostToDevice);
    size_t nb = (size_t)N * H * W * 24 * 2;
    for (int i = 0; i < 2; ++i) { cudaMalloc(&buf[i], nb); cudaMemset(buf[i], 0, nb); }
    const int tx = ceil_div(W, 32), ty = ceil_div(H, 16), ntiles = tx * ty * N;
    CUtensorMap map[2];
    for (int i = 0; i < 2; ++i) if (make_trunk_map(&map[i], buf[i], N, H, W) != cudaSuccess) { printf("map failed\n"); return 1; }
    size_t smem = tc5v3::smem_bytes(M1P);
    cudaFuncSetAttribute(wdsr_block_tc5p_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int ctas = std::min(148, ntiles);
    std::vector<float> t;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int rep = 0; rep < 12; ++rep) {
        cudaEventRecord(a);
        for (int l = 0; l < 16; ++l)
        {
#ifdef NO_PDL
            wdsr_block_tc5p_kernel<3><<<ctas, tc5v3::NTHREADS, smem>>>(map[l & 1], buf[l & 1], buf[(l & 1) ^ 1], dimg, M1P, N, H, W, tx, ty, ntiles);
#else
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(ctas), cfg.blockDim = dim3(tc5v3::NTHREADS), cfg.dynamicSmemBytes = smem, cfg.stream = 0;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = attr, cfg.numAttrs = 1;
            cudaLaunchKernelEx(&cfg, wdsr_block_tc5p_kernel<3>, map[l & 1], (const bf16 *)buf[l & 1], buf[(l & 1) ^ 1], (const uint8_t *)dimg, M1P, N, H, W, tx, ty, ntiles);
#endif
        }
        cudaEventRecord(b);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("FAILED: %s\n", cudaGetErrorString(e)); return 1; }
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (rep >= 2) t.push_back(ms * 1e3f / 16);
    }
    std::sort(t.begin(), t.end());
    printf("%s: %dx%dx%d  %d tiles  block kernel median %.2f us/launch (min %.2f max %.2f)\n", argv[0], N, H, W, ntiles, t[t.size() / 2], t.front(), t.back());
    return 0;
}
