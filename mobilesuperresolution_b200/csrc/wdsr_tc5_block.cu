// wdsr_tc5_block.cu -- launchers of the tcgen05 fused residual-block kernels.
#include "launch.h"
#include "wdsr_tc5.cuh"

namespace b200sr {

cudaError_t launch_block_tc5(int variant, const void *in, void *out, const uint8_t *wimg, int M1P, int N, int H, int W,
                             cudaStream_t st) {
    using namespace tc5cfg;
    (void)variant;
    auto kern = wdsr_block_tc5_seq_kernel;
    const size_t smem = wdsr_block_tc5_seq_smem(M1P);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int tx = ceil_div(W, TW), ty = ceil_div(H, TH);
    const int ntiles = tx * ty * N;
    int ctas = sm_count();  // persistent, one CTA per SM (shared memory bound)
    if (ctas > ntiles) ctas = ntiles;
    kern<<<ctas, 128, smem, st>>>((const bf16 *)in, (bf16 *)out, wimg, M1P, N, H, W, tx, ty, ntiles);
    return cudaGetLastError();
}

}  // namespace b200sr
