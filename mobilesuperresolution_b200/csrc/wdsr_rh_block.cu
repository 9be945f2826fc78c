// wdsr_rh_block.cu -- launcher of the row-streaming fused residual block with the reduce 1x1 in registers (wdsr_rh.cuh).
#include <cstdlib>

#include "launch.h"
#include "wdsr_rh.cuh"
#include "wdsr_rs_pack.h"

namespace b200sr {

cudaError_t launch_block_rh(const void *in, void *out, const uint8_t *wimg, int M1P, int M2, int N, int H, int W, cudaStream_t st) {
    const int nstrips = rs::num_strips(N, W);
    const int total_rows = nstrips * H;
    // one persistent CTA per SM; a CTA's range costs 2 extra t2 rows per unit, so tiny problems use fewer, longer ranges
    int ctas = sm_count();
    static const int cap = [] { const char *e = getenv("B200SR_RS_MAX_CTAS"); return e ? atoi(e) : 0; }();   // developer switch (timing sweeps)
    if (cap > 0 && ctas > cap) ctas = cap;
    const int min_rows = 4;
    if (ctas > (total_rows + min_rows - 1) / min_rows) ctas = (total_rows + min_rows - 1) / min_rows;
    const int nc2 = rs_nc2(M2);
    const bool pack = rs_pack20(M2);
    const bool dense = M1P == 144;
    void (*kern)(const bf16 *, bf16 *, const uint8_t *, int, int, int, int, int) =
        dense ? (nc2 == 1 ? wdsr_block_rh_kernel<1, false, 9> : nc2 == 2 ? wdsr_block_rh_kernel<2, false, 9>
                 : pack ? wdsr_block_rh_kernel<3, true, 9> : wdsr_block_rh_kernel<3, false, 9>)
              : (nc2 == 1 ? wdsr_block_rh_kernel<1, false, 0> : nc2 == 2 ? wdsr_block_rh_kernel<2, false, 0>
                 : pack ? wdsr_block_rh_kernel<3, true, 0> : wdsr_block_rh_kernel<3, false, 0>);
    const int ki = (nc2 == 3 ? (pack ? 3 : 2) : nc2 - 1) + (dense ? 4 : 0);
    const size_t smem = rh::smem_bytes(M1P);
    static thread_local size_t smem_set[64][8] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || smem_set[dev][ki] < smem) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) smem_set[dev][ki] = smem;
    }
    // programmatic stream serialization: the prologue overlaps the previous kernel's tail; the loaders order their first trunk copy
    // (and with it everything downstream, the output stores included) behind griddepcontrol.wait
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctas), cfg.blockDim = dim3(rh::NTHREADS), cfg.dynamicSmemBytes = smem, cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr, cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, (const bf16 *)in, (bf16 *)out, wimg, M1P, N, H, W, total_rows);
}

}  // namespace b200sr
