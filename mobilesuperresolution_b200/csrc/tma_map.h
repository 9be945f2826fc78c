// tma_map.h -- host-side construction of the TMA tensor map of the NHWC bf16 trunk (driver entry point fetched at run time).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <mutex>

#include "wdsr_tc5_layout.cuh"

namespace b200sr {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    });
    return fn;
}

// 5-D view of the NHWC bf16 trunk [N][H][W][24] as (8 ch, 3 chunks, W, H, N): a box {8, 1, 34, 18, 1} is one 8-channel
// plane of a tile + halo and lands contiguously ([row][px][16 B]) in shared memory; out-of-image elements read as zero.
static cudaError_t make_trunk_map(CUtensorMap *map, const void *trunk, int N, int H, int W, int box_w = tc5cfg::HW_, int box_h = tc5cfg::HH_) {
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return cudaErrorNotSupported;
    const cuuint64_t dims[5] = {8, 3, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    const cuuint64_t strides[4] = {16, 48, (cuuint64_t)W * 48, (cuuint64_t)H * W * 48};
    const cuuint32_t box[5] = {8, 1, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void *>(trunk), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

}  // namespace b200sr
