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

// The tcgen05 path keeps the bf16 trunk in the planar-8 layout [N][3 planes][H][W][8 channels] ("NC/8HW8"): one 8-channel plane
// of a tile + halo is then box_h rows of box_w * 16 CONTIGUOUS bytes.  The tensor map views it as 4-D (4*W uint32, H, 3, N); a box
// {4 * box_w, box_h, 1, 1} lands as [row][px][16 B] in shared memory, out-of-image elements read as zero.
// (The first form used the NHWC trunk through a 5-D map with an 8-element inner box: the same shared-memory image, but every
//  16-byte pixel row was its own TMA request -- 1,836 per block tile, 2,880 per tail tile -- and the tail kernel was bound by it.)
static cudaError_t make_trunk_map(CUtensorMap *map, const void *trunk, int N, int H, int W, int box_w = tc5cfg::HW_, int box_h = tc5cfg::HH_) {
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return cudaErrorNotSupported;
    const cuuint64_t dims[4] = {(cuuint64_t)W * 4, (cuuint64_t)H, 3, (cuuint64_t)N};
    const cuuint64_t strides[3] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)H * W * 48};
    const cuuint32_t box[4] = {(cuuint32_t)box_w * 4, (cuuint32_t)box_h, 1, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, const_cast<void *>(trunk), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

}  // namespace b200sr
