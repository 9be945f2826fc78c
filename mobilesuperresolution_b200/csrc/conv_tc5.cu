// conv_tc5.cu -- launcher of the tcgen05 3x3 64 -> 64 convolution (BasicVSR trunks, conv_hr).
#include "conv_tc5.cuh"

#include "launch.h"
#include "tma_map.h"

namespace b200sr {

// NHWC bf16 activation window (64 channels starting at a 16-byte aligned offset inside pixels of `cs` channels) viewed as
// 5-D (8 channels, 8 chunks, W, H, N); a box {8, 1, 32, 10, 1} lands as [row][pixel][16 B]; out-of-image pixels read as zero.
static cudaError_t make_nhwc64_map(CUtensorMap *map, const void *base, int N, int H, int W, int cs) {
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return cudaErrorNotSupported;
    const cuuint64_t dims[5] = {8, 8, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    const cuuint64_t strides[4] = {16, (cuuint64_t)cs * 2, (cuuint64_t)W * cs * 2, (cuuint64_t)H * W * cs * 2};
    const cuuint32_t box[5] = {8, 1, (cuuint32_t)tc5conv::BW, (cuuint32_t)tc5conv::BH, 1};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void *>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

bool conv_tc5_eligible(const ConvArgs &a) {
    return a.cin == 64 && a.cout == 64 && a.shuffle == 1 && a.x_cs % 8 == 0 && a.x_co % 8 == 0 && a.y_cs % 8 == 0 && a.y_co % 8 == 0 &&
           (!a.residual || (a.r_cs % 8 == 0 && a.r_co % 8 == 0)) && (reinterpret_cast<uintptr_t>(a.x) & 15) == 0 &&
           (reinterpret_cast<uintptr_t>(a.y) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.residual) & 15) == 0;
}

cudaError_t launch_conv3x3_c64_tc5(const ConvArgs &a, const uint8_t *wimg, cudaStream_t st) {
    using namespace tc5conv;
    // the recurrent trunks cycle through a handful of activation buffers: tensor maps are cached per (pointer, geometry)
    struct MapKey { const void *p; int n, h, w, cs; CUtensorMap map; };
    constexpr int NCACHE = 32;
    static thread_local MapKey cache[NCACHE];
    static thread_local int next_slot = 0;
    const void *base = reinterpret_cast<const bf16 *>(a.x) + a.x_co;
    const CUtensorMap *mapp = nullptr;
    for (auto &c : cache)
        if (c.p == base && c.n == a.n && c.h == a.h && c.w == a.w_ && c.cs == a.x_cs) { mapp = &c.map; break; }
    cudaError_t e;
    if (!mapp) {
        MapKey &c = cache[next_slot++ % NCACHE];
        e = make_nhwc64_map(&c.map, base, a.n, a.h, a.w_, a.x_cs);
        if (e != cudaSuccess) { c.p = nullptr; return e; }
        c.p = base, c.n = a.n, c.h = a.h, c.w = a.w_, c.cs = a.x_cs;
        mapp = &c.map;
    }
    static thread_local bool set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !set[dev]) {
        e = cudaFuncSetAttribute(conv3x3_c64_tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes());
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) set[dev] = true;
    }
    const int tx = ceil_div(a.w_, TWO), ty = ceil_div(a.h, TH), ntiles = tx * ty * a.n;
    int ctas = sm_count();
    if (ctas > ntiles) ctas = ntiles;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctas), cfg.blockDim = dim3(NTHREADS), cfg.dynamicSmemBytes = smem_bytes(), cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr, cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, conv3x3_c64_tc5_kernel, *mapp, a, wimg, tx, ty, ntiles);
}

}  // namespace b200sr
