// conv7_tc5.cu -- launcher of the tcgen05 7x7 convolution (SPyNet BasicModule layers).
#include "conv7_tc5.cuh"

#include "launch.h"
#include "tma_map.h"

namespace b200sr {

// NHWC bf16 window (cin channels at a 16-byte aligned offset inside pixels of `cs` channels) as 4-D (cin, W, H, N), box {8, 32, 22, 1};
// planar-8 [N][planes][H][W][8] as 4-D (4 W uint32, H, planes, N), box {128, 22, 1, 1}.  Both land as [row][pixel][16 B].
static cudaError_t make_map7(CUtensorMap *map, const ConvArgs &a) {
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return cudaErrorNotSupported;
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const cuuint64_t N = a.n, H = a.h, W = a.w_;
    CUresult r;
    if (a.x_planar) {
        const cuuint64_t planes = a.cin / 8;
        const cuuint64_t dims[4] = {W * 4, H, planes, N};
        const cuuint64_t strides[3] = {W * 16, H * W * 16, H * W * 16 * planes};
        const cuuint32_t box[4] = {(cuuint32_t)tc5conv7::BW * 4, (cuuint32_t)tc5conv7::BH, 1, 1};
        r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, const_cast<void *>(a.x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
        const cuuint64_t cs = a.x_cs;
        const cuuint64_t dims[4] = {(cuuint64_t)a.cin, W, H, N};
        const cuuint64_t strides[3] = {cs * 2, W * cs * 2, H * W * cs * 2};
        const cuuint32_t box[4] = {8, (cuuint32_t)tc5conv7::BW, (cuuint32_t)tc5conv7::BH, 1};
        const void *base = reinterpret_cast<const bf16 *>(a.x) + a.x_co;
        r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

// (chunks of the input, output channels) the kernel is instantiated for: SPyNet's 8|16 -> 32, 32 -> 64, 64 -> 32, 32 -> 16
int conv7_tc5_nch(int cin) { return cin <= 16 ? 2 : cin <= 32 ? 4 : cin <= 64 ? 8 : 0; }
bool conv7_tc5_shape_ok(int cin, int cout) {
    const int nch = conv7_tc5_nch(cin);
    return (nch == 2 && cout == 32) || (nch == 4 && (cout == 64 || cout == 16)) || (nch == 8 && cout == 32);
}

bool conv7_tc5_eligible(const ConvArgs &a) {
    const auto al16 = [](const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    if (!conv7_tc5_shape_ok(a.cin, a.cout) || a.shuffle != 1 || a.residual || !al16(a.x) || !al16(a.y)) return false;
    if (a.x_planar ? (a.cin % 16 != 0) : (a.x_cs % 8 || a.x_co % 8)) return false;
    return a.y_planar || (a.y_cs % 8 == 0 && a.y_co % 8 == 0);
}

template <int NCH, int NOUT>
static cudaError_t launch7_t(const ConvArgs &a, const CUtensorMap &map, const uint8_t *wimg, cudaStream_t st) {
    using namespace tc5conv7;
    using C = Cfg<NCH, NOUT>;
    auto kern = conv7x7_tc5_kernel<NCH, NOUT>;
    static thread_local bool set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::smem_bytes());
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) set[dev] = true;
    }
    const int tx = ceil_div(a.w_, TWO), ty = ceil_div(a.h, TH), ntiles = tx * ty * a.n;
    int ctas = sm_count();
    if (ctas > ntiles) ctas = ntiles;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctas), cfg.blockDim = dim3(NTHREADS), cfg.dynamicSmemBytes = C::smem_bytes(), cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr, cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, map, a, wimg, tx, ty, ntiles);
}

cudaError_t launch_conv7x7_tc5(const ConvArgs &a, const uint8_t *wimg, cudaStream_t st) {
    struct MapKey { const void *p; int n, h, w, cs, cin; CUtensorMap map; };
    constexpr int NCACHE = 32;
    static thread_local MapKey cache[NCACHE];
    static thread_local int next_slot = 0;
    const void *base = a.x_planar ? a.x : reinterpret_cast<const bf16 *>(a.x) + a.x_co;
    const int cs = a.x_planar ? -1 : a.x_cs;
    const CUtensorMap *mapp = nullptr;
    for (auto &c : cache)
        if (c.p == base && c.n == a.n && c.h == a.h && c.w == a.w_ && c.cs == cs && c.cin == a.cin) { mapp = &c.map; break; }
    if (!mapp) {
        MapKey &c = cache[next_slot++ % NCACHE];
        cudaError_t e = make_map7(&c.map, a);
        if (e != cudaSuccess) { c.p = nullptr; return e; }
        c.p = base, c.n = a.n, c.h = a.h, c.w = a.w_, c.cs = cs, c.cin = a.cin;
        mapp = &c.map;
    }
    const int nch = conv7_tc5_nch(a.cin);
    if (nch == 2 && a.cout == 32) return launch7_t<2, 32>(a, *mapp, wimg, st);
    if (nch == 4 && a.cout == 64) return launch7_t<4, 64>(a, *mapp, wimg, st);
    if (nch == 4 && a.cout == 16) return launch7_t<4, 16>(a, *mapp, wimg, st);
    if (nch == 8 && a.cout == 32) return launch7_t<8, 32>(a, *mapp, wimg, st);
    return cudaErrorInvalidValue;
}

}  // namespace b200sr
