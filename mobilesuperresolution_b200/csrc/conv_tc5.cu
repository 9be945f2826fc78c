// conv_tc5.cu -- launcher of the tcgen05 3x3 (64 | 65..80) -> 64 convolution (BasicVSR trunks, conv_hr).
#include "conv_tc5.cuh"

#include <cstdlib>

#include "launch.h"
#include "tma_map.h"

namespace b200sr {

// NHWC bf16 activation window (cin channels starting at a 16-byte aligned offset inside pixels of `cs` channels) viewed as
// 4-D (cin, W, H, N); a box {8, 32, 10, 1} at channel 8c lands as [row][pixel][16 B]; out-of-image pixels and channels >= cin
// read as zero.
static cudaError_t make_nhwc_map(CUtensorMap *map, const void *base, int N, int H, int W, int cs, int cin, int bh) {
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return cudaErrorNotSupported;
    const cuuint64_t dims[4] = {(cuuint64_t)cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    const cuuint64_t strides[3] = {(cuuint64_t)cs * 2, (cuuint64_t)W * cs * 2, (cuuint64_t)H * W * cs * 2};
    const cuuint32_t box[4] = {8, (cuuint32_t)tc5conv::BW, (cuuint32_t)bh, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

// planar-8 activation [N][8 planes][H][W][8 channels] viewed as 4-D (4 * W uint32, H, 8, N); a box {128, 10, 1, 1} is ten rows of
// 512 contiguous bytes and lands as the same [row][pixel][16 B] image.
static cudaError_t make_planar_map(CUtensorMap *map, const void *base, int N, int H, int W, int bh) {
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return cudaErrorNotSupported;
    const cuuint64_t dims[4] = {(cuuint64_t)W * 4, (cuuint64_t)H, 8, (cuuint64_t)N};
    const cuuint64_t strides[3] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)H * W * 128};
    const cuuint32_t box[4] = {(cuuint32_t)tc5conv::BW * 4, (cuuint32_t)bh, 1, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, const_cast<void *>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

bool conv_tc5_eligible(const ConvArgs &a) {
    const auto al16 = [](const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    if (a.ks == 1)   // 1x1 form: 128 -> 64 k (the fusion conv on cat([backward, forward] features)), NHWC input
        return a.cin == 128 && a.cout % 64 == 0 && a.cout <= 256 && a.shuffle == 1 && !a.base && !a.x_planar && al16(a.x) && al16(a.y) &&
               al16(a.residual) && a.x_cs % 8 == 0 && a.x_co % 8 == 0 && (!a.residual || (a.r_cs % 8 == 0 && a.r_co % 8 == 0)) &&
               (a.y_planar ? a.cout == 64 : (a.y_cs % 8 == 0 && a.y_co % 8 == 0));
    if (a.ks != 3) return false;
    if (a.base)   // "rgb" form: 64 -> 3, fp32 NCHW output + bilinear x4 base of an (H/4, W/4) image
        return a.cout == 3 && a.cin == 64 && a.shuffle == 1 && !a.residual && a.act == 0 && a.h % 4 == 0 && a.w_ % 4 == 0 && al16(a.x) &&
               (a.x_planar || (a.x_cs % 8 == 0 && a.x_co % 8 == 0));
    if (a.cout % 64 || a.cout > 256 || a.cin < 64 || a.cin > 80 || !al16(a.x) || !al16(a.y) || !al16(a.residual)) return false;
    if (a.shuffle == 2 ? a.residual != nullptr : a.shuffle != 1) return false;
    if (a.x_planar) {
        if (a.cin != 64 || (a.residual && a.cout != 64)) return false;
    } else if (a.x_cs % 8 || a.x_co % 8 || (a.residual && (a.r_cs % 8 || a.r_co % 8))) {
        return false;
    }
    return a.y_planar ? (a.shuffle == 2 ? a.cout % 32 == 0 : a.cout == 64) : (a.y_cs % 8 == 0 && a.y_co % 8 == 0);
}

static int grid_ctas(const ConvArgs &a, int G) {
    const int cap = a.max_ctas;   // two concurrent streams of small launches share the SMs better with half-size grids (b200sr.h)
    return (cap > 0 && cap < sm_count() ? cap : sm_count()) / G * G;
}

template <int NCH, int NOUT, int MT, int KS = 3>
static cudaError_t launch_t(const ConvArgs &a, const CUtensorMap &map, const uint8_t *wimg, cudaStream_t st) {
    using namespace tc5conv;
    using C = Cfg<NCH, NOUT, MT, KS>;
    auto kern = conv3x3_c64_tc5_kernel<NCH, NOUT, MT, KS>;
    static thread_local bool set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::smem_bytes());
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) set[dev] = true;
    }
    const int tx = ceil_div(a.w_, C::TWO), ty = ceil_div(a.h, C::TH), ntiles = tx * ty * a.n;
    const int G = NOUT == 64 ? a.cout / 64 : 1;      // output-channel groups: a CTA serves one (conv_tc5.cuh)
    int ctas = grid_ctas(a, G);
    if (ctas < G) ctas = G;
    if (ctas > ntiles * G) ctas = ntiles * G;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctas), cfg.blockDim = dim3(NTHREADS), cfg.dynamicSmemBytes = C::smem_bytes(), cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr, cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, map, a, wimg, tx, ty, ntiles);
}

cudaError_t launch_conv3x3_c64_tc5(const ConvArgs &a, const uint8_t *wimg, cudaStream_t st) {
    using namespace tc5conv;
    // tile height: 30 x 8 outputs normally; 30 x 4 when a CTA would get fewer than 6 of those (wave quantisation of small launches)
    const int G = a.base ? 1 : a.cout / 64;
    int ctas = grid_ctas(a, G) / G;
    if (ctas < 1) ctas = 1;
    const long long tiles8 = (long long)ceil_div(a.w_, 30) * ceil_div(a.h, 8) * a.n;
    const char *mt_env = getenv("B200SR_CONV_MT");   // developer / test switch (read per call): force 1 or 2 M-tiles per tile
    const int force_mt = mt_env ? atoi(mt_env) : 0;
    const int mt = a.cin != 64 || a.base || a.ks != 3 ? 2 : (force_mt == 1 || force_mt == 2) ? force_mt : (tiles8 < 6ll * ctas ? 1 : 2);
    const int bh = 4 * mt + 2 * (a.ks / 2);
    // the recurrent trunks cycle through a handful of activation buffers: tensor maps are cached per (pointer, geometry)
    struct MapKey { const void *p; int n, h, w, cs, cin, bh; CUtensorMap map; };
    constexpr int NCACHE = 32;
    static thread_local MapKey cache[NCACHE];
    static thread_local int next_slot = 0;
    const void *base = a.x_planar ? a.x : reinterpret_cast<const bf16 *>(a.x) + a.x_co;
    const int cs = a.x_planar ? -1 : a.x_cs;
    const CUtensorMap *mapp = nullptr;
    for (auto &c : cache)
        if (c.p == base && c.n == a.n && c.h == a.h && c.w == a.w_ && c.cs == cs && c.cin == a.cin && c.bh == bh) { mapp = &c.map; break; }
    if (!mapp) {
        MapKey &c = cache[next_slot++ % NCACHE];
        cudaError_t e = a.x_planar ? make_planar_map(&c.map, base, a.n, a.h, a.w_, bh) : make_nhwc_map(&c.map, base, a.n, a.h, a.w_, a.x_cs, a.cin, bh);
        if (e != cudaSuccess) { c.p = nullptr; return e; }
        c.p = base, c.n = a.n, c.h = a.h, c.w = a.w_, c.cs = cs, c.cin = a.cin, c.bh = bh;
        mapp = &c.map;
    }
    if (a.ks == 1) return launch_t<16, 64, 2, 1>(a, *mapp, wimg, st);
    if (a.base) return launch_t<8, 16, 2>(a, *mapp, wimg, st);
    if (a.cin != 64) return launch_t<10, 64, 2>(a, *mapp, wimg, st);
    return mt == 1 ? launch_t<8, 64, 1>(a, *mapp, wimg, st) : launch_t<8, 64, 2>(a, *mapp, wimg, st);
}

}  // namespace b200sr
