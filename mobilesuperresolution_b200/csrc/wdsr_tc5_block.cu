// wdsr_tc5_block.cu -- launchers of the tcgen05 fused residual-block kernels.
#include <cuda.h>

#include <cstdlib>

#include <mutex>

#include "launch.h"
#include "wdsr_tc5.cuh"
#include "wdsr_tc5p.cuh"
#include "wdsr_tc5q.cuh"
#include "wdsr_tc5c.cuh"
#include "tma_map.h"

namespace b200sr {

bool block_tc5_g3_packed(int M2) {
    static const bool off = [] { const char *e = getenv("B200SR_G3_NOPACK"); return e && e[0] == '1'; }();   // developer A/B switch
    return !off && M2 > 16 && M2 <= 20;
}

cudaError_t launch_block_tc5(int variant, const void *in, void *out, const uint8_t *wimg, int M1P, int M2, int N, int H, int W,
                             cudaStream_t st) {
    using namespace tc5cfg;
    const int tx = ceil_div(W, TW), ty = ceil_div(H, TH);
    const int ntiles = tx * ty * N;
    int ctas = sm_count();  // persistent, one CTA per SM (shared memory bound)
    if (ctas > ntiles) ctas = ntiles;
    if (variant == 0) {
        const size_t smem = wdsr_block_tc5_seq_smem(M1P);
        cudaError_t e = cudaFuncSetAttribute(wdsr_block_tc5_seq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        wdsr_block_tc5_seq_kernel<<<ctas, 128, smem, st>>>((const bf16 *)in, (bf16 *)out, wimg, M1P, N, H, W, tx, ty, ntiles);
        return cudaGetLastError();
    }
    // host-side launch cost matters (the kernel runs ~30 us): tensor maps are cached per (pointer, shape) -- the forward
    // ping-pongs between two trunk buffers -- and the shared-memory opt-in is done once per device
    struct MapKey { const void *p; int n, h, w; CUtensorMap map; };
    static thread_local MapKey cache[8];
    static thread_local int next_slot = 0;
    const CUtensorMap *mapp = nullptr;
    for (auto &c : cache)
        if (c.p == in && c.n == N && c.h == H && c.w == W) mapp = &c.map;
    cudaError_t e;
    if (!mapp) {
        MapKey &c = cache[next_slot++ & 7];
        e = make_trunk_map(&c.map, in, N, H, W);
        if (e != cudaSuccess) { c.p = nullptr; return e; }
        c.p = in, c.n = N, c.h = H, c.w = W;
        mapp = &c.map;
    }
    const int nc2 = M2 <= 8 ? 1 : M2 <= 16 ? 2 : 3;   // 8-channel chunks of t2 (the image was packed for exactly these, b200sr.cu)
    // variant 1: wdsr_tc5p.cuh (expand accumulator re-used in place as the reduce operand); 2: wdsr_tc5q.cuh (decoupled staging halves)
    const bool q = variant == 2;
    // the decoupled form takes two staging slots per expand half where TMEM has room for them (3 M1P + 128 <= 512 columns: the pruned widths)
    const bool q2 = q && M1P <= 128;
    auto kern = q2 ? (nc2 == 3 ? wdsr_block_tc5q_kernel<3, 2> : nc2 == 2 ? wdsr_block_tc5q_kernel<2, 2> : wdsr_block_tc5q_kernel<1, 2>)
                : q ? (nc2 == 3 ? wdsr_block_tc5q_kernel<3, 1> : nc2 == 2 ? wdsr_block_tc5q_kernel<2, 1> : wdsr_block_tc5q_kernel<1, 1>)
                    : (nc2 == 3 ? (block_tc5_g3_packed(M2) ? wdsr_block_tc5p_kernel<4> : wdsr_block_tc5p_kernel<3>)
                                : nc2 == 2 ? wdsr_block_tc5p_kernel<2> : wdsr_block_tc5p_kernel<1>);
    const size_t smem = tc5v3::smem_bytes(M1P);
    static_assert(tc5v3::CTRL_BYTES == tc5v4::CTRL_BYTES, "the two forms share one shared-memory layout");
    static thread_local size_t smem_set[64][10] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    const int ki = (!q && nc2 == 3 && block_tc5_g3_packed(M2)) ? 9 : nc2 - 1 + (q2 ? 6 : q ? 3 : 0);
    if (dev < 0 || dev >= 64 || smem_set[dev][ki] < smem) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) smem_set[dev][ki] = smem;
    }
    const CUtensorMap &map = *mapp;
    // programmatic stream serialization: this grid may start (and run its prologue) while the previous kernel in the stream
    // drains; the kernel orders its first trunk load behind griddepcontrol.wait
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctas), cfg.blockDim = dim3(tc5v3::NTHREADS), cfg.dynamicSmemBytes = smem, cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr, cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, map, (const bf16 *)in, (bf16 *)out, wimg, M1P, N, H, W, tx, ty, ntiles);
}

// All `nlayers` blocks (same M1P / M2) in ONE persistent cooperative launch (wdsr_tc5c.cuh): layer l reads buf[l & 1], writes buf[(l + 1) & 1].
cudaError_t launch_block_chain_tc5(void *buf_a, void *buf_b, const uint8_t *const *wimgs, int nlayers, unsigned *gsync, int M1P, int M2, int N,
                                   int H, int W, cudaStream_t st) {
    using namespace tc5cfg;
    if (nlayers < 1 || nlayers > tc5v5::MAX_LAYERS) return cudaErrorInvalidValue;
    const int tx = ceil_div(W, TW), ty = ceil_div(H, TH);
    const int ntiles = tx * ty * N;
    int ctas = sm_count();
    if (ctas > ntiles) ctas = ntiles;
    struct MapKey { const void *p; int n, h, w; CUtensorMap map; };
    static thread_local MapKey cache[8];
    static thread_local int next_slot = 0;
    auto get_map = [&](const void *ptr, const CUtensorMap **out) -> cudaError_t {
        for (auto &c : cache)
            if (c.p == ptr && c.n == N && c.h == H && c.w == W) { *out = &c.map; return cudaSuccess; }
        MapKey &c = cache[next_slot++ & 7];
        cudaError_t e = make_trunk_map(&c.map, ptr, N, H, W);
        if (e != cudaSuccess) { c.p = nullptr; return e; }
        c.p = ptr, c.n = N, c.h = H, c.w = W;
        *out = &c.map;
        return cudaSuccess;
    };
    const CUtensorMap *ma = nullptr, *mb = nullptr;
    cudaError_t e;
    if ((e = get_map(buf_a, &ma)) != cudaSuccess) return e;
    const CUtensorMap map_a = *ma;   // (copied: the second lookup may evict the first entry)
    if ((e = get_map(buf_b, &mb)) != cudaSuccess) return e;
    const CUtensorMap map_b = *mb;
    const int nc2 = M2 <= 8 ? 1 : M2 <= 16 ? 2 : 3;
    auto kern = nc2 == 3 ? wdsr_chain_tc5_kernel<3> : nc2 == 2 ? wdsr_chain_tc5_kernel<2> : wdsr_chain_tc5_kernel<1>;
    const size_t smem = tc5v5::smem_bytes(M1P);
    static thread_local size_t smem_set[64][3] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || smem_set[dev][nc2 - 1] < smem) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) smem_set[dev][nc2 - 1] = smem;
    }
    tc5v5::ChainImages imgs = {};
    for (int i = 0; i < nlayers; ++i) imgs.img[i] = wimgs[i];
    if ((e = cudaMemsetAsync(gsync, 0, sizeof(unsigned), st)) != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctas), cfg.blockDim = dim3(tc5v5::NTHREADS), cfg.dynamicSmemBytes = smem, cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;   // the layer barrier needs every CTA resident
    attr[0].val.cooperative = 1;
    cfg.attrs = attr, cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, map_a, map_b, (bf16 *)buf_a, (bf16 *)buf_b, imgs, nlayers, gsync, M1P, N, H, W, tx, ty, ntiles);
}

}  // namespace b200sr
