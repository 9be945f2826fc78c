// wdsr_tc5_tail.cu -- launcher of the tcgen05 fused tail kernel.
#include "launch.h"
#include "tma_map.h"
#include "wdsr_tc5_head.cuh"
#include "wdsr_tc5_tail.cuh"

namespace b200sr {

template <typename TIN, typename TOUT, int S>
static cudaError_t tail_tc5_t(const void *trunk, const void *x, void *y, const uint8_t *wimg, int N, int H, int W, float mean, float out_add,
                              cudaStream_t st) {
    using namespace tc5tail;
    constexpr int NOP = round_up(3 * S * S, 16);
    struct MapKey { const void *p; int n, h, w; CUtensorMap map; };
    static thread_local MapKey cache[4];
    static thread_local int next_slot = 0;
    const CUtensorMap *mapp = nullptr;
    for (auto &c : cache)
        if (c.p == trunk && c.n == N && c.h == H && c.w == W) mapp = &c.map;
    cudaError_t e;
    if (!mapp) {
        MapKey &c = cache[next_slot++ & 3];
        e = make_trunk_map(&c.map, trunk, N, H, W, TW, TH + 2);
        if (e != cudaSuccess) { c.p = nullptr; return e; }
        c.p = trunk, c.n = N, c.h = H, c.w = W;
        mapp = &c.map;
    }
    auto kern = wdsr_tail_tc5_kernel<TIN, TOUT, S>;
    const size_t smem = smem_bytes(NOP);
    static thread_local SmemOptIn optin;   // per device (launch.h)
    if ((e = optin.ensure(kern, smem)) != cudaSuccess) return e;
    const int tx = ceil_div(W, TW), ty = ceil_div(H, TH), ntiles = tx * ty * N;
    int ctas = tc5tail::CTAS_PER_SM * sm_count();
    if (ctas > ntiles) ctas = ntiles;
    kern<<<ctas, NTHREADS, smem, st>>>(*mapp, (const TIN *)x, (TOUT *)y, wimg, N, H, W, tx, ty, ntiles, mean, out_add);
    return cudaGetLastError();
}

template <int S>
static cudaError_t tail_tc5_io(int xd, int yd, const void *trunk, const void *x, void *y, const uint8_t *wimg, int N, int H, int W, float mean,
                               float out_add, cudaStream_t st) {
    if (xd == kF32 && yd == kF32) return tail_tc5_t<float, float, S>(trunk, x, y, wimg, N, H, W, mean, out_add, st);
    if (xd == kF32 && yd == kBF16) return tail_tc5_t<float, bf16, S>(trunk, x, y, wimg, N, H, W, mean, out_add, st);
    if (xd == kBF16 && yd == kF32) return tail_tc5_t<bf16, float, S>(trunk, x, y, wimg, N, H, W, mean, out_add, st);
    if (xd == kBF16 && yd == kBF16) return tail_tc5_t<bf16, bf16, S>(trunk, x, y, wimg, N, H, W, mean, out_add, st);
    if (xd == kF32 && yd == kU8) return tail_tc5_t<float, uint8_t, S>(trunk, x, y, wimg, N, H, W, mean, out_add, st);
    if (xd == kBF16 && yd == kU8) return tail_tc5_t<bf16, uint8_t, S>(trunk, x, y, wimg, N, H, W, mean, out_add, st);
    return cudaErrorInvalidValue;
}

cudaError_t launch_tail_tc5(int S, int xd, int yd, const void *trunk, const void *x, void *y, const uint8_t *wimg, int N, int H, int W,
                            float mean, float out_add, cudaStream_t st) {
    switch (S) {
        case 2: return tail_tc5_io<2>(xd, yd, trunk, x, y, wimg, N, H, W, mean, out_add, st);
        case 3: return tail_tc5_io<3>(xd, yd, trunk, x, y, wimg, N, H, W, mean, out_add, st);
        case 4: return tail_tc5_io<4>(xd, yd, trunk, x, y, wimg, N, H, W, mean, out_add, st);
    }
    return cudaErrorInvalidValue;
}


template <typename TIN>
static cudaError_t head_tc5_t(const void *x, void *trunk, const uint8_t *wimg, int N, int H, int W, float mean, cudaStream_t st) {
    using namespace tc5head;
    auto kern = wdsr_head_tc5_kernel<TIN>;
    const size_t smem = smem_bytes();
    static thread_local SmemOptIn optin;   // per device (launch.h); ~30 KB needs no opt-in today, kept for symmetry
    if (cudaError_t e = optin.ensure(kern, smem); e != cudaSuccess) return e;
    const int tx = ceil_div(W, TW), ty = ceil_div(H, TH), ntiles = tx * ty * N;
    int ctas = CTAS_PER_SM * sm_count();
    if (ctas > ntiles) ctas = ntiles;
    kern<<<ctas, NTHREADS, smem, st>>>((const TIN *)x, (bf16 *)trunk, wimg, N, H, W, tx, ty, ntiles, mean);
    return cudaGetLastError();
}

cudaError_t launch_head_tc5(int xd, const void *x, void *trunk, const uint8_t *wimg, int N, int H, int W, float mean, cudaStream_t st) {
    if (xd == kF32) return head_tc5_t<float>(x, trunk, wimg, N, H, W, mean, st);
    if (xd == kBF16) return head_tc5_t<bf16>(x, trunk, wimg, N, H, W, mean, st);
    return cudaErrorInvalidValue;
}

}  // namespace b200sr
