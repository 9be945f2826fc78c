// sync_cost.cu -- which synchronisation primitives stall behind a warp's OUTSTANDING GLOBAL STORES? (developer probe)
// One warp: [3 x st.global.v4 per lane to fresh lines] then primitive X, timed with clock64.  Compare with no stores before X.
#include <cstdio>
#include "tc5.cuh"
using namespace b200sr;
__device__ unsigned long long g_out[64];
__device__ unsigned g_sink;

__global__ void __launch_bounds__(128, 1) bench(uint4 *buf, size_t stride) {
    __shared__ __align__(16) unsigned long long bars[4];
    __shared__ uint32_t tptr;
    __shared__ __align__(16) uint4 scratch[128];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar = smem_u32(&bars[0]), bar2 = smem_u32(&bars[1]);
    if (tid == 0) { tc5::mbar_init(bar, 1); tc5::mbar_init(bar2, 32 * 1000000); tc5::mbar_init_fence(); }
    if (warp == 0) tc5::tmem_alloc(smem_u32(&tptr), 32);
    tc5::fence_before_sync(); __syncthreads(); tc5::fence_after_sync();
    const uint32_t tmem = tptr;
    if (tid == 0) tc5::mbar_arrive(bar);   // phase 0 of `bar` complete -> test/try_wait(parity 0) succeed immediately
    __syncthreads();
    if (warp == 0) {
        unsigned acc = 0;
        for (int with_st = 0; with_st < 2; ++with_st)
            for (int prim = 0; prim < 10; ++prim) {
                unsigned long long tot = 0;
                for (int r = 0; r < 16; ++r) {
                    __syncwarp();
                    // let earlier stores drain
                    for (volatile int s = 0; s < 3000; ++s) {}
                    if (with_st) {
                        uint4 *p = buf + ((size_t)(prim * 16 + r) * 3 * 32 + lane) ;
#pragma unroll
                        for (int q = 0; q < 3; ++q) p[q * 32 + (size_t)with_st * stride] = make_uint4(r, q, lane, prim);
                    }
                    const long long t0 = clock64();
                    switch (prim) {
                    case 0: acc += tc5::mbar_test(bar, 0); break;                       // test_wait (acquire)
                    case 1: tc5::mbar_wait(bar, 0); break;                              // try_wait (acquire)
                    case 2: tc5::mbar_arrive(bar2); break;                              // arrive.release.cta
                    case 3: tc5::mbar_arrive_relaxed(bar2); break;                      // arrive.relaxed
                    case 4: tc5::fence_proxy_async(); break;
                    case 5: tc5::fence_after_sync(); break;
                    case 6: tc5::fence_before_sync(); break;
                    case 7: { uint32_t v[16]; tc5::tmem_ld16(tmem + ((uint32_t)0 << 16), v); tc5::tmem_wait_ld(); acc += v[1]; } break;
                    case 8: { scratch[tid] = make_uint4(r, 1, 2, 3); acc += scratch[(tid + 1) & 31].x; } break;  // st.shared + ld.shared
                    case 9: acc += __any_sync(0xffffffffu, lane == r); break;
                    }
                    const long long t1 = clock64();
                    tot += (unsigned long long)(t1 - t0);
                }
                if (lane == 0) g_out[with_st * 10 + prim] = tot / 16;
            }
        if (acc == 0x1234567) g_sink = acc;
    }
    __syncthreads();
    if (warp == 0) tc5::tmem_free(tmem, 32);
}

int main() {
    uint4 *buf; size_t n = 1 << 22; cudaMalloc(&buf, n * sizeof(uint4) * 2);
    bench<<<1, 128>>>(buf, n);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    unsigned long long out[64]; cudaMemcpyFromSymbol(out, g_out, sizeof out);
    const char *names[] = {"mbarrier.test_wait (acquire)", "mbarrier.try_wait (acquire)", "mbarrier.arrive (release.cta)", "mbarrier.arrive.relaxed", "fence.proxy.async",
                           "tcgen05.fence::after_thread_sync", "tcgen05.fence::before_thread_sync", "tcgen05.ld x16 + wait::ld", "st.shared + ld.shared", "__any_sync"};
    for (int i = 0; i < 10; ++i) printf("%-36s alone %6llu clk    right after 3 x st.global.v4 %6llu clk\n", names[i], out[i], out[10 + i]);
    return 0;
}
