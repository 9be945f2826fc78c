// poll_cost.cu -- cost of probing an mbarrier whose phase is still PENDING (developer probe).
#include <cstdio>
#include "tc5.cuh"
using namespace b200sr;
__device__ unsigned long long g_out[16];
__global__ void bench() {
    __shared__ __align__(8) unsigned long long bars[2];
    const uint32_t bar = smem_u32(&bars[0]);
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { tc5::mbar_init(bar, 1); tc5::mbar_init_fence(); }
    __syncthreads();
    unsigned acc = 0;
    for (int mode = 0; mode < 6; ++mode) {
        unsigned long long tot = 0;
        for (int r = 0; r < 32; ++r) {
            __syncwarp();
            const long long t0 = clock64();
            if (mode == 0) acc += tc5::mbar_test(bar, 0);                                   // test_wait, 32 lanes, pending
            else if (mode == 1) { if (lane == 0) acc += tc5::mbar_test(bar, 0); }           // test_wait, 1 lane, pending
            else if (mode == 2) acc += tc5::mbar_try_wait(bar, 0);                          // try_wait, 32 lanes, pending (returns false after HW limit)
            else if (mode == 3) { if (lane == 0) acc += tc5::mbar_try_wait(bar, 0); }
            else if (mode == 4) acc += tc5::mbar_test(bar, 1);                              // test_wait on the COMPLETE parity (fresh barrier: parity 1 passes)
            else { volatile unsigned long long *p = &bars[0]; acc += (unsigned)(*p >> 32); } // plain ld.shared of the barrier word
            __syncwarp();
            tot += (unsigned long long)(clock64() - t0);
        }
        if (lane == 0) g_out[mode] = tot / 32;
    }
    if (acc == 0x1234567) g_out[15] = acc;
    if (threadIdx.x == 0) { g_out[8] = bars[0]; }
    __syncwarp();
    if (threadIdx.x == 0) { tc5::mbar_arrive(bar); g_out[9] = bars[0]; tc5::mbar_arrive(bar); g_out[10] = bars[0]; }
}
int main() {
    bench<<<1, 32>>>();
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    unsigned long long out[16]; cudaMemcpyFromSymbol(out, g_out, sizeof out);
    const char *names[] = {"test_wait pending, 32 lanes", "test_wait pending, 1 lane", "try_wait pending, 32 lanes", "try_wait pending, 1 lane", "test_wait complete, 32 lanes", "plain ld.shared.u64 of the barrier"};
    for (int i = 0; i < 6; ++i) printf("%-40s %6llu clk (incl. ~40 clk of timing overhead)\n", names[i], out[i]);
    printf("raw barrier word: init(1) %016llx  after 1st phase %016llx  after 2nd phase %016llx\n", out[8], out[9], out[10]);
    return 0;
}
