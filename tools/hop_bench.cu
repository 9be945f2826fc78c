// hop_bench.cu -- latency of the hand-offs the fused block kernel is built from (developer probe):
//   A  issue 1 tcgen05.mma (SS N=32) + commit -> same thread's try_wait returns
//   B  same, TS N=32           C  G1-like (2 x SS N=144)
//   D  MMA + commit -> another warp (4 warps waiting with try_wait) -> tcgen05.ld x32 + wait -> per-warp elected arrive -> issuer's try_wait
//   E  like D, but the consumer warps poll with test_wait + __any_sync (+ nanosleep)
//   F  pure mbarrier ping-pong between two warps (no MMA)
#include <cstdio>
#include "tc5.cuh"
using namespace b200sr;
__device__ unsigned long long g_out[16];

__global__ void __launch_bounds__(704, 1) bench(int reps, int npoll, int poll_mode) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *ctrl = smem, *opa = smem + 1024;     // 64 KB of zeros as operands
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t b_full = smem_u32(ctrl), b_back = smem_u32(ctrl + 8);
    const uint32_t b_never = smem_u32(ctrl + 16);
    volatile int *stop = reinterpret_cast<volatile int *>(ctrl + 128);
    if (tid == 0) { tc5::mbar_init(b_full, 1); tc5::mbar_init(b_back, 4); tc5::mbar_init(b_never, 1); tc5::mbar_init(b_never + 8, 1); *stop = 0; tc5::mbar_init_fence(); }
    __syncwarp();
    if (warp == 0) tc5::tmem_alloc(smem_u32(ctrl + 64), 512);
    for (int i = tid; i < 65536 / 16; i += 704) *reinterpret_cast<uint4 *>(opa + i * 16) = make_uint4(0, 0, 0, 0);
    tc5::fence_proxy_async(); tc5::fence_before_sync(); __syncthreads(); tc5::fence_after_sync();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t *>(ctrl + 64);
    const uint32_t idesc32 = tc5::idesc_bf16_f32(128, 32), idesc144 = tc5::idesc_bf16_f32(128, 144);
    const uint64_t ad = tc5::smem_desc(smem_u32(opa), 128, 256), bd = tc5::smem_desc(smem_u32(opa) + 32768, 128, 256);
    uint32_t ph_full = 0, ph_back = 0;
    for (int mode = 0; mode < 6; ++mode) {
        __syncthreads();
        const bool two_hop = mode >= 3;
        if (warp == 0) {
            const bool leader = tc5::elect_one();
            const long long t0 = clock64();
            for (int r = 0; r < reps; ++r) {
                if (mode != 5) {
                    if (leader) {
                        if (mode == 1) tc5::mma_ts(tmem + 256, tmem, bd, idesc32, false);
                        else if (mode == 2) { tc5::mma_ss(tmem, ad, bd, idesc144, false); tc5::mma_ss(tmem, ad, bd, idesc144, true); }
                        else tc5::mma_ss(tmem + 256, ad, bd, idesc32, false);
                        tc5::commit(b_full);
                    }
                    __syncwarp();
                } else {
                    if (leader) tc5::mbar_arrive(b_full);
                    __syncwarp();
                }
                if (!two_hop) { tc5::mbar_wait(b_full, ph_full); ph_full ^= 1; tc5::fence_after_sync(); }
                else { tc5::mbar_wait(b_back, ph_back); ph_back ^= 1; tc5::fence_after_sync(); }
            }
            if (lane == 0) { g_out[mode] = (unsigned long long)(clock64() - t0) / reps; *stop = mode + 1; }
        } else if (warp >= 2 && two_hop) {   // 4 consumer warps (warps 2..5), TMEM lanes by warp % 4
            const uint32_t tb = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 256;
            unsigned acc = 0;
            for (int r = 0; r < reps; ++r) {
                if (mode == 4) {
                    while (!__any_sync(0xffffffffu, tc5::mbar_test(b_full, ph_full))) __nanosleep(40);
                }
                tc5::mbar_wait(b_full, ph_full); ph_full ^= 1;
                if (mode != 5) {
                    tc5::fence_after_sync();
                    uint32_t v[32]; tc5::tmem_ld32(tb, v); tc5::tmem_wait_ld(); acc += v[5];
                    tc5::fence_before_sync();
                }
                __syncwarp();
                if (lane == 0) tc5::mbar_arrive(b_back);
            }
            if (acc == 0x12345) g_out[15] = acc;
        }
        else if (warp >= 6 && warp < 6 + npoll) {
            // pollers: spin on a barrier that never completes, the way the kernel's waiting warps do
            const uint32_t bn = b_never + 8 * (warp & 1);
            while (*stop != mode + 1) {
                if (poll_mode == 0) { for (int i = 0; i < 4; ++i) tc5::mbar_try_wait(bn, 0); }                       // all 32 lanes try_wait
                else if (poll_mode == 1) { if (lane == 0) for (int i = 0; i < 4; ++i) tc5::mbar_try_wait(bn, 0); __syncwarp(); }   // one lane
                else { for (int i = 0; i < 4; ++i) { __any_sync(0xffffffffu, tc5::mbar_test(bn, 0)); __nanosleep(40); } }          // test + any + sleep
            }
        }
        if (two_hop && warp < 2) { ph_full = (ph_full + reps) & 1; }          // keep every thread's parity bookkeeping in step
        if (!two_hop && warp >= 2) { ph_full = (ph_full + reps) & 1; }
        if (two_hop && warp >= 1) { ph_back = (ph_back + reps) & 1; }
        if (warp == 1 && two_hop) {}
    }
    __syncthreads();
    if (warp == 0) tc5::tmem_free(tmem, 512);
}

int main() {
    const int reps = 64;   // even: parities return to 0 after every mode
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 1024 + 65536);
  for (int poll_mode = 0; poll_mode < 3; ++poll_mode)
  for (int npoll = 0; npoll <= 16; npoll += 8) {
    if (npoll == 0 && poll_mode > 0) continue;
    bench<<<1, 704, 1024 + 65536>>>(reps, npoll, poll_mode);
    printf("---- %d poller warps, mode %s: %s\n", npoll, poll_mode == 0 ? "32-lane try_wait" : poll_mode == 1 ? "1-lane try_wait" : "test_wait+any+nanosleep", cudaGetErrorString(cudaDeviceSynchronize()));
    unsigned long long out[16]; cudaMemcpyFromSymbol(out, g_out, sizeof out);
    const char *names[] = {"A  1 x SS N=32 + commit -> own wait", "B  1 x TS N=32 + commit -> own wait", "C  2 x SS N=144 + commit -> own wait",
                           "D  MMA+commit -> 4 warps try_wait -> ld x32 -> arrive -> issuer", "E  as D, consumers poll test_wait+any_sync+nanosleep(40)",
                           "F  mbarrier ping-pong only (arrive -> 4 warps -> arrive back)"};
    for (int i = 0; i < 6; ++i) printf("%-66s %6llu clk per round trip\n", names[i], out[i]);
  }
    return 0;
}
