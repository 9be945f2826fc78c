// rs_umma_bench.cu -- cost of the exact G1 / G2 / G3 instruction sequences of wdsr_rs.cuh, issued alone (developer probe).
#include <cstdio>
#include "tc5.cuh"
#include "wdsr_rs_layout.cuh"
using namespace b200sr;
using namespace b200sr::rs;
__device__ unsigned long long g_out[64];

__global__ void __launch_bounds__(128, 1) bench(int reps) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *ctrl = smem, *xs = smem + 256, *t2 = xs + X_BYTES, *wsm = t2 + T2_BYTES;
    const BlockRsLayout L(144);
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t bar = smem_u32(ctrl);
    if (tid == 0) { tc5::mbar_init(bar, 1); for (int i = 1; i < 5; ++i) tc5::mbar_init(bar + 8 * i, 1); tc5::mbar_init_fence(); }
    __syncwarp();
    if (warp == 0) tc5::tmem_alloc(smem_u32(ctrl + 16), 512);
    for (int i = tid; i < (X_BYTES + T2_BYTES + L.total) / 16; i += 128) *reinterpret_cast<uint4 *>(xs + i * 16) = make_uint4(0, 0, 0, 0);
    tc5::fence_proxy_async(); tc5::fence_before_sync(); __syncthreads(); tc5::fence_after_sync();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t *>(ctrl + 16);
    const uint32_t xs_u = smem_u32(xs), t2_u = smem_u32(t2), w_u = smem_u32(wsm);
    uint32_t phase = 0;
    const uint32_t idesc1 = tc5::idesc_bf16_f32(128, 144), idesc32 = tc5::idesc_bf16_f32(128, 32), idesc64 = tc5::idesc_bf16_f32(128, 64),
                   idesc96 = tc5::idesc_bf16_f32(128, 96);
    const uint64_t bw1a = tc5::smem_desc(w_u + L.w1, 128, 512), bw1b = tc5::smem_desc(w_u + L.w1 + 256, 128, 512);
    const uint64_t bw2 = tc5::smem_desc(w_u + L.w2, 128, L.sbo2), bw3 = tc5::smem_desc(w_u + L.w3, 128, L.sbo3);
    const uint64_t grp = (uint64_t)((4 * L.sbo3) >> 4);
    auto g1 = [&](int s) {
        const int slot = s % NX, e = s & 1;
        const uint32_t base = xs_u + slot * XSLOT;
        tc5::mma_ss(tmem + d1_col(e), tc5::smem_desc(base, XPLANE, 128), bw1a, idesc1, false);
        tc5::mma_ss(tmem + d1_col(e), tc5::smem_desc(base + 2 * XPLANE, X_ONE - slot * XSLOT - 2 * XPLANE, 128), bw1b, idesc1, true);
    };
    auto g2 = [&](int s) {
        const int e = s & 1;
        const uint32_t d2 = tmem + d2_col(e), a2 = tmem + d1_col(e);
        tc5::mma_ts(d2, a2, bw2, idesc32, false);
#pragma unroll 4
        for (int j = 1; j < 9; ++j) tc5::mma_ts(d2, a2 + (j < 4 ? 8 * j : 104 + 8 * (j - 4)), bw2 + (uint64_t)(16 * j), idesc32, true);
    };
    // PACK slices: (c0|c1) at dx 0,1,2 ; (c2' dx0 | c2' dx2)
    auto aslice = [&](int s, int i, bool aligned) -> uint64_t {
        const uint32_t base = t2_u + (s % NT) * T2SLOT;
        if (aligned) return tc5::smem_desc(base + (i < 3 ? 0 : 2 * 2048), i < 3 ? 2048 : 128, 128);
        return i < 3 ? tc5::smem_desc(base + i * 16, T2PLANE, 128) : tc5::smem_desc(base + 2 * T2PLANE, 32, 128);
    };
    auto g3 = [&](int s, int a, bool aligned) {
        if (a <= 2) {
            for (int i = 0; i < 4; ++i) tc5::mma_ss(tmem + out_col(a), aslice(s, i, aligned), bw3 + (uint64_t)(16 * i), idesc96, true);
        } else if (a == 3) {
            for (int i = 0; i < 4; ++i) tc5::mma_ss(tmem + out_col(3), aslice(s, i, aligned), bw3 + (uint64_t)(16 * i), idesc64, true);
            for (int i = 0; i < 4; ++i) tc5::mma_ss(tmem + out_col(0), aslice(s, i, aligned), bw3 + 2 * grp + (uint64_t)(16 * i), idesc32, true);
        } else {
            for (int i = 0; i < 4; ++i) tc5::mma_ss(tmem + out_col(4), aslice(s, i, aligned), bw3 + (uint64_t)(16 * i), idesc32, true);
            for (int i = 0; i < 4; ++i) tc5::mma_ss(tmem + out_col(0), aslice(s, i, aligned), bw3 + grp + (uint64_t)(16 * i), idesc64, true);
        }
    };
    {   // four issuer warps as in the kernel (A even / odd steps, B even / odd steps), no waits: pure tensor-queue time per step
        __syncthreads();
        const long long t0 = clock64();
        if (tc5::elect_one()) {
            for (int s = warp & 1; s < 2 * reps; s += 2) {
                if (warp < 2) { g2(s); g1(s + 2); }
                else g3(s, (s + 4) % 5, false);
            }
            tc5::commit(bar + 8 + 8 * warp);
        }
        __syncwarp();
        tc5::mbar_wait(bar + 8 + 8 * warp, 0);
        const long long t1 = clock64();
        __syncthreads();
        if (tid == 0) g_out[40] = (unsigned long long)(t1 - t0);
        if (tid == 64) g_out[41] = (unsigned long long)(t1 - t0);
        // A issuers alone
        const long long t2 = clock64();
        if (warp < 2) {
            if (tc5::elect_one()) {
                for (int s = warp & 1; s < 2 * reps; s += 2) { g2(s); g1(s + 2); }
                tc5::commit(bar + 8 + 8 * warp);
            }
            __syncwarp();
            tc5::mbar_wait(bar + 8 + 8 * warp, 1);
        }
        const long long t3 = clock64();
        __syncthreads();
        if (tid == 0) g_out[42] = (unsigned long long)(t3 - t2);
        const long long t4 = clock64();
        if (warp >= 2) {
            if (tc5::elect_one()) {
                for (int s = warp & 1; s < 2 * reps; s += 2) g3(s, (s + 4) % 5, false);
                tc5::commit(bar + 8 + 8 * warp);
            }
            __syncwarp();
            tc5::mbar_wait(bar + 8 + 8 * warp, 1);
        }
        const long long t5 = clock64();
        __syncthreads();
        if (tid == 64) g_out[43] = (unsigned long long)(t5 - t4);
    }
    for (int mode = 0; mode < 9; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            long long t0 = 0, t1 = 0;
            if (warp == 0) {
                tc5::fence_after_sync();
                t0 = clock64();
                if (tc5::elect_one()) {
                    for (int s = 0; s < reps; ++s) {
                        if (mode == 0) g1(s);
                        else if (mode == 1) g2(s);
                        else if (mode == 2) g3(s, 1, false);
                        else if (mode == 3) g3(s, 1, true);
                        else if (mode == 4) g3(s, 3, false);
                        else if (mode == 5) g3(s, 4, false);
                        else if (mode == 6) { g2(s); g1(s + 2); g3(s, (s + 4) % 5, false); }
                        else if (mode == 7) { g2(s); g1(s + 2); }
                        else { for (int i = 0; i < 4; ++i) tc5::mma_ss(tmem + out_col(0), aslice(s, 0, true), bw3, idesc96, true); }
                    }
                    tc5::commit(bar);
                }
                t1 = clock64();
                __syncwarp();
            }
            if (warp != 0 || tid == 0) tc5::mbar_wait(bar, phase);
            phase ^= 1;
            if (tid == 0 && rep == 1) { g_out[2 * mode] = (unsigned long long)(t1 - t0); g_out[2 * mode + 1] = (unsigned long long)(clock64() - t0); }
            tc5::fence_before_sync();
            __syncthreads();
        }
    }
    if (warp == 0) tc5::tmem_free(tmem, 512);
}

int main() {
    const int reps = 20;
    size_t smem = 256 + X_BYTES + T2_BYTES + BlockRsLayout(144).total;
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    bench<<<1, 128, smem>>>(reps);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    unsigned long long out[64]; cudaMemcpyFromSymbol(out, g_out, sizeof out);
    const char *names[] = {"G1 (2 SS N=144)", "G2 (9 TS N=32)", "G3 non-wrap (4 SS N=96, real A)", "G3 non-wrap, 128B-aligned A", "G3 wrap a=3 (4 N=64 + 4 N=32)",
                           "G3 wrap a=4 (4 N=32 + 4 N=64)", "full step mix G2+G1+G3 (kernel order, one thread)", "G2+G1 only", "4 SS N=96 same aligned A"};
    const double cnt[] = {2, 9, 4, 4, 8, 8, 15 + 8.0 * 2 / 5, 11, 4};
    printf("four issuer warps, no waits: %.0f clk per step (A view) %.0f (B view); A issuers alone %.0f; B issuers alone %.0f\n", (double)out[40] / (2 * reps),
           (double)out[41] / (2 * reps), (double)out[42] / (2 * reps), (double)out[43] / (2 * reps));
    for (int i = 0; i < 9; ++i)
        printf("%-52s per step: issue %6.0f  complete %6.0f clk   (%.1f clk/instr)\n", names[i], (double)out[2 * i] / reps, (double)out[2 * i + 1] / reps,
               (double)out[2 * i + 1] / reps / cnt[i]);
    return 0;
}
