// tc5p_probe.cu -- barrier-wait / stage timers of the production tcgen05 block kernel (developer tool).
#include <cstdio>
#include <vector>
#include <cstring>
#include <algorithm>
#include "wdsr_tc5p.cuh"
#include "tma_map.h"
using namespace b200sr;
int main() {
    const int N = 64, H = 96, W = 96, M1P = 144;
    BlockTc5Layout L(M1P);
    std::vector<uint8_t> img(L.total, 0);
    reinterpret_cast<float *>(img.data() + L.b2)[31] = 3.f;   // dense block: all three 8-channel chunks of t2 (see b200sr.cu)
    uint8_t *dimg; bf16 *din, *dout;
    cudaMalloc(&dimg, L.total); cudaMemcpy(dimg, img.data(), L.total, cudaMemcpyHostToDevice);
    size_t nb = (size_t)N * H * W * 24 * 2;
    cudaMalloc(&din, nb); cudaMalloc(&dout, nb); cudaMemset(din, 0, nb);
    const int tx = ceil_div(W, 32), ty = ceil_div(H, 16), ntiles = tx * ty * N;
    CUtensorMap map; if (make_trunk_map(&map, din, N, H, W) != cudaSuccess) { printf("map failed\n"); return 1; }
    size_t smem = tc5v3::smem_bytes(M1P);
    cudaFuncSetAttribute(wdsr_block_tc5p_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    unsigned *dbg_h = nullptr, *dbg_d = nullptr;
    cudaHostAlloc(&dbg_h, 148 * 32 * 16, cudaHostAllocMapped); memset(dbg_h, 0, 148 * 32 * 16); cudaHostGetDevicePointer(&dbg_d, dbg_h, 0);
    cudaMemcpyToSymbol(tc5::g_tc5_dbg, &dbg_d, sizeof dbg_d);
    for (int rep = 0; rep < 6; ++rep) {
        unsigned long long z[64] = {0};
        cudaMemcpyToSymbol(g_tc5p_prof, z, sizeof z);
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a);
        wdsr_block_tc5p_kernel<3><<<148, tc5v3::NTHREADS, smem>>>(map, din, dout, dimg, M1P, N, H, W, tx, ty, ntiles);
        cudaEventRecord(b);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            printf("FAILED: %s\n", cudaGetErrorString(e));
            for (int i = 0; i < 148 * 32; ++i) if ((dbg_h[4 * i] >> 16) == 0xDEAD) printf("  cta %u warp %u timed out on barrier byte-offset %u (index %u) parity %u\n", dbg_h[4 * i + 3], dbg_h[4 * i] & 0xffff, dbg_h[4 * i + 1] & 0xff, (dbg_h[4 * i + 1] & 0xff) / 8, dbg_h[4 * i + 2]);
            return 1;
        }
        float ms; cudaEventElapsedTime(&ms, a, b);
        unsigned long long p[64]; cudaMemcpyFromSymbol(p, g_tc5p_prof, sizeof p);
        const int t = (ntiles - 1) / 148 + 1;
        printf("v3 rep %d: %s, %.1f us; MMA warp per tile: loop %llu | waits A2_FULL %llu D2_EMPTY %llu XS_FULL %llu T2R_FULL %llu D3_EMPTY %llu\n",
               rep, cudaGetErrorString(e), ms * 1e3, p[5] / t, p[0] / t, p[1] / t, p[2] / t, p[3] / t, p[4] / t);
        printf("   CTA0 total %llu clk (=> %.0f MHz effective), setup %llu clk\n", p[56], p[56] / (ms * 1e3), p[57]);
        { static unsigned long long c[1024][3]; cudaMemcpyFromSymbol(c, g_tc5p_cta, sizeof c); unsigned long long t0 = ~0ull, t1 = 0; for (int i = 0; i < 148; ++i) { if (c[i][0] < t0) t0 = c[i][0]; if (c[i][1] > t1) t1 = c[i][1]; }
          double sumdur = 0, maxstart = 0, mindur = 1e18, maxdur = 0; for (int i = 0; i < 148; ++i) { double d = (double)(c[i][1] - c[i][0]); sumdur += d; if (d < mindur) mindur = d; if (d > maxdur) maxdur = d; double st = (double)(c[i][0] - t0); if (st > maxstart) maxstart = st; }
          printf("   CTAs: first start -> last end %.1f us; CTA duration min %.1f avg %.1f max %.1f us; latest CTA start +%.1f us; sm of cta0,1,147: %llu %llu %llu\n", (t1 - t0) * 1e-3, mindur * 1e-3, sumdur / 148 * 1e-3, maxdur * 1e-3, maxstart * 1e-3, c[0][2], c[1][2], c[147][2]); }
        printf("   MMA-B (G3) per tile: loop %llu | wait G3_READY %llu\n", p[45] / t, p[43] / t);
        for (int w = 0; w < 2; ++w) { // w=0: warp 4 (WG1, E1), w=1: warp 16 (WG4, E2 odd)
            const unsigned long long *q = p + 8 + 16 * w;
            printf("   warp%d per tile: waits D1_FULL %llu D2_FULL %llu T2R_FREE %llu D3_FULL %llu XS_FULL %llu | work E1 %llu E2 %llu E3 %llu\n", w,
                   q[0] / t, q[1] / t, q[2] / t, q[3] / t, q[4] / t, q[5] / t, q[6] / t, q[7] / t);
        }
        { const unsigned long long *q = p + 32; printf("   E3 warp per tile: waits D3_FULL %llu XS_FULL %llu | work E3 %llu\n", q[3] / t, q[4] / t, q[7] / t); }
    }
    {   // timeline of CTA 0 (last rep): events of warps 1 (MMA-A), 2 (MMA-B), 4 (WG1 E1 lo), 8 (WG2 E1 hi), 12 (WG3 E2 even), 16 (WG4 E2 odd), 20 (WG5 E3)
        static unsigned long long ev[24][2048]; int evn[24];
        cudaMemcpyFromSymbol(ev, g_tc5p_evt, sizeof ev); cudaMemcpyFromSymbol(evn, g_tc5p_evtn, sizeof evn);
        struct E { unsigned long long t; int w, id; };
        std::vector<E> all;
        const int ws[] = {1, 2, 4, 8, 12, 16, 20};
        for (int w : ws) for (int i = 0; i < evn[w]; ++i) all.push_back({ev[w][i] & 0xFFFFFFFFFFFFull, w, (int)(ev[w][i] >> 48)});
        std::sort(all.begin(), all.end(), [](const E &a, const E &b) { return a.t < b.t; });
        // print tiles 3..4 window: find the 4th occurrence of id 100 (G2 step m=0)
        int seen = 0; unsigned long long t0 = 0, t1 = 0;
        for (auto &e : all) if (e.w == 1 && e.id == 100) { ++seen; if (seen == 4) t0 = e.t; if (seen == 5) t1 = e.t; }
        printf("timeline of tile 3 (cycles since its first G2 step; tile period %llu):\n", t1 - t0);
        for (auto &e : all) if (e.t >= t0 && e.t < t1) printf("  %6llu  warp %2d  evt %d\n", e.t - t0, e.w, e.id);
    }
    return 0;
}
