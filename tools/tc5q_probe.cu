// tc5q_probe.cu -- event timeline of the decoupled-staging tile form (developer tool; build: tools/mk.sh tc5q_probe tc5q_probe -DB200SR_TC5_PROF)
#include <cstdio>
#include <vector>
#include <cstring>
#include <algorithm>
#include <map>
#include "wdsr_tc5q.cuh"
#include "tma_map.h"
using namespace b200sr;
int main(int argc, char **argv) {
    const int N = 64, H = 96, W = 96, M1P = 144;
    const int which = argc > 1 ? atoi(argv[1]) : 1;   // 0: tc5p, 1: tc5q
    BlockTc5Layout L(M1P);
    std::vector<uint8_t> img(L.total, 0);
    uint8_t *dimg; bf16 *din, *dout;
    cudaMalloc(&dimg, L.total); cudaMemcpy(dimg, img.data(), L.total, cudaMemcpyHostToDevice);
    size_t nb = (size_t)N * H * W * 24 * 2;
    cudaMalloc(&din, nb); cudaMalloc(&dout, nb); cudaMemset(din, 0, nb);
    const int tx = ceil_div(W, 32), ty = ceil_div(H, 16), ntiles = tx * ty * N;
    CUtensorMap map; if (make_trunk_map(&map, din, N, H, W) != cudaSuccess) { printf("map failed\n"); return 1; }
    size_t smem = tc5v3::smem_bytes(M1P);
    auto kern = which ? wdsr_block_tc5q_kernel<3, 1> : wdsr_block_tc5p_kernel<3>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int rep = 0; rep < 4; ++rep) {
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a);
        kern<<<148, tc5v3::NTHREADS, smem>>>(map, din, dout, dimg, M1P, N, H, W, tx, ty, ntiles);
        cudaEventRecord(b);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("FAILED: %s\n", cudaGetErrorString(e)); return 1; }
        float ms; cudaEventElapsedTime(&ms, a, b);
        printf("%s rep %d: %.1f us\n", which ? "tc5q" : "tc5p", rep, ms * 1e3);
    }
#ifdef B200SR_TC5_PROF
    static unsigned long long ev[24][2048]; int evn[24];
    cudaMemcpyFromSymbol(ev, g_tc5p_evt, sizeof ev); cudaMemcpyFromSymbol(evn, g_tc5p_evtn, sizeof evn);
    for (int w : {1, 2, 3, 4, 8, 12, 16, 20}) {
        std::map<std::pair<int, int>, std::pair<double, int>> acc;
        for (int i = 12; i + 12 < evn[w]; ++i) {
            const int a = (int)(ev[w][i - 1] >> 48), b = (int)(ev[w][i] >> 48);
            const double d = (double)((ev[w][i] & 0xFFFFFFFFFFFFull) - (ev[w][i - 1] & 0xFFFFFFFFFFFFull));
            auto &p = acc[{a, b}]; p.first += d; p.second++;
        }
        printf("warp %2d (%d events):", w, evn[w]);
        for (auto &kv : acc) printf("  %d->%d: %.0f (x%d)", kv.first.first, kv.first.second, kv.second.first / kv.second.second, kv.second.second);
        printf("\n");
    }
    if (argc > 2) {
        struct E { unsigned long long t; int w, id; };
        std::vector<E> all;
        for (int w : {1, 2, 3, 4, 8, 12, 16, 20}) for (int i = 0; i < evn[w]; ++i) all.push_back({ev[w][i] & 0xFFFFFFFFFFFFull, w, (int)(ev[w][i] >> 48)});
        std::sort(all.begin(), all.end(), [](const E &a, const E &b) { return a.t < b.t; });
        std::vector<unsigned long long> ta;
        for (auto &e : all) if (e.w == 1 && e.id == 100) ta.push_back(e.t);
        const int first = atoi(argv[2]);
        if ((int)ta.size() > first + 1) {
            printf("timeline of one tile (warp 1 G1 issuer, 2 G3 issuer, 3 G2 issuer, 4 / 8 E1 halves, 12 / 16 E2, 20 E3)\n");
            for (auto &e : all) if (e.t >= ta[first] && e.t < ta[first + 1]) printf("  %6llu  warp %2d  evt %d\n", e.t - ta[first], e.w, e.id);
        }
    }
#endif
    return 0;
}
