// rh_probe.cu -- event timeline of the rh block kernel (developer tool; build: tools/mk.sh rh_probe rh_probe -DB200SR_RS_PROF)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cstring>
#include <algorithm>
#include <map>
#include "wdsr_rh.cuh"
#include "wdsr_rs_pack.h"
using namespace b200sr;
int main(int argc, char **argv) {
    const int N = argc > 1 ? atoi(argv[1]) : 64, H = argc > 2 ? atoi(argv[2]) : 96, W = argc > 3 ? atoi(argv[3]) : 96, M1 = 144, M2 = 20, M1P = 144;
    std::vector<float> w1(M1 * 24, 0.01f), b1(M1, 0.f), w2(M2 * M1, 0.01f), b2(M2, 0.f), w3(24 * M2 * 9, 0.01f), b3(24, 0.f);
    std::vector<uint8_t> img;
    pack_block_rs(img, 24, M1, M2, M1P, w1.data(), b1.data(), w2.data(), b2.data(), w3.data(), b3.data());
    uint8_t *dimg; bf16 *din, *dout;
    cudaMalloc(&dimg, img.size()); cudaMemcpy(dimg, img.data(), img.size(), cudaMemcpyHostToDevice);
    size_t nb = (size_t)N * H * W * 24 * 2;
    cudaMalloc(&din, nb); cudaMalloc(&dout, nb); cudaMemset(din, 0, nb);
    const int total_rows = rs::num_strips(N, W) * H;
    size_t smem = rh::smem_bytes(M1P);
    auto kern = wdsr_block_rh_kernel<3, true, 9>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int ctas = 148;
    if (ctas > (total_rows + 3) / 4) ctas = (total_rows + 3) / 4;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a);
        kern<<<ctas, rh::NTHREADS, smem>>>(din, dout, dimg, M1P, N, H, W, total_rows);
        cudaEventRecord(b);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("FAILED: %s\n", cudaGetErrorString(e)); return 1; }
        float ms; cudaEventElapsedTime(&ms, a, b);
        printf("rep %d: %.1f us, %d CTAs, %d rows (%.1f per CTA)\n", rep, ms * 1e3, ctas, total_rows, (double)total_rows / ctas);
    }
#ifdef B200SR_RS_PROF
    static unsigned long long ev[32][4096]; int evn[32];
    cudaMemcpyFromSymbol(ev, g_rs_evt, sizeof ev); cudaMemcpyFromSymbol(evn, g_rs_evtn, sizeof evn);
    // per warp: mean delta between consecutive events (id a -> id b), skipping the first and last 3 iterations
    for (int w : {0, 1, 2, 3, 4, 8, 12}) {
        std::map<std::pair<int, int>, std::pair<double, int>> acc;
        for (int i = 8; i + 8 < evn[w]; ++i) {
            const int a = (int)(ev[w][i - 1] >> 48), b = (int)(ev[w][i] >> 48);
            const double d = (double)((ev[w][i] & 0xFFFFFFFFFFFFull) - (ev[w][i - 1] & 0xFFFFFFFFFFFFull));
            auto &p = acc[{a, b}]; p.first += d; p.second++;
        }
        printf("warp %2d (%d events):", w, evn[w]);
        for (auto &kv : acc) printf("  %d->%d: %.0f clk (x%d)", kv.first.first, kv.first.second, kv.second.first / kv.second.second, kv.second.second);
        printf("\n");
    }
    if (argc > 4) {   // raw timeline of a window: argv[4] = first issuer-A iteration, argv[5] = iterations
        struct E { unsigned long long t; int w, id, k; };
        std::vector<E> all;
        for (int w : {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12}) { std::map<int, int> cnt; for (int i = 0; i < evn[w]; ++i) { int id = (int)(ev[w][i] >> 48); all.push_back({ev[w][i] & 0xFFFFFFFFFFFFull, w, id, cnt[id]++}); } }
        std::sort(all.begin(), all.end(), [](const E &a, const E &b) { return a.t < b.t; });
        std::vector<unsigned long long> ta;
        for (auto &e : all) if (e.w == 0 && e.id == 100) ta.push_back(e.t);
        const int first = atoi(argv[4]), n = argc > 5 ? atoi(argv[5]) : 2;
        if ((int)ta.size() > first + n) {
            const unsigned long long t0 = ta[first], t1 = ta[first + n];
            printf("timeline: warp 0/1 issuer A even/odd, 2 issuer B, 4 E1 even, 8 E1 odd, 12 E3, 16 loader; (k) = occurrence\n");
            for (auto &e : all) if (e.t >= t0 && e.t < t1) if (e.id == 100 || e.id == 102 || e.id == 300 || e.id == 301 || e.id == 303) printf("  %6llu  warp %2d  evt %d (%d)\n", e.t - t0, e.w, e.id, e.k);
        }
    }
#endif
    return 0;
}
