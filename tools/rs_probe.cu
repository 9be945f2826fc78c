// rs_probe.cu -- event timeline of the row-streaming tcgen05 block kernel (developer tool; build: tools/mk.sh rs_probe rs_probe -DB200SR_RS_PROF)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cstring>
#include <algorithm>
#include <initializer_list>
#include "wdsr_rs.cuh"
#include "wdsr_rs_pack.h"
using namespace b200sr;
int main(int argc, char **argv) {
    const int N = argc > 1 ? atoi(argv[1]) : 64, H = argc > 2 ? atoi(argv[2]) : 96, W = argc > 3 ? atoi(argv[3]) : 96, M1 = 144, M2 = 20, M1P = 144;
    const int first = argc > 4 ? atoi(argv[4]) : 14, nshow = argc > 5 ? atoi(argv[5]) : 3;
    std::vector<float> w1(M1 * 24, 0.01f), b1(M1, 0.f), w2(M2 * M1, 0.01f), b2(M2, 0.f), w3(24 * M2 * 9, 0.01f), b3(24, 0.f);
    std::vector<uint8_t> img;
    pack_block_rs(img, 24, M1, M2, M1P, w1.data(), b1.data(), w2.data(), b2.data(), w3.data(), b3.data());
    uint8_t *dimg; bf16 *din, *dout;
    cudaMalloc(&dimg, img.size()); cudaMemcpy(dimg, img.data(), img.size(), cudaMemcpyHostToDevice);
    size_t nb = (size_t)N * H * W * 24 * 2;
    cudaMalloc(&din, nb); cudaMalloc(&dout, nb); cudaMemset(din, 0, nb);
    const int total_rows = rs::num_strips(N, W) * H;
    size_t smem = rs::smem_bytes(M1P);
    auto kern = wdsr_block_rs_kernel<3, true>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int ctas = 148;
    if (ctas > (total_rows + 3) / 4) ctas = (total_rows + 3) / 4;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a);
        kern<<<ctas, rs::NTHREADS, smem>>>(din, dout, dimg, M1P, N, H, W, total_rows);
        cudaEventRecord(b);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("FAILED: %s\n", cudaGetErrorString(e)); return 1; }
        float ms; cudaEventElapsedTime(&ms, a, b);
        printf("rep %d: %.1f us, %d CTAs, %d rows (%.1f per CTA)\n", rep, ms * 1e3, ctas, total_rows, (double)total_rows / ctas);
    }
#ifdef B200SR_RS_TIMERS
    {
        unsigned long long tm[32][8]; cudaMemcpyFromSymbol(tm, g_rs_tm, sizeof tm);
        const double steps = (double)total_rows / ctas + 4;   // ~ steps of CTA 0
        auto row = [&](const char *name, int w, std::initializer_list<const char *> labels) {
            printf("  %-22s (warp %2d) per step:", name, w); int i = 0; double sum = 0;
            for (auto l : labels) { printf("  %s %.0f", l, tm[w][i] / steps); sum += tm[w][i]; ++i; }
            printf("  | other %.0f  total %.0f\n", tm[w][7] / steps, (sum + tm[w][7]) / steps);
        };
        printf("CTA 0 cycle accounting, clk per step (%.1f steps); issuers and E2 run every OTHER step (their per-iteration time is 2x):\n", steps);
        row("issuer even", 0, {"wait G2_READY", "wait G3_READY", "issue"});
        row("issuer odd", 1, {"wait G2_READY", "wait G3_READY", "issue"});
        for (int q = 0; q < 4; ++q) row("E1", 4 + 4 * q, {"wait D1_FULL", "tmem ld", "cvt+st+arrive", "issue copy+commit", "wait_group", "fence.proxy", "wait X_EMPTY"});
        row("E2 even", 20, {"wait D2_FULL", "ld+arrive", "math", "wait STEP_DONE", "st.shared+arrive"});
        row("E2 odd", 24, {"wait D2_FULL", "ld+arrive", "math", "wait STEP_DONE", "st.shared+arrive"});
        row("E3", 28, {"wait STEP_DONE", "ld+zero+arrive", "residual+store"});
    }
#endif
#ifdef B200SR_RS_PROF
    static unsigned long long ev[32][4096]; int evn[32];
    cudaMemcpyFromSymbol(ev, g_rs_evt, sizeof ev); cudaMemcpyFromSymbol(evn, g_rs_evtn, sizeof evn);
    struct E { unsigned long long t; int w, id, k; };
    std::vector<E> all;
    const int ws[] = {0, 1, 2, 3, 4, 8, 12, 16, 20, 24, 28};
    for (int w : ws) { int cnt[1024] = {0}; for (int i = 0; i < evn[w]; ++i) { int id = (int)(ev[w][i] >> 48); all.push_back({ev[w][i] & 0xFFFFFFFFFFFFull, w, id, cnt[id]++}); } }
    std::sort(all.begin(), all.end(), [](const E &a, const E &b) { return a.t < b.t; });
    {   // issue durations: A even (warp 0), A odd (warp 1): 100 -> 102;  B (warp 2): 200 -> 201
        for (int w : {0, 1, 2}) {
            const int a = w < 2 ? 100 : 200, b = w < 2 ? 102 : 201;
            std::vector<unsigned long long> ta_, tb_;
            for (auto &e : all) if (e.w == w) { if (e.id == a) ta_.push_back(e.t); if (e.id == b) tb_.push_back(e.t); }
            const size_t n = std::min(ta_.size(), tb_.size());
            if (n < 8) continue;
            double dur = 0, per = 0; size_t cnt = 0;
            for (size_t i = 4; i + 2 < n; ++i) { dur += (double)(tb_[i] - ta_[i]); per += (double)(ta_[i + 1] - ta_[i]); ++cnt; }
            printf("warp %d: %zu iterations, mean issue %.0f clk, mean period %.0f clk (wait = %.0f)\n", w, n, dur / cnt, per / cnt, (per - dur) / cnt);
        }
    }
    // per-step period as seen by issuer A
    std::vector<unsigned long long> ta;
    for (auto &e : all) if (e.w == 1 && e.id == 100) ta.push_back(e.t);
    printf("issuer A wake times (delta clk):");
    for (size_t i = 1; i < ta.size(); ++i) printf(" %llu", ta[i] - ta[i - 1]);
    printf("\n");
    for (int id : {600, 601, 602}) { printf("producer evt %d since issuer A step 0:", id); for (auto &e : all) if (e.w == 0 && e.id == id && !ta.empty()) printf(" %lld", (long long)e.t - (long long)ta[0]); printf("\n"); }
    if ((int)ta.size() > first + nshow) {
        const unsigned long long t0 = ta[first], t1 = ta[first + nshow];
        printf("timeline from issuer-A step %d (%d steps, %llu clk): warp 0 producer, 1 issuer A, 2 issuer B, 4/8/12/16 E1, 20/24 E2, 28 E3; (k) = occurrence = step for most ids\n", first, nshow, t1 - t0);
        for (auto &e : all) if (e.t >= t0 && e.t < t1) printf("  %6llu  warp %2d  evt %d (%d)\n", e.t - t0, e.w, e.id, e.k);
    }
#endif
    return 0;
}
