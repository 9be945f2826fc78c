// b200sr.cu -- C ABI (include/b200sr.h): plan objects, host-side weight packing, forward orchestration.
//
// No CPU fallback lives here: every compute entry point launches the sm_100a kernels or fails.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/b200sr.h"
#include "common.cuh"
#include "conv.cuh"
#include "launch.h"
#include "wdsr_bf16.cuh"
#include "wdsr_f32.cuh"
#include "wdsr_tc5_layout.cuh"
#include "wdsr_rs_pack.h"
#include "wdsr_tc5_head.cuh"
#include "wdsr_tc5_tail.cuh"

#include <cstdlib>

using namespace b200sr;

namespace {

thread_local std::string g_err;

int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
int cuda_fail(cudaError_t e, const char *what) {
    return fail((int)e, "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
}
#define CU(call)                                        \
    do {                                                \
        cudaError_t e__ = (call);                       \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
    } while (0)

uint16_t f2bf(float f) {  // round-to-nearest-even, same as __float2bfloat16_rn for finite values
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}

size_t esize(int dtype) { return dtype == B200SR_F32 ? 4 : dtype == B200SR_U8 ? 1 : 2; }

// B200SR_CONV_IMPL = mma keeps the video path's 3x3 64 -> 64 convolutions on the mma.sync kernel (developer A/B switch, read per call)
bool conv_tc5_enabled() {
    const char *e = getenv("B200SR_CONV_IMPL");
    return !e || strcmp(e, "mma") != 0;
}

struct BlockW {
    std::vector<float> w1, b1, w2, b2, w3, b3;
    bool set = false;
};

}  // namespace

struct b200sr_wdsr {
    int scale = 0, nb = 0, cin = 0, cp = 0, add_mean = 1;
    float mean = 0.5f;
    int no = 0;  // 3*s*s
    std::vector<int> m1, m2, m1p, m2p_f32, m2p_bf16;
    std::vector<float> head_w, head_b, tail_w, tail_b, skip_w, skip_b;
    bool head_set = false, tail_set = false, committed = false;
    std::vector<BlockW> blocks;
    int device = -1;
    // device images
    float *d_head = nullptr;
    std::vector<float *> d_blk_f32;
    std::vector<uint8_t *> d_blk_bf16;
    std::vector<uint8_t *> d_blk_tc5;  // tcgen05 operand images (nullptr where the block is not eligible); w3 in the form launch_block_tc5 variant 1 expects
    std::vector<uint8_t *> d_blk_tc5u; // the same with the plain 27-slice w3 (sequential / decoupled / chained forms): == d_blk_tc5[i] unless that one is packed
    std::vector<uint8_t *> d_blk_rs;   // row-streaming tcgen05 operand images (wdsr_rs.cuh), same eligibility
    int block_impl = 0;                // 0 = mma.sync kernel, 1 = tcgen05 sequential form, 2 = tcgen05 tile form, 3 = tcgen05 row-streaming form, 4 = row-streaming with the reduce on mma.sync (wdsr_rh.cuh)
    float *d_tail_f32 = nullptr;
    uint8_t *d_tail_bf16 = nullptr;
    uint8_t *d_head_tc5 = nullptr;   // tcgen05 head image (trunk padded to 24 channels)
    uint8_t *d_tail_tc5 = nullptr;   // tcgen05 tail image (nullptr unless the trunk is padded to 24 channels)
    int tail_impl = 1;               // 0 = mma.sync kernel, 1 = tcgen05 kernel
    int head_impl = 0;               // head of the tcgen05 path: 0 = mma.sync kernel (wdsr_head_mma.cu), 1 = tcgen05 kernel (B200SR_HEAD_IMPL=tc5)
    // The bf16 forward is one of two uniform paths: head/block/tail all tcgen05 on a planar-8 trunk [N][3][H][W][8] (tma_map.h), or
    // all mma.sync (+ the sequential tcgen05 reference block) on an NHWC trunk.  Never a mixture: the kernels disagree on the layout.
    bool tc5_path() const {
        if (cp != 24 || block_impl < 2 || !tail_impl || !d_head_tc5 || !d_tail_tc5) return false;
        for (auto b : d_blk_tc5)
            if (!b) return false;
        return true;
    }
    mutable int launches = 0;

    void free_device() {
        if (d_head) cudaFree(d_head), d_head = nullptr;
        for (auto p : d_blk_f32) cudaFree(p);
        for (auto p : d_blk_bf16) cudaFree(p);
        for (size_t i = 0; i < d_blk_tc5u.size(); ++i)
            if (d_blk_tc5u[i] && d_blk_tc5u[i] != d_blk_tc5[i]) cudaFree(d_blk_tc5u[i]);
        for (auto p : d_blk_tc5)
            if (p) cudaFree(p);
        for (auto p : d_blk_rs)
            if (p) cudaFree(p);
        d_blk_f32.clear(), d_blk_bf16.clear(), d_blk_tc5.clear(), d_blk_tc5u.clear(), d_blk_rs.clear();
        if (d_tail_f32) cudaFree(d_tail_f32), d_tail_f32 = nullptr;
        if (d_tail_bf16) cudaFree(d_tail_bf16), d_tail_bf16 = nullptr;
        if (d_tail_tc5) cudaFree(d_tail_tc5), d_tail_tc5 = nullptr;
        if (d_head_tc5) cudaFree(d_head_tc5), d_head_tc5 = nullptr;
    }
};

extern "C" {

int b200sr_version(void) { return B200SR_VERSION; }
const char *b200sr_last_error(void) { return g_err.c_str(); }
int b200sr_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int b200sr_wdsr_create(const b200sr_wdsr_desc *d, b200sr_wdsr_t **out) {
    if (!d || !out) return fail(B200SR_E_INVAL, "wdsr_create: null argument");
    if (d->scale < 2 || d->scale > 4) return fail(B200SR_E_UNSUPPORTED, "wdsr_create: scale %d not in {2,3,4}", d->scale);
    if (d->c_trunk < 1 || d->c_trunk > 24)
        return fail(B200SR_E_UNSUPPORTED, "wdsr_create: trunk width %d not in [1,24]", d->c_trunk);
    if (d->num_blocks < 0 || (d->num_blocks > 0 && (!d->m1 || !d->m2))) return fail(B200SR_E_INVAL, "wdsr_create: bad block list");
    b200sr_wdsr *p = new (std::nothrow) b200sr_wdsr();
    if (!p) return fail(B200SR_E_INVAL, "wdsr_create: out of memory");
    p->scale = d->scale, p->nb = d->num_blocks, p->cin = d->c_trunk, p->cp = round_up(d->c_trunk, 8);
    p->add_mean = d->add_mean, p->mean = d->image_mean, p->no = 3 * d->scale * d->scale;
    for (int i = 0; i < p->nb; ++i) {
        const int m1 = d->m1[i], m2 = d->m2[i];
        if (m1 < 1 || m1 > 160 || m2 < 1 || m2 > 24) {
            delete p;
            return fail(B200SR_E_UNSUPPORTED, "wdsr_create: block %d widths (M1=%d, M2=%d) outside [1,160]x[1,24]", i, m1, m2);
        }
        p->m1.push_back(m1), p->m2.push_back(m2);
        p->m1p.push_back(round_up(m1, 16));
        p->m2p_f32.push_back(round_up(m2, 4) < 8 ? 8 : round_up(m2, 4));
        p->m2p_bf16.push_back(round_up(m2, 8));
    }
    {
        // Pruned trunks (IN < 24) are padded to 24 channels whenever every block then fits the tcgen05 kernels (M1 <= 144):
        // the padded channels are exact zeros and cost HBM bytes, but the tcgen05 head / block / tail beat the mma.sync
        // specialisations by 1.2-1.5x even so (cfg3 P1 360p x 8: 1176 -> 791 us, P2: 164 -> 115 us per frame on B200).
        // B200SR_TRUNK_PAD24=0 keeps the narrow trunk (developer switch, A/B timing).
        bool fits = true;
        for (int m1 : p->m1) fits = fits && round_up(m1, 16) <= 144;
        const char *e = getenv("B200SR_TRUNK_PAD24");
        if (fits && !(e && e[0] == '0')) p->cp = 24;
    }
    p->blocks.resize(p->nb);
    *out = p;
    return 0;
}

void b200sr_wdsr_destroy(b200sr_wdsr_t *p) {
    if (!p) return;
    p->free_device();
    delete p;
}

int b200sr_wdsr_set_head(b200sr_wdsr_t *p, const float *w, const float *b) {
    if (!p || !w || !b) return fail(B200SR_E_INVAL, "set_head: null argument");
    p->head_w.assign(w, w + (size_t)p->cin * 27);
    p->head_b.assign(b, b + p->cin);
    p->head_set = true, p->committed = false;
    return 0;
}

int b200sr_wdsr_set_block(b200sr_wdsr_t *p, int i, const float *w1, const float *b1, const float *w2, const float *b2,
                          const float *w3, const float *b3) {
    if (!p || !w1 || !b1 || !w2 || !b2 || !w3 || !b3) return fail(B200SR_E_INVAL, "set_block: null argument");
    if (i < 0 || i >= p->nb) return fail(B200SR_E_INVAL, "set_block: block %d out of range [0,%d)", i, p->nb);
    BlockW &k = p->blocks[i];
    const int m1 = p->m1[i], m2 = p->m2[i], c = p->cin;
    k.w1.assign(w1, w1 + (size_t)m1 * c), k.b1.assign(b1, b1 + m1);
    k.w2.assign(w2, w2 + (size_t)m2 * m1), k.b2.assign(b2, b2 + m2);
    k.w3.assign(w3, w3 + (size_t)c * m2 * 9), k.b3.assign(b3, b3 + c);
    k.set = true, p->committed = false;
    return 0;
}

int b200sr_wdsr_set_tail(b200sr_wdsr_t *p, const float *wt, const float *bt, const float *ws, const float *bs) {
    if (!p || !wt || !bt || !ws || !bs) return fail(B200SR_E_INVAL, "set_tail: null argument");
    p->tail_w.assign(wt, wt + (size_t)p->no * p->cin * 9), p->tail_b.assign(bt, bt + p->no);
    p->skip_w.assign(ws, ws + (size_t)p->no * 75), p->skip_b.assign(bs, bs + p->no);
    p->tail_set = true, p->committed = false;
    return 0;
}

static int upload(const void *host, size_t bytes, void **dev) {
    CU(cudaMalloc(dev, bytes));
    CU(cudaMemcpy(*dev, host, bytes, cudaMemcpyHostToDevice));
    return 0;
}

int b200sr_wdsr_commit(b200sr_wdsr_t *p) {
    if (!p) return fail(B200SR_E_INVAL, "commit: null plan");
    if (!p->head_set || !p->tail_set) return fail(B200SR_E_STATE, "commit: head/tail weights not set");
    for (int i = 0; i < p->nb; ++i)
        if (!p->blocks[i].set) return fail(B200SR_E_STATE, "commit: block %d weights not set", i);
    if (b200sr_device_count() <= 0) return fail(B200SR_E_STATE, "commit: no CUDA device (this library has no CPU fallback)");
    CU(cudaGetDevice(&p->device));
    p->free_device();
    const int C = p->cin, CP = p->cp;
    int rc;
    {   // head: [27][CP] (k = c*9+ky*3+kx) + bias[CP]
        std::vector<float> h((size_t)27 * CP + CP, 0.f);
        for (int o = 0; o < C; ++o) {
            for (int k = 0; k < 27; ++k) h[(size_t)k * CP + o] = p->head_w[(size_t)o * 27 + k];
            h[(size_t)27 * CP + o] = p->head_b[o];
        }
        if ((rc = upload(h.data(), h.size() * 4, (void **)&p->d_head))) return rc;
    }
    for (int i = 0; i < p->nb; ++i) {
        const BlockW &k = p->blocks[i];
        const int M1 = p->m1[i], M2 = p->m2[i], M1P = p->m1p[i];
        {   // fp32 image
            const int M2P = p->m2p_f32[i];
            BlockF32Layout L(CP, M1P, M2P);
            std::vector<float> f((size_t)L.total, 0.f);
            for (int m = 0; m < M1; ++m) {
                for (int c = 0; c < C; ++c) f[L.w1 + (size_t)m * CP + c] = k.w1[(size_t)m * C + c];
                f[L.b1 + m] = k.b1[m];
                for (int j = 0; j < M2; ++j) f[L.w2 + (size_t)m * M2P + j] = k.w2[(size_t)j * M1 + m];
            }
            for (int j = 0; j < M2; ++j) f[L.b2 + j] = k.b2[j];
            for (int o = 0; o < C; ++o) {
                for (int j = 0; j < M2; ++j)
                    for (int tap = 0; tap < 9; ++tap)
                        f[L.w3 + ((size_t)tap * M2P + j) * CP + o] = k.w3[((size_t)o * M2 + j) * 9 + tap];
                f[L.b3 + o] = k.b3[o];
            }
            float *d = nullptr;
            if ((rc = upload(f.data(), f.size() * 4, (void **)&d))) return rc;
            p->d_blk_f32.push_back(d);
        }
        {   // bf16 image (verbatim shared-memory layout)
            const int M2P = p->m2p_bf16[i];
            BlockBf16Layout L(CP, M1P, M2P);
            std::vector<uint8_t> img((size_t)L.total, 0);
            uint16_t *w1 = (uint16_t *)(img.data() + L.w1), *w2 = (uint16_t *)(img.data() + L.w2),
                     *w3 = (uint16_t *)(img.data() + L.w3);
            float *b1 = (float *)(img.data() + L.b1), *b2 = (float *)(img.data() + L.b2), *b3 = (float *)(img.data() + L.b3);
            for (int m = 0; m < M1; ++m) {
                for (int c = 0; c < C; ++c) w1[(size_t)m * L.s1 + c] = f2bf(k.w1[(size_t)m * C + c]);
                b1[m] = k.b1[m];
            }
            for (int j = 0; j < M2; ++j) {
                for (int m = 0; m < M1; ++m) w2[(size_t)j * L.s2 + m] = f2bf(k.w2[(size_t)j * M1 + m]);
                b2[j] = k.b2[j];
            }
            for (int o = 0; o < C; ++o) {
                for (int j = 0; j < M2; ++j)
                    for (int tap = 0; tap < 9; ++tap)
                        w3[((size_t)tap * CP + o) * L.s3 + j] = f2bf(k.w3[((size_t)o * M2 + j) * 9 + tap]);
                b3[o] = k.b3[o];
            }
            uint8_t *d = nullptr;
            if ((rc = upload(img.data(), img.size(), (void **)&d))) return rc;
            p->d_blk_bf16.push_back(d);
        }
        if (CP == 24 && M2 <= 24 && M1P <= 144) {   // tcgen05 operand image (TMEM holds two 144-column expand accumulators) (interleaved K-major core matrices, see wdsr_tc5.cuh)
            BlockTc5Layout L(M1P);
            std::vector<uint8_t> img((size_t)L.total, 0);
            auto at = [&](int off) { return (uint16_t *)(img.data() + off); };
            for (int n = 0; n < M1; ++n) {
                for (int c = 0; c < C; ++c) at(L.w1 + (n / 8) * 512 + (c / 8) * 128 + (n % 8) * 16)[c % 8] = f2bf(k.w1[(size_t)n * C + c]);
                const uint16_t hi = f2bf(k.b1[n]);
                uint32_t hu = (uint32_t)hi << 16;
                float hf;
                memcpy(&hf, &hu, 4);
                uint16_t *bc = at(L.w1 + (n / 8) * 512 + 3 * 128 + (n % 8) * 16);
                bc[0] = hi, bc[1] = f2bf(k.b1[n] - hf);          // b1 = hi + lo against the two constant-one channels
            }
            for (int j = 0; j < M2; ++j)
                for (int m = 0; m < M1; ++m)
                    at(L.w2 + (j / 8) * L.sbo2 + (m / 8) * 128 + (j % 8) * 16)[m % 8] = f2bf(k.w2[(size_t)j * M1 + m]);
            const int NC2 = (M2 + 7) / 8;   // 8-channel chunks of t2 (3 dense; pruned blocks skip the 3x3 MMAs of absent chunks)
            auto pack_w3 = [&](bool packed) {   // (re)writes the w3 region
                memset(img.data() + L.w3, 0, (size_t)4 * 28 * 128);
                for (int o = 0; o < C; ++o)
                    for (int j = 0; j < M2; ++j)
                        for (int dy = 0; dy < 3; ++dy)
                            for (int dx = 0; dx < 3; ++dx) {
                                int q, kk;
                                if (packed && j >= 16) {
                                    // last chunk (channels 16..19) packed two horizontal taps per K = 8 half (wdsr_tc5p.cuh, NC2 = 4): slice 18 + 2 dy
                                    // = [tap dx 0 | tap dx 1], slice 19 + 2 dy = [tap dx 2 | zero]
                                    q = 18 + 2 * dy + (dx == 2 ? 1 : 0);
                                    kk = (j - 16) + (dx == 1 ? 4 : 0);
                                } else if (packed) {
                                    q = (dx * 3 + dy) * 2 + j / 8, kk = j % 8;
                                } else {
                                    q = (dx * 3 + dy) * NC2 + j / 8, kk = j % 8;   // only the (tap, chunk) slices t2 really has
                                }
                                at(L.w3 + (o / 8) * (28 * 128) + q * 128 + (o % 8) * 16)[kk] = f2bf(k.w3[((size_t)o * M2 + j) * 9 + dy * 3 + dx]);
                            }
            };
            pack_w3(false);
            float *b2 = (float *)(img.data() + L.b2), *b3 = (float *)(img.data() + L.b3);
            for (int j = 0; j < M2; ++j) b2[j] = k.b2[j];
            b2[31] = (float)NC2;   // informational (the launcher derives the same count from M2 and picks the kernel instantiation)
            for (int o = 0; o < C; ++o) b3[o] = k.b3[o];
            uint8_t *d = nullptr;
            if ((rc = upload(img.data(), img.size(), (void **)&d))) return rc;
            p->d_blk_tc5u.push_back(d);
            if (block_tc5_g3_packed(M2)) {   // the tile form's packed 3x3 (12 instead of 14 MMAs per M-tile) reads its own image
                pack_w3(true);
                d = nullptr;
                if ((rc = upload(img.data(), img.size(), (void **)&d))) return rc;
            }
            p->d_blk_tc5.push_back(d);
            std::vector<uint8_t> rimg;   // row-streaming form: same w1 / w2 images, the 3x3 with dy stacked in N (wdsr_rs_pack.h)
            pack_block_rs(rimg, C, M1, M2, M1P, k.w1.data(), k.b1.data(), k.w2.data(), k.b2.data(), k.w3.data(), k.b3.data());
            d = nullptr;
            if ((rc = upload(rimg.data(), rimg.size(), (void **)&d))) return rc;
            p->d_blk_rs.push_back(d);
        } else {
            p->d_blk_tc5.push_back(nullptr);
            p->d_blk_tc5u.push_back(nullptr);
            p->d_blk_rs.push_back(nullptr);
        }
    }
    {
        // default: the tcgen05 kernel wherever a block is eligible (trunk padded to 24, M2 <= 24), the mma.sync kernel
        // elsewhere.  B200SR_BLOCK_IMPL = mma | tc5seq | tc5 | tc5q | rs | rh is a developer switch (A/B timing, reference form):
        // tc5 = the tile form (wdsr_tc5p.cuh, the default); rs = the row-streaming form (wdsr_rs.cuh; cfg2 28.4 vs 26.9 us, 1080p 74-82 vs
        // 80 us per launch); rh = row-streaming with the reduce 1x1 on mma.sync out of registers (wdsr_rh.cuh; 32.5 / 94 us: bound by the
        // legacy HMMA rate, profiles/r02_block_rh.md); tc5q = the tile form with decoupled expand staging (wdsr_tc5q.cuh; 27.7 / 86 us).
        // All of them are parity-green against the oracle and each other.
        const char *e = getenv("B200SR_BLOCK_IMPL");
        p->block_impl = !e ? 2 : !strcmp(e, "mma") ? 0 : !strcmp(e, "tc5seq") ? 1 : !strcmp(e, "rs") ? 3 : !strcmp(e, "rh") ? 4 : !strcmp(e, "tc5q") ? 5 : !strcmp(e, "chain") ? 6 : 2;
    }
    const int NO = p->no;
    {   // tail fp32: Wt[9][CP][NOP4] | Ws[75][NOP4] | bias[NOP4]
        const int NOP = round_up(NO, 4);
        std::vector<float> f((size_t)9 * CP * NOP + 75 * NOP + NOP, 0.f);
        float *wt = f.data(), *ws = wt + (size_t)9 * CP * NOP, *bias = ws + (size_t)75 * NOP;
        for (int o = 0; o < NO; ++o) {
            for (int c = 0; c < C; ++c)
                for (int tap = 0; tap < 9; ++tap) wt[((size_t)tap * CP + c) * NOP + o] = p->tail_w[((size_t)o * C + c) * 9 + tap];
            for (int k = 0; k < 75; ++k) ws[(size_t)k * NOP + o] = p->skip_w[(size_t)o * 75 + k];
            bias[o] = p->tail_b[o] + p->skip_b[o];
        }
        if ((rc = upload(f.data(), f.size() * 4, (void **)&p->d_tail_f32))) return rc;
    }
    {   // tail bf16: Wt[9][NOP8][st] | Ws[NOP8][120] (k = ky*24 + dx*4 + c) | bias[NOP8]
        const int NOP = round_up(NO, 8);
        TailBf16Layout L(CP, NOP);
        std::vector<uint8_t> img((size_t)L.total, 0);
        uint16_t *wt = (uint16_t *)(img.data() + L.wt), *ws = (uint16_t *)(img.data() + L.ws);
        float *bias = (float *)(img.data() + L.bias);
        for (int o = 0; o < NO; ++o) {
            for (int c = 0; c < C; ++c)
                for (int tap = 0; tap < 9; ++tap)
                    wt[((size_t)tap * NOP + o) * L.st + c] = f2bf(p->tail_w[((size_t)o * C + c) * 9 + tap]);
            for (int c = 0; c < 3; ++c)
                for (int ky = 0; ky < 5; ++ky)
                    for (int kx = 0; kx < 5; ++kx)
                        ws[(size_t)o * 120 + ky * 24 + kx * 4 + c] = f2bf(p->skip_w[(size_t)o * 75 + c * 25 + ky * 5 + kx]);
            bias[o] = p->tail_b[o] + p->skip_b[o];
        }
        if ((rc = upload(img.data(), img.size(), (void **)&p->d_tail_bf16))) return rc;
    }
    if (CP == 24) {   // tcgen05 head image (wdsr_tc5_head.cuh): k = window pixel (ky*3+kx) * 4 + channel
        HeadTc5Layout L;
        std::vector<uint8_t> img((size_t)L.total, 0);
        for (int o = 0; o < C; ++o) {
            for (int c = 0; c < 3; ++c)
                for (int t = 0; t < 9; ++t) {
                    const int k = t * 4 + c;
                    ((uint16_t *)(img.data() + L.w + (o / 8) * (6 * 128) + (k / 8) * 128 + (o % 8) * 16))[k % 8] = f2bf(p->head_w[(size_t)o * 27 + c * 9 + t]);
                }
            ((float *)(img.data() + L.bias))[o] = p->head_b[o];
        }
        if ((rc = upload(img.data(), img.size(), (void **)&p->d_head_tc5))) return rc;
    }
    if (CP == 24) {   // tcgen05 tail image (wdsr_tc5_tail.cuh): 3x3 chunks ordered (dx, c, dy); skip = 15 (window row, tap pair) chunks x [2 taps x 4 channels]
        const int NOP = round_up(NO, 16);
        TailTc5Layout L(NOP);
        std::vector<uint8_t> img((size_t)L.total, 0);
        auto at = [&](int off) { return (uint16_t *)(img.data() + off); };
        for (int o = 0; o < NO; ++o) {
            for (int c = 0; c < C; ++c)
                for (int dy = 0; dy < 3; ++dy)
                    for (int dx = 0; dx < 3; ++dx) {
                        const int q = (dx * 3 + c / 8) * 3 + dy;
                        at(L.wt + (o / 8) * L.sbo_t + q * 128 + (o % 8) * 16)[c % 8] = f2bf(p->tail_w[((size_t)o * C + c) * 9 + dy * 3 + dx]);
                    }
            for (int c = 0; c < 3; ++c)
                for (int ky = 0; ky < 5; ++ky)
                    for (int kx = 0; kx < 5; ++kx) {
                        // chunk (ky, j = kx / 2) holds taps 2j | 2j+1 x 4 channels; instruction order (wdsr_tc5_tail.cuh): i = ky: (ky,0) | (ky,1);
                        // then the j = 2 chunks in pairs: (0,2) | (1,2), (2,2) | (3,2), (4,2) | zero
                        const int j = kx / 2;
                        const int chunk = j < 2 ? 2 * ky + j : 10 + ky;
                        const int k = (kx & 1) * 4 + c;
                        at(L.ws + (o / 8) * L.sbo_s + chunk * 128 + (o % 8) * 16)[k] = f2bf(p->skip_w[(size_t)o * 75 + c * 25 + ky * 5 + kx]);
                    }
            ((float *)(img.data() + L.bias))[o] = p->tail_b[o] + p->skip_b[o];
        }
        if ((rc = upload(img.data(), img.size(), (void **)&p->d_tail_tc5))) return rc;
    }
    {
        const char *e = getenv("B200SR_TAIL_IMPL");   // developer switch: mma | tc5
        p->tail_impl = (e && !strcmp(e, "mma")) ? 0 : 1;
        const char *eh = getenv("B200SR_HEAD_IMPL");   // developer switch: mma (default) | tc5
        p->head_impl = (eh && !strcmp(eh, "tc5")) ? 1 : 0;
    }
    p->committed = true;
    return 0;
}

int b200sr_wdsr_trunk_channels(const b200sr_wdsr_t *p) { return p ? p->cp : 0; }
int b200sr_wdsr_trunk_layout(const b200sr_wdsr_t *p, int precision) {
    return p && precision == B200SR_BF16 && p->tc5_path() ? B200SR_TRUNK_PLANAR8 : B200SR_TRUNK_NHWC;
}
int b200sr_wdsr_launches_per_forward(const b200sr_wdsr_t *p) { return p ? p->launches : 0; }

int b200sr_wdsr_pack_block_image(int c, int m1, int m2, const float *w1, const float *b1, const float *w2, const float *b2, const float *w3,
                                 const float *b3, void *img_host, size_t cap, size_t *bytes) {
    if (!w1 || !b1 || !w2 || !b2 || !w3 || !b3 || !bytes) return fail(B200SR_E_INVAL, "pack_block_image: null argument");
    if (c < 1 || c > 24 || m1 < 1 || m1 > 144 || m2 < 1 || m2 > 24)
        return fail(B200SR_E_UNSUPPORTED, "pack_block_image: (c=%d, m1=%d, m2=%d) outside [1,24]x[1,144]x[1,24]", c, m1, m2);
    std::vector<uint8_t> img;
    pack_block_rs(img, c, m1, m2, round_up(m1, 16), w1, b1, w2, b2, w3, b3);
    *bytes = img.size();
    if (img_host) {
        if (cap < img.size()) return fail(B200SR_E_WORKSPACE, "pack_block_image: buffer %zu < %zu bytes", cap, img.size());
        memcpy(img_host, img.data(), img.size());
    }
    return 0;
}

size_t b200sr_wdsr_workspace_bytes(const b200sr_wdsr_t *p, int n, int h, int w, int precision) {
    if (!p || n <= 0 || h <= 0 || w <= 0) return 0;
    const size_t trunk = round_up((int)(((size_t)n * h * w * p->cp * esize(precision) + 255) / 256), 1) * (size_t)256;
    return 2 * trunk + 256;   // two ping-pong trunk buffers + the layer counter of the chained block launch (wdsr_tc5c.cuh)
}

static int check_common(const b200sr_wdsr_t *p, int n, int h, int w, int precision, const char *who) {
    if (!p) return fail(B200SR_E_INVAL, "%s: null plan", who);
    if (!p->committed) return fail(B200SR_E_STATE, "%s: weights not committed", who);
    if (n <= 0 || h <= 0 || w <= 0) return fail(B200SR_E_INVAL, "%s: bad shape n=%d h=%d w=%d", who, n, h, w);
    if (precision != B200SR_F32 && precision != B200SR_BF16) return fail(B200SR_E_INVAL, "%s: bad precision %d", who, precision);
    if ((long long)n * h * w * p->scale * p->scale * 3 >= (1ll << 40)) return fail(B200SR_E_INVAL, "%s: tensor too large", who);
    return 0;
}

int b200sr_wdsr_head(const b200sr_wdsr_t *p, const void *x, int x_dtype, void *trunk, int n, int h, int w, int precision,
                     void *stream) {
    int rc = check_common(p, n, h, w, precision, "wdsr_head");
    if (rc) return rc;
    if (!x || !trunk) return fail(B200SR_E_INVAL, "wdsr_head: null tensor");
    if (precision == B200SR_BF16 && p->tc5_path()) {
        if (p->head_impl == 1)
            CU(launch_head_tc5(x_dtype, x, trunk, p->d_head_tc5, n, h, w, p->mean, (cudaStream_t)stream));
        else
            CU(launch_head_mma(x_dtype, x, trunk, p->d_head, n, h, w, p->mean, (cudaStream_t)stream));
    } else
        CU(launch_head(p->cp, x_dtype, precision, x, trunk, p->d_head, n, h, w, p->mean, (cudaStream_t)stream));
    return 0;
}

int b200sr_wdsr_block(const b200sr_wdsr_t *p, int i, const void *tin, void *tout, int n, int h, int w, int precision,
                      void *stream) {
    int rc = check_common(p, n, h, w, precision, "wdsr_block");
    if (rc) return rc;
    if (i < 0 || i >= p->nb) return fail(B200SR_E_INVAL, "wdsr_block: block %d out of range", i);
    if (!tin || !tout || tin == tout) return fail(B200SR_E_INVAL, "wdsr_block: in/out must be distinct non-null buffers");
    if (precision == B200SR_F32)
        CU(launch_block_f32(p->cp, p->m2p_f32[i], (const float *)tin, (float *)tout, p->d_blk_f32[i], p->m1p[i], n, h, w,
                            (cudaStream_t)stream));
    else if (p->tc5_path() && p->block_impl == 3 && block_rs_eligible(n, h, w))
        CU(launch_block_rs(tin, tout, p->d_blk_rs[i], p->m1p[i], p->m2[i], n, h, w, (cudaStream_t)stream));
    else if (p->tc5_path() && p->block_impl == 4 && block_rs_eligible(n, h, w))
        CU(launch_block_rh(tin, tout, p->d_blk_rs[i], p->m1p[i], p->m2[i], n, h, w, (cudaStream_t)stream));
    else if (p->tc5_path())
        CU(launch_block_tc5(p->block_impl == 5 ? 2 : 1, tin, tout, p->block_impl == 5 ? p->d_blk_tc5u[i] : p->d_blk_tc5[i], p->m1p[i], p->m2[i], n, h, w,
                            (cudaStream_t)stream));
    else if (p->block_impl == 1 && p->d_blk_tc5[i] && p->m2[i] > 16)   // sequential tcgen05 reference form (NHWC trunk, all 27 w3 slices; developer switch)
        CU(launch_block_tc5(0, tin, tout, p->d_blk_tc5u[i], p->m1p[i], p->m2[i], n, h, w, (cudaStream_t)stream));
    else
        CU(launch_block_bf16(p->cp, p->m2p_bf16[i], tin, tout, p->d_blk_bf16[i], p->m1p[i], n, h, w, (cudaStream_t)stream));
    return 0;
}

int b200sr_wdsr_tail(const b200sr_wdsr_t *p, const void *trunk, const void *x, int x_dtype, void *y, int y_dtype, int n,
                     int h, int w, int precision, void *stream) {
    int rc = check_common(p, n, h, w, precision, "wdsr_tail");
    if (rc) return rc;
    if (!trunk || !x || !y) return fail(B200SR_E_INVAL, "wdsr_tail: null tensor");
    if (y_dtype == B200SR_U8 && (precision != B200SR_BF16 || !p->tc5_path()))
        return fail(B200SR_E_UNSUPPORTED, "wdsr_tail: the 8-bit frame output is written by the tcgen05 tail only (bf16 precision, trunk of 24 channels)");
    const float out_add = p->add_mean ? p->mean : 0.f;
    if (precision == B200SR_F32)
        CU(launch_tail_f32(p->cp, p->scale, x_dtype, y_dtype, (const float *)trunk, x, y, p->d_tail_f32, n, h, w, p->mean,
                           out_add, (cudaStream_t)stream));
    else if (p->tc5_path())
        CU(launch_tail_tc5(p->scale, x_dtype, y_dtype, trunk, x, y, p->d_tail_tc5, n, h, w, p->mean, out_add, (cudaStream_t)stream));
    else
        CU(launch_tail_bf16(p->cp, p->scale, x_dtype, y_dtype, trunk, x, y, p->d_tail_bf16, n, h, w, p->mean, out_add,
                            (cudaStream_t)stream));
    return 0;
}

int b200sr_wdsr_forward(const b200sr_wdsr_t *p, const void *x, int x_dtype, void *y, int y_dtype, int n, int h, int w,
                        int precision, void *ws, size_t ws_bytes, void *stream) {
    int rc = check_common(p, n, h, w, precision, "wdsr_forward");
    if (rc) return rc;
    if (!x || !y || !ws) return fail(B200SR_E_INVAL, "wdsr_forward: null tensor/workspace");
    const size_t need = b200sr_wdsr_workspace_bytes(p, n, h, w, precision);
    if (ws_bytes < need) return fail(B200SR_E_WORKSPACE, "wdsr_forward: workspace %zu < %zu bytes", ws_bytes, need);
    const size_t trunk = (need - 256) / 2;
    uint8_t *a = (uint8_t *)ws, *b = a + trunk;
    int launches = 0;
    if ((rc = b200sr_wdsr_head(p, x, x_dtype, a, n, h, w, precision, stream))) return rc;
    ++launches;
    bool chained = false;
    if (precision == B200SR_BF16 && p->block_impl == 6 && p->tc5_path() && p->nb >= 1 && p->nb <= 32) {
        // every block in ONE persistent cooperative launch (wdsr_tc5c.cuh) when they all have the same shape
        bool uniform = true;
        for (int i = 0; i < p->nb; ++i) uniform = uniform && p->d_blk_tc5[i] && p->m1p[i] == p->m1p[0] && p->m2[i] == p->m2[0];
        if (uniform) {
            CU(launch_block_chain_tc5(a, b, p->d_blk_tc5u.data(), p->nb, (unsigned *)((uint8_t *)ws + 2 * trunk), p->m1p[0], p->m2[0], n, h, w,
                                      (cudaStream_t)stream));
            launches += 2;   // the counter's memset node + the kernel
            if (p->nb & 1) a = b;
            chained = true;
        }
    }
    for (int i = 0; i < p->nb && !chained; ++i) {
        if ((rc = b200sr_wdsr_block(p, i, a, b, n, h, w, precision, stream))) return rc;
        ++launches;
        uint8_t *t = a;
        a = b, b = t;
    }
    if ((rc = b200sr_wdsr_tail(p, a, x, x_dtype, y, y_dtype, n, h, w, precision, stream))) return rc;
    p->launches = launches + 1;
    return 0;
}

int b200sr_wdsr_forward_host(const b200sr_wdsr_t *p, const void *xh, int x_dtype, void *yh, int y_dtype, int n, int h, int w,
                             int precision, void *xd, void *yd, void *ws, size_t ws_bytes, void *stream) {
    int rc = check_common(p, n, h, w, precision, "wdsr_forward_host");
    if (rc) return rc;
    if (!xh || !yh || !xd || !yd) return fail(B200SR_E_INVAL, "wdsr_forward_host: null buffer");
    const size_t xb = (size_t)n * 3 * h * w * esize(x_dtype);
    const size_t yb = (size_t)n * 3 * h * w * p->scale * p->scale * esize(y_dtype);
    CU(cudaMemcpyAsync(xd, xh, xb, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    if ((rc = b200sr_wdsr_forward(p, xd, x_dtype, yd, y_dtype, n, h, w, precision, ws, ws_bytes, stream))) return rc;
    CU(cudaMemcpyAsync(yh, yd, yb, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return 0;
}

int b200sr_flow_warp_nchw(const float *x, const float *flow, int64_t fs_n, int64_t fs_h, int64_t fs_w, int64_t fs_c, float *y,
                          int n, int c, int h, int w, int padding_mode, void *stream) {
    if (!x || !flow || !y) return fail(B200SR_E_INVAL, "flow_warp: null tensor");
    if (n < 0 || c < 0 || h <= 0 || w <= 0) return fail(B200SR_E_INVAL, "flow_warp: bad shape");
    if (padding_mode != B200SR_PAD_ZEROS && padding_mode != B200SR_PAD_BORDER)
        return fail(B200SR_E_UNSUPPORTED, "flow_warp: padding_mode %d (only zeros/border)", padding_mode);
    CU(launch_flow_warp_nchw(x, flow, fs_n, fs_h, fs_w, fs_c, y, n, c, h, w, padding_mode == B200SR_PAD_BORDER,
                             (cudaStream_t)stream));
    return 0;
}

int b200sr_flow_warp_nhwc(const void *x, const float *flow, void *y, int n, int c, int h, int w, int padding_mode, int dtype,
                          void *stream) {
    if (!x || !flow || !y) return fail(B200SR_E_INVAL, "flow_warp_nhwc: null tensor");
    if (padding_mode != B200SR_PAD_ZEROS && padding_mode != B200SR_PAD_BORDER)
        return fail(B200SR_E_UNSUPPORTED, "flow_warp_nhwc: padding_mode %d", padding_mode);
    cudaError_t e = launch_flow_warp_nhwc(x, flow, y, n, c, h, w, padding_mode == B200SR_PAD_BORDER, dtype, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "flow_warp_nhwc (c must be a multiple of 8 (bf16) / 4 (f32), c*esize/16 in {1,2,3,4,6,8,16})");
    return 0;
}

int b200sr_flow_warp_nhwc_into(const void *x, const float *flow, void *y, int y_cs, int y_co, int n, int c, int h, int w, int padding_mode,
                               int dtype, void *stream) {
    if (!x || !flow || !y) return fail(B200SR_E_INVAL, "flow_warp_nhwc_into: null tensor");
    if (padding_mode != B200SR_PAD_ZEROS && padding_mode != B200SR_PAD_BORDER)
        return fail(B200SR_E_UNSUPPORTED, "flow_warp_nhwc_into: padding_mode %d", padding_mode);
    if (y_cs < c || y_co < 0 || y_co + c > y_cs) return fail(B200SR_E_INVAL, "flow_warp_nhwc_into: channel window outside tensor");
    cudaError_t e = launch_flow_warp_nhwc(x, flow, y, n, c, h, w, padding_mode == B200SR_PAD_BORDER, dtype, (cudaStream_t)stream, y_cs, y_co);
    if (e != cudaSuccess) return cuda_fail(e, "flow_warp_nhwc_into (c, y_cstride, y_coff multiples of 8 (bf16) / 4 (f32); c*esize/16 in {1,2,3,4,6,8,16})");
    return 0;
}

int b200sr_flow_warp_nhwc_windows(const void *x, int x_cs, int x_co, const float *flow, void *y, int y_cs, int y_co, int n, int c, int h, int w,
                                  int padding_mode, int dtype, void *stream) {
    if (!x || !flow || !y) return fail(B200SR_E_INVAL, "flow_warp_nhwc_windows: null tensor");
    if (padding_mode != B200SR_PAD_ZEROS && padding_mode != B200SR_PAD_BORDER)
        return fail(B200SR_E_UNSUPPORTED, "flow_warp_nhwc_windows: padding_mode %d", padding_mode);
    if (y_cs < c || y_co < 0 || y_co + c > y_cs || x_cs < c || x_co < 0 || x_co + c > x_cs)
        return fail(B200SR_E_INVAL, "flow_warp_nhwc_windows: channel window outside tensor");
    cudaError_t e = launch_flow_warp_nhwc(x, flow, y, n, c, h, w, padding_mode == B200SR_PAD_BORDER, dtype, (cudaStream_t)stream, y_cs, y_co, x_cs, x_co);
    if (e != cudaSuccess) return cuda_fail(e, "flow_warp_nhwc_windows (c, strides, offsets multiples of 8 (bf16) / 4 (f32); c*esize/16 in {1,2,3,4,6,8,16})");
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// Split_Block (the fork's searchable block body)
// ---------------------------------------------------------------------------------------------------------
struct b200sr_split {
    int c = 0;
    std::vector<float> params;   // packed image, passed to the kernel by value (constant bank)
    std::vector<uint8_t> tc;     // parameter image of the bf16 tensor-core arm (split_block_tc.cu), derived from params ...
    void *d_tc = nullptr;        // ... and its copy in device memory (the device that was current at b200sr_split_create)
    bool ffma_only = false;      // B200SR_SPLIT_IMPL=ffma at create time: keep the fp32-FFMA kernel for bf16 tensors too (developer switch)
};

static int split_upload_tc(b200sr_split *b) {
    split_tc_pack(b->c, b->params.data(), b->tc);
    if (b->tc.empty()) return 0;
    if (!b->d_tc) CU(cudaMalloc(&b->d_tc, b->tc.size()));
    CU(cudaMemcpy(b->d_tc, b->tc.data(), b->tc.size(), cudaMemcpyHostToDevice));
    return 0;
}

// bf16 tensors with 16-byte rows take the tensor-core arm; B200SR_SPLIT_IMPL=ffma keeps the fp32-FFMA kernel (developer switch)
static cudaError_t run_split(const b200sr_split *b, int dtype, const void *x, void *y, int n, int h, int w, cudaStream_t st) {
    if (!b->ffma_only && b->d_tc && split_tc_eligible(dtype, x, y, w)) return launch_split_block_tc(b->c, x, y, (const uint8_t *)b->d_tc, n, h, w, st);
    return launch_split_block(b->c, dtype, x, y, b->params.data(), n, h, w, st);
}

int b200sr_split_create(int C, const float *dw3, const float *dw5, const float *dw7, const float *dwb, const float *pw, const float *pwb,
                        const float *e, const float *prob, b200sr_split_t **out) {
    if (!dw3 || !dw5 || !dw7 || !dwb || !pw || !pwb || !e || !prob || !out) return fail(B200SR_E_INVAL, "split_create: null argument");
    if (C != 8 && C != 16 && C != 24 && C != 32) return fail(B200SR_E_UNSUPPORTED, "split_create: channels=%d (8, 16, 24 or 32)", C);
    if (b200sr_device_count() <= 0) return fail(B200SR_E_STATE, "split_create: no CUDA device (this library has no CPU fallback)");
    std::vector<float> f((size_t)split_param_floats(C), 0.f);
    float *q = f.data();
    memcpy(q, dw3, sizeof(float) * C * 9), q += C * 9;
    memcpy(q, dw5, sizeof(float) * C * 25), q += C * 25;
    memcpy(q, dw7, sizeof(float) * C * 49), q += C * 49;
    memcpy(q, dwb, sizeof(float) * 3 * C), q += 3 * C;
    for (int k = 0; k < 3; ++k)            // [k][out][in] -> [k][in][out]: a thread reads one input channel's row of outputs
        for (int o = 0; o < C; ++o)
            for (int i = 0; i < C; ++i) q[((size_t)k * C + i) * C + o] = pw[((size_t)k * C + o) * C + i];
    q += 3 * C * C;
    memcpy(q, pwb, sizeof(float) * 3 * C), q += 3 * C;
    memcpy(q, e, sizeof(float) * C), q += C;
    memcpy(q, prob, sizeof(float) * 3), q += 4;
    for (int c = 0; c < C; ++c) q[c] = 1.f;   // pre-mask: none (b200sr_split_set_premask)
    b200sr_split *b = new (std::nothrow) b200sr_split();
    if (!b) return fail(B200SR_E_INVAL, "split_create: out of memory");
    b->c = C;
    b->params.swap(f);
    {
        const char *e = getenv("B200SR_SPLIT_IMPL");
        b->ffma_only = e && !strcmp(e, "ffma");
    }
    if (int rc = split_upload_tc(b)) {
        b200sr_split_destroy(b);
        return rc;
    }
    *out = b;
    return 0;
}

void b200sr_split_destroy(b200sr_split_t *b) {
    if (b && b->d_tc) cudaFree(b->d_tc);
    delete b;
}

int b200sr_split_set_premask(b200sr_split_t *b, const float *g) {
    if (!b) return fail(B200SR_E_INVAL, "split_set_premask: null block");
    float *q = b->params.data() + b->params.size() - b->c;
    for (int c = 0; c < b->c; ++c) q[c] = g ? g[c] : 1.f;
    return split_upload_tc(b);   // (synchronous copy: call outside stream capture, before the forward that uses it)
}

size_t b200sr_nas_workspace_bytes(const b200sr_wdsr_t *p, int n, int h, int w, int precision) {
    if (!p || n <= 0 || h <= 0 || w <= 0) return 0;
    const size_t es = esize(precision), px = (size_t)n * h * w;
    const size_t trunk = ((px * p->cp * es + 255) / 256) * 256, nchw = ((px * p->cin * es + 255) / 256) * 256;
    return trunk + 2 * nchw;
}

int b200sr_nas_forward(const b200sr_wdsr_t *p, const b200sr_split_t *const *blocks, int nblocks, const void *x, int x_dtype, void *y,
                       int y_dtype, int n, int h, int w, int precision, void *ws, size_t ws_bytes, void *stream) {
    int rc = check_common(p, n, h, w, precision, "nas_forward");
    if (rc) return rc;
    if (!x || !y || !ws || (nblocks > 0 && !blocks)) return fail(B200SR_E_INVAL, "nas_forward: null tensor/workspace");
    if (p->nb != 0) return fail(B200SR_E_STATE, "nas_forward: the plan must hold head / tail / skip only (0 classic blocks)");
    for (int i = 0; i < nblocks; ++i)
        if (!blocks[i] || blocks[i]->c != p->cin) return fail(B200SR_E_INVAL, "nas_forward: block %d is null or not %d channels wide", i, p->cin);
    const size_t need = b200sr_nas_workspace_bytes(p, n, h, w, precision);
    if (ws_bytes < need) return fail(B200SR_E_WORKSPACE, "nas_forward: workspace %zu < %zu bytes", ws_bytes, need);
    const size_t es = esize(precision), px = (size_t)n * h * w;
    const size_t trunk_b = ((px * p->cp * es + 255) / 256) * 256, nchw_b = ((px * p->cin * es + 255) / 256) * 256;
    uint8_t *trunk = (uint8_t *)ws, *a = trunk + trunk_b, *b = a + nchw_b;
    const int tl = b200sr_wdsr_trunk_layout(p, precision);   // 0 NHWC, 1 planar-8
    cudaStream_t st = (cudaStream_t)stream;
    int launches = 0;
    if ((rc = b200sr_wdsr_head(p, x, x_dtype, trunk, n, h, w, precision, stream))) return rc;
    ++launches;
    if (nblocks > 0) {   // the Split_Block kernel works on NCHW planes (depthwise convolutions): two layout changes per forward
        CU(launch_trunk_convert(trunk, tl, a, 2, precision, n, p->cin, p->cp, h, w, st));
        for (int i = 0; i < nblocks; ++i) {
            CU(run_split(blocks[i], precision, a, b, n, h, w, st));
            uint8_t *t = a;
            a = b, b = t;
        }
        CU(launch_trunk_convert(a, 2, trunk, tl, precision, n, p->cin, p->cp, h, w, st));
        launches += nblocks + 2;
    }
    if ((rc = b200sr_wdsr_tail(p, trunk, x, x_dtype, y, y_dtype, n, h, w, precision, stream))) return rc;
    p->launches = launches + 1;
    return 0;
}

int b200sr_split_forward(const b200sr_split_t *b, const void *x, void *y, int n, int h, int w, int dtype, void *stream) {
    if (!b || !x || !y) return fail(B200SR_E_INVAL, "split_forward: null argument");
    if (x == y) return fail(B200SR_E_INVAL, "split_forward: x and y must be distinct buffers (a CTA reads a 3-pixel halo that its neighbours overwrite)");
    if (n <= 0 || h <= 0 || w <= 0) return fail(B200SR_E_INVAL, "split_forward: bad shape");
    if (dtype != B200SR_F32 && dtype != B200SR_BF16) return fail(B200SR_E_INVAL, "split_forward: bad dtype %d", dtype);
    cudaError_t e = run_split(b, dtype, x, y, n, h, w, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "split_forward");
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// video path
// ---------------------------------------------------------------------------------------------------------
struct b200sr_conv {
    int cin = 0, cout = 0, k = 0, cinp_f32 = 0, coutp_f32 = 0, cinp_bf16 = 0, coutp_bf16 = 0, nt = 0;
    float *d_w_f32 = nullptr, *d_bias = nullptr;
    uint16_t *d_w_bf16 = nullptr;
    int max_ctas = 0;             // grid cap of the tcgen05 kernels (0 = all SMs)
    uint8_t *d_w_tc5 = nullptr;   // tcgen05 operand image of a 3x3 64 -> 64 filter (conv_tc5.cuh)
};

int b200sr_conv_create(int cin, int cout, int k, const float *w, const float *bias, b200sr_conv_t **out) {
    if (!w || !out) return fail(B200SR_E_INVAL, "conv_create: null argument");
    if (cin < 1 || cout < 1 || (k != 1 && k != 3 && k != 5 && k != 7)) return fail(B200SR_E_UNSUPPORTED, "conv_create: cin=%d cout=%d k=%d (k in {1,3,5,7})", cin, cout, k);
    if (b200sr_device_count() <= 0) return fail(B200SR_E_STATE, "conv_create: no CUDA device (this library has no CPU fallback)");
    b200sr_conv *c = new (std::nothrow) b200sr_conv();
    if (!c) return fail(B200SR_E_INVAL, "conv_create: out of memory");
    c->cin = cin, c->cout = cout, c->k = k;
    c->cinp_f32 = round_up(cin, 8), c->coutp_f32 = round_up(cout, 32);
    c->nt = cout <= 8 ? 1 : cout <= 16 ? 2 : cout <= 32 ? 4 : 8;
    c->cinp_bf16 = round_up(cin, 16), c->coutp_bf16 = round_up(cout, 8 * c->nt);
    const int kk = k * k, bp = c->coutp_f32 > c->coutp_bf16 ? c->coutp_f32 : c->coutp_bf16;
    std::vector<float> wf((size_t)kk * c->cinp_f32 * c->coutp_f32, 0.f), bb((size_t)bp, 0.f);
    std::vector<uint16_t> wb((size_t)kk * c->coutp_bf16 * c->cinp_bf16, 0);
    for (int o = 0; o < cout; ++o) {
        for (int i = 0; i < cin; ++i)
            for (int t = 0; t < kk; ++t) {
                const float v = w[((size_t)o * cin + i) * kk + t];
                wf[((size_t)t * c->cinp_f32 + i) * c->coutp_f32 + o] = v;
                wb[((size_t)t * c->coutp_bf16 + o) * c->cinp_bf16 + i] = f2bf(v);
            }
        if (bias) bb[o] = bias[o];
    }
    int rc;
    if (cin >= 64 && cin <= 80 && cout % 64 == 0 && cout <= 256 && k == 3) {
        // per group of 64 output channels: K-major core matrices [8-output-channel group][(tap, 8-input-channel chunk) slice]
        // [8 rows][16 B] (conv_tc5.cuh); 8 chunks for 64 input channels, 10 (zero rows past cin) for the trunk's first conv on
        // [x_i | warped features]
        const int nch = cin == 64 ? 8 : 10;
        const size_t img = (size_t)8 * 9 * nch * 64;
        std::vector<uint16_t> wi(img * (cout / 64), 0);
        for (int o = 0; o < cout; ++o)
            for (int i = 0; i < cin; ++i)
                for (int t = 0; t < 9; ++t)
                    wi[(o / 64) * img + (((size_t)(o % 64 / 8) * 9 * nch + t * nch + i / 8) * 8 + o % 8) * 8 + i % 8] =
                        f2bf(w[((size_t)o * cin + i) * 9 + t]);
        if ((rc = upload(wi.data(), wi.size() * 2, (void **)&c->d_w_tc5))) {
            b200sr_conv_destroy(c);
            return rc;
        }
    }
    if (cin == 128 && cout % 64 == 0 && cout <= 256 && k == 1) {
        // 1x1 form of the tcgen05 kernel (fusion conv): per group of 64 outputs [8 row groups][16 chunk slices][8 rows][16 B]
        const size_t img = (size_t)8 * 16 * 64;
        std::vector<uint16_t> wi(img * (cout / 64), 0);
        for (int o = 0; o < cout; ++o)
            for (int i = 0; i < 128; ++i) wi[(o / 64) * img + (((size_t)(o % 64 / 8) * 16 + i / 8) * 8 + o % 8) * 8 + i % 8] = f2bf(w[(size_t)o * 128 + i]);
        if ((rc = upload(wi.data(), wi.size() * 2, (void **)&c->d_w_tc5))) {
            b200sr_conv_destroy(c);
            return rc;
        }
    }
    if (cin == 64 && cout == 3 && k == 3) {
        // "rgb" form of the tcgen05 3x3 kernel (conv_last): [2 row groups][72 slices][8 rows][16 B] with the horizontal taps in N --
        // accumulator row n = dx * 3 + c (9 of 16 used), slice (dy, 8-channel chunk) (24 of 72 used): conv_tc5.cuh
        std::vector<uint16_t> wi((size_t)2 * 72 * 64, 0);
        for (int o = 0; o < 3; ++o)
            for (int i = 0; i < 64; ++i)
                for (int t = 0; t < 9; ++t) {
                    const int dy = t / 3, dx = t % 3, n = dx * 3 + o;
                    wi[(((size_t)(n / 8) * 72 + dy * 8 + i / 8) * 8 + n % 8) * 8 + i % 8] = f2bf(w[((size_t)o * 64 + i) * 9 + t]);
                }
        if ((rc = upload(wi.data(), wi.size() * 2, (void **)&c->d_w_tc5))) {
            b200sr_conv_destroy(c);
            return rc;
        }
    }
    if (k == 7 && conv7_tc5_shape_ok(cin, cout)) {
        // [7 tap rows ky][cout/8 row groups][7 taps kx x nch chunks][8 rows][16 B]: one contiguous stage per tap row (conv7_tc5.cuh)
        // G horizontal taps in N (conv7_tc5.cuh, Cfg): accumulator row n = (kx % G) * cout + o, tap group kx / G
        const int nch = conv7_tc5_nch(cin), G = cout == 32 ? 2 : cout == 16 ? 4 : 1, NG = (7 + G - 1) / G, ND = G * cout;
        std::vector<uint16_t> wi((size_t)7 * (ND / 8) * NG * nch * 64, 0);
        for (int o = 0; o < cout; ++o)
            for (int i = 0; i < cin; ++i)
                for (int ky = 0; ky < 7; ++ky)
                    for (int kx = 0; kx < 7; ++kx) {
                        const int n = (kx % G) * cout + o, j = kx / G;
                        wi[(((((size_t)ky * (ND / 8) + n / 8) * NG + j) * nch + i / 8) * 8 + n % 8) * 8 + i % 8] =
                            f2bf(w[((size_t)o * cin + i) * 49 + ky * 7 + kx]);
                    }
        if ((rc = upload(wi.data(), wi.size() * 2, (void **)&c->d_w_tc5))) {
            b200sr_conv_destroy(c);
            return rc;
        }
    }
    if ((rc = upload(wf.data(), wf.size() * 4, (void **)&c->d_w_f32)) || (rc = upload(wb.data(), wb.size() * 2, (void **)&c->d_w_bf16)) ||
        (rc = upload(bb.data(), bb.size() * 4, (void **)&c->d_bias))) {
        b200sr_conv_destroy(c);
        return rc;
    }
    *out = c;
    return 0;
}

void b200sr_conv_destroy(b200sr_conv_t *c) {
    if (!c) return;
    if (c->d_w_f32) cudaFree(c->d_w_f32);
    if (c->d_w_bf16) cudaFree(c->d_w_bf16);
    if (c->d_bias) cudaFree(c->d_bias);
    if (c->d_w_tc5) cudaFree(c->d_w_tc5);
    delete c;
}

int b200sr_conv_set_max_ctas(b200sr_conv_t *c, int max_ctas) {
    if (!c || max_ctas < 0) return fail(B200SR_E_INVAL, "conv_set_max_ctas: bad argument");
    c->max_ctas = max_ctas;
    return 0;
}

int b200sr_conv_tcgen05_ok(const b200sr_conv_t *c) { return c && c->d_w_tc5 && conv_tc5_enabled() ? 1 : 0; }

int b200sr_conv_forward_layout(const b200sr_conv_t *c, const void *x, int x_layout, int x_cs, int x_co, void *y, int y_layout, int y_cs,
                               int y_co, const void *res, int r_cs, int r_co, int n, int h, int w, int act, int shuffle, int in_dtype,
                               int out_dtype, int precision, void *stream) {
    if (!c || !x || !y) return fail(B200SR_E_INVAL, "conv_forward: null argument");
    if (n <= 0 || h <= 0 || w <= 0) return fail(B200SR_E_INVAL, "conv_forward: bad shape");
    if (shuffle != 1 && shuffle != 2) return fail(B200SR_E_UNSUPPORTED, "conv_forward: shuffle %d (1 or 2)", shuffle);
    if (shuffle == 2 && (c->cout % 4)) return fail(B200SR_E_INVAL, "conv_forward: PixelShuffle(2) needs cout %% 4 == 0");
    if (act < 0 || act > 2) return fail(B200SR_E_INVAL, "conv_forward: bad activation %d", act);
    const bool xp = x_layout == B200SR_TRUNK_PLANAR8, yp = y_layout == B200SR_TRUNK_PLANAR8;
    if ((xp && c->cin % 8) || (yp && (shuffle == 2 ? c->cout % 32 : c->cout % 8))) return fail(B200SR_E_INVAL, "conv_forward: planar-8 needs channel counts that are multiples of 8");
    if ((!xp && x_layout != B200SR_TRUNK_NHWC) || (!yp && y_layout != B200SR_TRUNK_NHWC)) return fail(B200SR_E_INVAL, "conv_forward: bad layout");
    if ((!xp && x_co + c->cin > x_cs) || (!yp && y_co + (shuffle == 2 ? c->cout / 4 : c->cout) > y_cs))
        return fail(B200SR_E_INVAL, "conv_forward: channel window outside tensor");
    ConvArgs a;
    a.x = x, a.y = y, a.residual = res, a.bias = c->d_bias;
    a.n = n, a.h = h, a.w_ = w, a.cin = c->cin, a.cout = c->cout, a.x_cs = x_cs, a.x_co = x_co, a.y_cs = y_cs, a.y_co = y_co, a.r_cs = r_cs,
    a.r_co = r_co, a.act = act, a.shuffle = shuffle, a.x_planar = xp, a.y_planar = yp, a.max_ctas = c->max_ctas, a.ks = c->k;
    if (precision == B200SR_F32) {
        a.w = c->d_w_f32, a.cinp = c->cinp_f32, a.coutp = c->coutp_f32;
    } else {
        a.w = c->d_w_bf16, a.cinp = c->cinp_bf16, a.coutp = c->coutp_bf16;
    }
    if (precision == B200SR_BF16 && in_dtype == B200SR_BF16 && out_dtype == B200SR_BF16 && c->d_w_tc5 && conv_tc5_enabled() && c->k == 7 &&
        conv7_tc5_eligible(a)) {
        cudaError_t e7 = launch_conv7x7_tc5(a, c->d_w_tc5, (cudaStream_t)stream);
        if (e7 != cudaSuccess) return cuda_fail(e7, "conv_forward (tcgen05 7x7)");
        return 0;
    }
    if (precision == B200SR_BF16 && in_dtype == B200SR_BF16 && out_dtype == B200SR_BF16 && c->d_w_tc5 && conv_tc5_enabled() &&
        (c->k == 3 || c->k == 1) && conv_tc5_eligible(a)) {
        cudaError_t e5 = launch_conv3x3_c64_tc5(a, c->d_w_tc5, (cudaStream_t)stream);
        if (e5 != cudaSuccess) return cuda_fail(e5, "conv_forward (tcgen05 3x3 -> 64)");
        return 0;
    }
    if (xp || yp)
        return fail(B200SR_E_UNSUPPORTED, "conv_forward: the planar-8 layout is served by the tcgen05 bf16 kernels only (3x3 64 -> 64, SPyNet 7x7 layers)");
    cudaError_t e = launch_conv(a, c->k, c->nt, in_dtype, out_dtype, precision, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "conv_forward (fp32 precision needs float32 tensors; bf16 precision: bf16|f32 in, bf16|f32 out)");
    return 0;
}

int b200sr_vsr_conv_last_base(const b200sr_conv_t *c, const void *x, int x_layout, int x_cs, int x_co, const float *base, int64_t base_nstride,
                              float *y, int64_t y_nstride, int n, int H, int W, void *stream) {
    if (!c || !x || !base || !y) return fail(B200SR_E_INVAL, "vsr_conv_last_base: null argument");
    if (n <= 0 || H <= 0 || W <= 0 || H % 4 || W % 4) return fail(B200SR_E_INVAL, "vsr_conv_last_base: bad shape (H, W multiples of 4)");
    ConvArgs a;
    a.x = x, a.y = y, a.residual = nullptr, a.bias = c->d_bias, a.w = nullptr;
    a.n = n, a.h = H, a.w_ = W, a.cin = c->cin, a.cinp = c->cinp_bf16, a.cout = c->cout, a.coutp = c->coutp_bf16, a.x_cs = x_cs, a.x_co = x_co,
    a.y_cs = 0, a.y_co = 0, a.r_cs = 0, a.r_co = 0, a.act = 0, a.shuffle = 1, a.max_ctas = c->max_ctas;
    a.x_planar = x_layout == B200SR_TRUNK_PLANAR8, a.y_planar = 0;
    a.base = base, a.base_nstride = base_nstride, a.y_nstride = y_nstride;
    if (!c->d_w_tc5 || c->k != 3 || !conv_tc5_enabled() || !conv_tc5_eligible(a))
        return fail(B200SR_E_UNSUPPORTED, "vsr_conv_last_base: needs a 3x3 64 -> 3 conv, bf16 NHWC (16-byte aligned window) or planar-8 input");
    cudaError_t e = launch_conv3x3_c64_tc5(a, c->d_w_tc5, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "vsr_conv_last_base");
    return 0;
}

int b200sr_vsr_trunk_forward(const b200sr_conv_t *first, const b200sr_conv_t *const *blocks, int num_block, const void *buf, int buf_cs,
                             void *t, void *o, void *out, int n, int h, int w, void *stream) {
    return b200sr_vsr_trunk_forward_into(first, blocks, num_block, buf, buf_cs, t, o, out, 64, 0, n, h, w, stream);
}

int b200sr_vsr_trunk_forward_into(const b200sr_conv_t *first, const b200sr_conv_t *const *blocks, int num_block, const void *buf, int buf_cs,
                                  void *t, void *o, void *out, int out_cs, int out_co, int n, int h, int w, void *stream) {
    if (!first || !blocks || num_block < 1 || !buf || !t || !o || !out) return fail(B200SR_E_INVAL, "vsr_trunk_forward: bad argument");
    if (out_cs < 64 || out_co < 0 || out_co + 64 > out_cs || out_cs % 8 || out_co % 8) return fail(B200SR_E_INVAL, "vsr_trunk_forward: output channel window");
    const int P = B200SR_TRUNK_PLANAR8, N_ = B200SR_TRUNK_NHWC, bf = B200SR_BF16;
    int rc = b200sr_conv_forward_layout(first, buf, N_, buf_cs, 0, t, P, 64, 0, nullptr, 0, 0, n, h, w, B200SR_ACT_LRELU01, 1, bf, bf, bf, stream);
    for (int k = 0; k < num_block && !rc; ++k) {
        rc = b200sr_conv_forward_layout(blocks[2 * k], t, P, 64, 0, o, P, 64, 0, nullptr, 0, 0, n, h, w, B200SR_ACT_RELU, 1, bf, bf, bf, stream);
        if (rc) break;
        const bool last = k + 1 == num_block;   // conv2 adds its residual in place (y = t + conv(o)); the last one writes the NHWC features
        rc = b200sr_conv_forward_layout(blocks[2 * k + 1], o, P, 64, 0, last ? out : t, last ? N_ : P, last ? out_cs : 64, last ? out_co : 0, t, 64, 0,
                                        n, h, w, B200SR_ACT_NONE, 1, bf, bf, bf, stream);
    }
    return rc;
}

int b200sr_conv_forward(const b200sr_conv_t *c, const void *x, int x_cs, int x_co, void *y, int y_cs, int y_co, const void *res, int r_cs,
                        int r_co, int n, int h, int w, int act, int shuffle, int in_dtype, int out_dtype, int precision, void *stream) {
    return b200sr_conv_forward_layout(c, x, B200SR_TRUNK_NHWC, x_cs, x_co, y, B200SR_TRUNK_NHWC, y_cs, y_co, res, r_cs, r_co, n, h, w, act,
                                      shuffle, in_dtype, out_dtype, precision, stream);
}

int b200sr_resize_bilinear_nchw(const void *x, int x_dtype, float *y, int n, int c, int h, int w, int oh, int ow, int align,
                                const float *sub4, const float *mul4, void *stream) {
    if (!x || !y) return fail(B200SR_E_INVAL, "resize_bilinear: null tensor");
    if (h <= 0 || w <= 0 || oh <= 0 || ow <= 0) return fail(B200SR_E_INVAL, "resize_bilinear: bad shape");
    const float z[4] = {0, 0, 0, 0}, o[4] = {1, 1, 1, 1};
    CU(launch_resize_bilinear_nchw(x, x_dtype, y, n, c, h, w, oh, ow, align, sub4 ? sub4 : z, mul4 ? mul4 : o, (cudaStream_t)stream));
    return 0;
}
int b200sr_vsr_deconv_tail(const void *t, int t_dtype, int t_cstride, const void *img, int img_dtype, int64_t img_nstride, float *y,
                           int64_t y_nstride, int n, int h, int w, int oh, int ow, void *stream) {
    if (!t || !img || !y) return fail(B200SR_E_INVAL, "vsr_deconv_tail: null tensor");
    if (n < 0 || h <= 0 || w <= 0 || oh <= 0 || ow <= 0 || t_cstride < 48) return fail(B200SR_E_INVAL, "vsr_deconv_tail: bad shape");
    CU(launch_deconv_tail_resize_add(t, t_dtype, t_cstride, img, img_dtype, img_nstride, y, y_nstride, n, h, w, oh, ow, (cudaStream_t)stream));
    return 0;
}
int b200sr_u8_to_unit(const uint8_t *x, void *y, int y_dtype, int64_t count, void *stream) {
    if (!x || !y || count < 0) return fail(B200SR_E_INVAL, "u8_to_unit: bad argument");
    if (y_dtype != B200SR_F32 && y_dtype != B200SR_BF16) return fail(B200SR_E_INVAL, "u8_to_unit: y_dtype %d", y_dtype);
    CU(launch_u8_to_unit(x, y, y_dtype, count, (cudaStream_t)stream));
    return 0;
}
int b200sr_ssd_u8(const uint8_t *a, const uint8_t *b, uint64_t *out, int n, int c, int h, int w, int shave, void *stream) {
    if (!a || !b || !out) return fail(B200SR_E_INVAL, "ssd_u8: null tensor");
    if (n < 0 || c <= 0 || shave < 0 || h - 2 * shave <= 0 || w - 2 * shave <= 0) return fail(B200SR_E_INVAL, "ssd_u8: bad shape / shave");
    CU(launch_ssd_u8(a, b, (unsigned long long *)out, n, c, h, w, shave, (cudaStream_t)stream));
    return 0;
}
int b200sr_avg_pool2_nchw(const float *x, float *y, int n, int c, int h, int w, void *stream) {
    if (!x || !y) return fail(B200SR_E_INVAL, "avg_pool2: null tensor");
    CU(launch_avg_pool2(x, y, n, c, h, w, (cudaStream_t)stream));
    return 0;
}
// ---------------------------------------------------------------------------------------------------------
// SpyNet.forward (models/spynet_arch.py:49-96) as ONE call: resize + normalise, 5 x avg-pool, six pyramid levels of
// (upsample x2 -> warp(border) -> concat8 -> 7x7 convs 8-32-64-32-16-2 -> + up), final resize + rescale.  ~100 launches sequenced in C on
// a caller-provided workspace; nothing is allocated, nothing synchronises.
// ---------------------------------------------------------------------------------------------------------
namespace {
struct SpyGeom {
    int h_up, w_up, hl[6], wl[6];
    size_t pyr_off[2][6], inp_off, up_off, flow_off[2], act_off[2], head_off, total;
};
inline size_t al256(size_t v) { return (v + 255) / 256 * 256; }
SpyGeom spynet_geom(int n, int h, int w, int precision) {
    SpyGeom g;
    g.h_up = (h + 31) / 32 * 32, g.w_up = (w + 31) / 32 * 32;
    g.hl[5] = g.h_up, g.wl[5] = g.w_up;
    for (int l = 4; l >= 0; --l) g.hl[l] = g.hl[l + 1] / 2, g.wl[l] = g.wl[l + 1] / 2;
    const size_t es = precision == B200SR_F32 ? 4 : 2;
    const int cs = precision == B200SR_F32 ? 8 : 16;
    size_t off = 0;
    for (int k = 0; k < 2; ++k)
        for (int l = 0; l < 6; ++l) g.pyr_off[k][l] = off, off += al256((size_t)n * 3 * g.hl[l] * g.wl[l] * 4);
    const size_t px = (size_t)n * g.h_up * g.w_up;
    g.inp_off = off, off += al256(px * cs * es);
    g.up_off = off, off += al256(px * 2 * 4);
    for (int k = 0; k < 2; ++k) g.flow_off[k] = off, off += al256(px * 2 * 4);
    for (int k = 0; k < 2; ++k) g.act_off[k] = off, off += al256(px * 64 * es);
    g.head_off = off, off += al256(px * 2 * 4);
    g.total = off;
    return g;
}
}  // namespace

size_t b200sr_spynet_workspace_bytes(int n, int h, int w, int precision) {
    if (n <= 0 || h <= 0 || w <= 0) return 0;
    return spynet_geom(n, h, w, precision).total;
}

int b200sr_spynet_forward(const b200sr_conv_t *const *convs, const void *ref, const void *supp, int img_dtype, float *flow_out, int n, int h,
                          int w, int precision, const float *mean4, const float *inv_std4, void *ws, size_t ws_bytes, void *stream) {
    if (!convs || !ref || !supp || !flow_out || !ws || !mean4 || !inv_std4) return fail(B200SR_E_INVAL, "spynet_forward: null argument");
    if (n <= 0 || h <= 0 || w <= 0) return fail(B200SR_E_INVAL, "spynet_forward: bad shape");
    if (precision != B200SR_F32 && precision != B200SR_BF16) return fail(B200SR_E_INVAL, "spynet_forward: bad precision %d", precision);
    for (int i = 0; i < 30; ++i)
        if (!convs[i]) return fail(B200SR_E_INVAL, "spynet_forward: conv %d is null (30 handles: level-major, layers 8-32-64-32-16-2)", i);
    const SpyGeom g = spynet_geom(n, h, w, precision);
    if (g.h_up < 64 || g.w_up < 64) return fail(B200SR_E_INVAL, "spynet_forward: SPyNet needs at least 33 pixels per side");
    if (ws_bytes < g.total) return fail(B200SR_E_WORKSPACE, "spynet_forward: workspace %zu < %zu bytes", ws_bytes, g.total);
    uint8_t *base = (uint8_t *)ws;
    cudaStream_t st = (cudaStream_t)stream;
    const int adt = precision, cs = precision == B200SR_F32 ? 8 : 16;
    const int N_ = B200SR_TRUNK_NHWC, P = B200SR_TRUNK_PLANAR8;
    const void *img[2] = {ref, supp};
    for (int k = 0; k < 2; ++k) {   // resize (align_corners=False) + ImageNet normalisation, then the 5 average pools  (:88-89, 45-47, 52-57)
        CU(launch_resize_bilinear_nchw(img[k], img_dtype, (float *)(base + g.pyr_off[k][5]), n, 3, h, w, g.h_up, g.w_up, 0, mean4, inv_std4, st));
        for (int l = 4; l >= 0; --l)
            CU(launch_avg_pool2((const float *)(base + g.pyr_off[k][l + 1]), (float *)(base + g.pyr_off[k][l]), n, 3, g.hl[l + 1], g.wl[l + 1], st));
    }
    const float *flow = nullptr;
    int ph = g.hl[0] / 2, pw = g.wl[0] / 2, rc;
    for (int l = 0; l < 6; ++l) {
        const int hl = g.hl[l], wl = g.wl[l];
        float *up = (float *)(base + g.up_off);
        CU(launch_spynet_level_input((const float *)(base + g.pyr_off[0][l]), (const float *)(base + g.pyr_off[1][l]), flow, base + g.inp_off, adt,
                                     up, n, hl, wl, ph, pw, cs, st));
        const b200sr_conv_t *const *L = convs + 5 * l;
        // bf16: layers 0-3 run on the tcgen05 7x7 kernel and keep their private tensors planar-8; the 16 -> 2 flow head reads NHWC
        bool planar = precision != B200SR_F32;
        for (int j = 0; j < 4 && planar; ++j) planar = b200sr_conv_tcgen05_ok(L[j]) != 0;
        const void *t = base + g.inp_off;
        int tcs = cs;
        for (int j = 0; j < 5; ++j) {
            const bool last = j == 4;
            void *y = last ? (void *)(base + g.head_off) : (void *)(base + g.act_off[j & 1]);
            const int xl = planar && j >= 1 && j <= 3 ? P : N_, yl = planar && j <= 2 ? P : N_;
            const int ycs = L[j]->cout;
            if ((rc = b200sr_conv_forward_layout(L[j], t, xl, xl == P ? L[j]->cin : tcs, 0, y, yl, ycs, 0, nullptr, 0, 0, n, hl, wl,
                                                 last ? B200SR_ACT_NONE : B200SR_ACT_RELU, 1, adt, last ? B200SR_F32 : adt, precision, stream)))
                return rc;
            t = y, tcs = ycs;
        }
        float *fnew = (float *)(base + g.flow_off[l & 1]);
        CU(launch_nhwc_plus_nchw((const float *)(base + g.head_off), up, fnew, n, 2, hl, wl, 2, st));
        flow = fnew, ph = hl, pw = wl;
    }
    const float zero[4] = {0, 0, 0, 0}, scale[4] = {(float)w / (float)g.w_up, (float)h / (float)g.h_up, 1.f, 1.f};   // (:91-94)
    CU(launch_resize_bilinear_nchw(flow, B200SR_F32, flow_out, n, 2, g.h_up, g.w_up, h, w, 0, zero, scale, st));
    return 0;
}

int b200sr_spynet_level_input(const float *ref, const float *supp, const float *flow_prev, void *out, int out_dtype, float *up, int n,
                              int h, int w, int ph, int pw, int cs, void *stream) {
    if (!ref || !supp || !out || !up) return fail(B200SR_E_INVAL, "spynet_level_input: null tensor");
    if (cs < 8 || ph < 1 || pw < 1 || h < 2 * ph || h > 2 * ph + 1 || w < 2 * pw || w > 2 * pw + 1)
        return fail(B200SR_E_INVAL, "spynet_level_input: level %dx%d does not follow previous flow %dx%d", h, w, ph, pw);
    CU(launch_spynet_level_input(ref, supp, flow_prev, out, out_dtype, up, n, h, w, ph, pw, cs, (cudaStream_t)stream));
    return 0;
}
int b200sr_nhwc_plus_nchw(const float *a, const float *b, float *y, int n, int c, int h, int w, int cs, void *stream) {
    if (!a || !y) return fail(B200SR_E_INVAL, "nhwc_plus_nchw: null tensor");
    CU(launch_nhwc_plus_nchw(a, b, y, n, c, h, w, cs, (cudaStream_t)stream));
    return 0;
}
int b200sr_zero_async(void *dst, size_t bytes, void *stream) {
    if (!dst && bytes) return fail(B200SR_E_INVAL, "zero_async: null pointer");
    if (bytes) CU(cudaMemsetAsync(dst, 0, bytes, (cudaStream_t)stream));
    return 0;
}
int b200sr_pad_bottom_right_async(const void *src, void *dst, int n, int h, int w, int px_bytes, void *stream) {
    if (!src || !dst || n < 0 || h <= 0 || w <= 0 || px_bytes <= 0) return fail(B200SR_E_INVAL, "pad_bottom_right_async: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t spitch = (size_t)w * px_bytes, dpitch = (size_t)(w + 1) * px_bytes;
    for (int i = 0; i < n; ++i) {
        const uint8_t *s = (const uint8_t *)src + (size_t)i * h * spitch;
        uint8_t *d = (uint8_t *)dst + (size_t)i * (h + 1) * dpitch;
        CU(cudaMemcpy2DAsync(d, dpitch, s, spitch, spitch, (size_t)h, cudaMemcpyDeviceToDevice, st));
        CU(cudaMemset2DAsync(d + spitch, dpitch, 0, (size_t)px_bytes, (size_t)h, st));   // the zero column
        CU(cudaMemsetAsync(d + (size_t)h * dpitch, 0, dpitch, st));                       // the zero row
    }
    return 0;
}
int b200sr_nchw3_to_nhwc(const void *x, int x_dtype, int64_t x_nstride, void *y, int y_dtype, int n, int h, int w, int cs, int co, void *stream) {
    if (!x || !y || co + 3 > cs) return fail(B200SR_E_INVAL, "nchw3_to_nhwc: bad argument");
    CU(launch_nchw3_to_nhwc(x, x_dtype, x_nstride, y, y_dtype, n, h, w, cs, co, (cudaStream_t)stream));
    return 0;
}
int b200sr_vsr_base_add(const void *a, int a_dtype, int cs, const void *img, int img_dtype, int64_t img_nstride, float *y, int64_t y_nstride,
                        int n, int h, int w, void *stream) {
    if (!a || !img || !y || cs < 3) return fail(B200SR_E_INVAL, "vsr_base_add: bad argument");
    CU(launch_vsr_base_add(a, a_dtype, cs, img, img_dtype, img_nstride, y, y_nstride, n, h, w, (cudaStream_t)stream));
    return 0;
}

int b200sr_vsr_shuffle4_base_add(const void *a, int a_dtype, int cs, const void *img, int img_dtype, int64_t img_nstride, float *y, int64_t y_nstride,
                                 int n, int h, int w, void *stream) {
    if (!a || !img || !y || cs < 48) return fail(B200SR_E_INVAL, "vsr_shuffle4_base_add: bad argument (a has at least 3 x 16 channels)");
    CU(launch_vsr_base_add(a, a_dtype, cs, img, img_dtype, img_nstride, y, y_nstride, n, h, w, (cudaStream_t)stream, true));
    return 0;
}

}  // extern "C"
