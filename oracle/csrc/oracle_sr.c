/*
 * oracle_sr.c -- plain-C restatement of the reference's super-resolution
 * forward arithmetic.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py):
 * never linked into, loaded by or called from the product library.
 *
 * Direct loops, float32 storage, float64 accumulation, NCHW like the
 * reference's tensors.  Each primitive cites the reference call site it
 * restates (paths relative to /root/reference).  The composition of these
 * primitives into whole forwards lives in oracle/c_oracle.py.
 *
 * Parity: PINNED against tests/golden/ (generated from the imported reference
 * by oracle/make_golden.py).  float64 accumulation puts this oracle within
 * ~2e-5 of the reference's own float32 result at random-init dynamic range
 * (SURVEY.md App. C "reference fp32 vs same model in fp64").
 *
 * Build: oracle/Makefile  ->  oracle/_build/liboracle_sr.so
 */
#include <math.h>
#include <stddef.h>
#include <string.h>

#define IDX4(n, c, h, w, C, H, W) ((((size_t)(n) * (C) + (c)) * (H) + (h)) * (W) + (w))

/* W[o,:] = g[o] * v[o,:] / ||v[o,:]||_2
 * torch.nn.utils.weight_norm(dim=0): models/basic_wdsr_b.py:23,32,55,68,108,119,129 */
void osr_weight_norm(const float *g, const float *v, int O, int inner, float *w) {
    for (int o = 0; o < O; ++o) {
        double s = 0.0;
        for (int i = 0; i < inner; ++i) s += (double)v[(size_t)o * inner + i] * v[(size_t)o * inner + i];
        float norm = (float)sqrt(s);
        float scale = g[o] / norm;
        for (int i = 0; i < inner; ++i) w[(size_t)o * inner + i] = v[(size_t)o * inner + i] * scale;
    }
}

/* y = act(conv2d(x, w, b, stride 1, zero pad K/2)), act: 0 none, 1 ReLU, 2 LeakyReLU(0.1).
 * nn.Conv2d call sites: models/basic_wdsr_b.py:32-42,108-138,55-78; models/spynet_arch.py:17-22;
 * models/basicvsr_arch_origin.py:110-137. */
void osr_conv2d(const float *x, int N, int C, int H, int W, const float *w, const float *b, int O, int K, int act,
                float *y) {
    const int P = K / 2;
#pragma omp parallel for collapse(2) schedule(static)
    for (int n = 0; n < N; ++n)
        for (int o = 0; o < O; ++o)
            for (int h = 0; h < H; ++h)
                for (int x0 = 0; x0 < W; ++x0) {
                    double acc = b ? (double)b[o] : 0.0;
                    for (int c = 0; c < C; ++c)
                        for (int ky = 0; ky < K; ++ky) {
                            int yy = h + ky - P;
                            if (yy < 0 || yy >= H) continue;
                            for (int kx = 0; kx < K; ++kx) {
                                int xx = x0 + kx - P;
                                if (xx < 0 || xx >= W) continue;
                                acc += (double)x[IDX4(n, c, yy, xx, C, H, W)] *
                                       (double)w[(((size_t)o * C + c) * K + ky) * K + kx];
                            }
                        }
                    float r = (float)acc;
                    if (act == 1) r = r > 0.f ? r : 0.f;
                    if (act == 2) r = r > 0.f ? r : 0.1f * r;
                    y[IDX4(n, o, h, x0, O, H, W)] = r;
                }
}

/* out[n,c,r*h+i,r*w+j] = in[n,c*r*r+i*r+j,h,w] + add    nn.PixelShuffle, models/basic_wdsr_b.py:80-83,91-92 */
void osr_pixel_shuffle_add(const float *in, int N, int C, int H, int W, int r, float add, float *out) {
    const int Co = C / (r * r);
    for (int n = 0; n < N; ++n)
        for (int c = 0; c < Co; ++c)
            for (int h = 0; h < H; ++h)
                for (int i = 0; i < r; ++i)
                    for (int w = 0; w < W; ++w)
                        for (int j = 0; j < r; ++j)
                            out[IDX4(n, c, r * h + i, r * w + j, Co, r * H, r * W)] =
                                in[IDX4(n, c * r * r + i * r + j, h, w, C, H, W)] + add;
}

/* flow_warp, models/spynet_arch.py:98-129: grid_sample(bilinear, align_corners=True) at (x+fx, y+fy).
 * flow is (N,H,W,2).  border!=0 -> padding_mode='border' (clamp the coordinate), else 'zeros'
 * (out-of-image corners contribute 0).  The round trip through the [-1,1] normalisation
 * (:123-124 and grid_sample's un-normalisation) is replayed in float32 so the sample
 * position matches the reference to the last ulp. */
void osr_flow_warp(const float *x, int N, int C, int H, int W, const float *flow, int border, float *y) {
    const float dw = (float)(W - 1 > 1 ? W - 1 : 1), dh = (float)(H - 1 > 1 ? H - 1 : 1);
#pragma omp parallel for collapse(2) schedule(static)
    for (int n = 0; n < N; ++n)
        for (int h = 0; h < H; ++h)
            for (int w = 0; w < W; ++w) {
                const float *f = flow + (((size_t)n * H + h) * W + w) * 2;
                float gx = 2.0f * ((float)w + f[0]) / dw - 1.0f;
                float gy = 2.0f * ((float)h + f[1]) / dh - 1.0f;
                float sx = ((gx + 1.f) / 2.f) * (float)(W - 1);
                float sy = ((gy + 1.f) / 2.f) * (float)(H - 1);
                if (border) {
                    sx = fminf(fmaxf(sx, 0.f), (float)(W - 1));
                    sy = fminf(fmaxf(sy, 0.f), (float)(H - 1));
                }
                float fx0 = floorf(sx), fy0 = floorf(sy);
                int x0 = (int)fx0, y0 = (int)fy0, x1 = x0 + 1, y1 = y0 + 1;
                float ax = sx - fx0, ay = sy - fy0;
                float w00 = (1.f - ax) * (1.f - ay), w01 = ax * (1.f - ay), w10 = (1.f - ax) * ay, w11 = ax * ay;
                int vx0 = x0 >= 0 && x0 < W, vx1 = x1 >= 0 && x1 < W, vy0 = y0 >= 0 && y0 < H, vy1 = y1 >= 0 && y1 < H;
                for (int c = 0; c < C; ++c) {
                    const float *p = x + ((size_t)n * C + c) * H * W;
                    double acc = 0.0;
                    if (vy0 && vx0) acc += (double)p[(size_t)y0 * W + x0] * w00;
                    if (vy0 && vx1) acc += (double)p[(size_t)y0 * W + x1] * w01;
                    if (vy1 && vx0) acc += (double)p[(size_t)y1 * W + x0] * w10;
                    if (vy1 && vx1) acc += (double)p[(size_t)y1 * W + x1] * w11;
                    y[IDX4(n, c, h, w, C, H, W)] = (float)acc;
                }
            }
}

/* F.interpolate(mode='bilinear'), both align_corners conventions:
 *   align_corners=0: src = max(0, (d+0.5)*in/out - 0.5)   models/spynet_arch.py:88-91, basicvsr_arch_origin.py:91,93
 *   align_corners=1: src = d*(in-1)/(out-1)                models/spynet_arch.py:65 */
void osr_resize_bilinear(const float *x, int N, int C, int H, int W, int OH, int OW, int align, float *y) {
    float sh, sw;
    if (align) {
        sh = OH > 1 ? (float)(H - 1) / (float)(OH - 1) : 0.f;
        sw = OW > 1 ? (float)(W - 1) / (float)(OW - 1) : 0.f;
    } else {
        sh = (float)H / (float)OH;
        sw = (float)W / (float)OW;
    }
#pragma omp parallel for schedule(static)
    for (int nc = 0; nc < N * C; ++nc) {
        const float *p = x + (size_t)nc * H * W;
        float *q = y + (size_t)nc * OH * OW;
        for (int oh = 0; oh < OH; ++oh) {
            float sy = align ? sh * (float)oh : fmaxf(sh * ((float)oh + 0.5f) - 0.5f, 0.f);
            int y0 = (int)sy;
            int y1 = y0 + (y0 < H - 1 ? 1 : 0);
            float ly = sy - (float)y0;
            for (int ow = 0; ow < OW; ++ow) {
                float sx = align ? sw * (float)ow : fmaxf(sw * ((float)ow + 0.5f) - 0.5f, 0.f);
                int x0 = (int)sx;
                int x1 = x0 + (x0 < W - 1 ? 1 : 0);
                float lx = sx - (float)x0;
                double v = (double)(1.f - ly) * ((double)(1.f - lx) * p[(size_t)y0 * W + x0] + (double)lx * p[(size_t)y0 * W + x1]) +
                           (double)ly * ((double)(1.f - lx) * p[(size_t)y1 * W + x0] + (double)lx * p[(size_t)y1 * W + x1]);
                q[(size_t)oh * OW + ow] = (float)v;
            }
        }
    }
}

/* F.avg_pool2d(kernel 2, stride 2, count_include_pad=False): models/spynet_arch.py:56-57.  Odd tail dropped. */
void osr_avg_pool2(const float *x, int N, int C, int H, int W, float *y) {
    const int OH = H / 2, OW = W / 2;
    for (int nc = 0; nc < N * C; ++nc) {
        const float *p = x + (size_t)nc * H * W;
        float *q = y + (size_t)nc * OH * OW;
        for (int h = 0; h < OH; ++h)
            for (int w = 0; w < OW; ++w)
                q[(size_t)h * OW + w] = (float)(((double)p[(size_t)(2 * h) * W + 2 * w] + p[(size_t)(2 * h) * W + 2 * w + 1] +
                                                 p[(size_t)(2 * h + 1) * W + 2 * w] + p[(size_t)(2 * h + 1) * W + 2 * w + 1]) * 0.25);
    }
}
