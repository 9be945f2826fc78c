// video_glue.cu -- the memory-bound kernels around the convolutions of the video path (SPyNet pyramid, BasicVSR glue).
// Every resampling formula is the closed form of the ATen op the reference calls; coordinates and
// flows are always fp32.
#include "common.cuh"
#include "launch.h"

namespace b200sr {

__device__ __forceinline__ void src_index(float scale, int d, int in_size, bool align, int &i0, int &i1, float &l) {
    const float s = align ? scale * (float)d : fmaxf(scale * ((float)d + 0.5f) - 0.5f, 0.f);
    i0 = min((int)s, in_size - 1);
    i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
    l = s - (float)i0;
}
__host__ __device__ inline float resize_scale(int in, int out, bool align) {
    if (align) return out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f;
    return (float)in / (float)out;
}

// F.interpolate(mode='bilinear') on NCHW, then y = (y - sub[c]) * mul[c]  (c indexes a 4-entry table, c % 4).
//   SPyNet pre-resize + ImageNet normalisation   models/spynet_arch.py:88-89,45-47   (align_corners=False)
//   SPyNet final flow resize + rescale           models/spynet_arch.py:91-94
//   BasicVSR final resize                        models/basicvsr_arch_origin.py:93
template <typename TIN>
__global__ void __launch_bounds__(256) resize_bilinear_nchw_kernel(const TIN *__restrict__ x, float *__restrict__ y, int NC, int C, int H,
                                                                   int W, int OH, int OW, int align, float4 sub, float4 mul) {
    const float sh = resize_scale(H, OH, align), sw = resize_scale(W, OW, align);
    const long long total = (long long)NC * OH * OW;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ox = (int)(i % OW), oy = (int)((i / OW) % OH), nc = (int)(i / ((long long)OW * OH));
        int y0, y1, x0, x1;
        float ly, lx;
        src_index(sh, oy, H, align, y0, y1, ly);
        src_index(sw, ox, W, align, x0, x1, lx);
        const TIN *p = x + (long long)nc * H * W;
        const float v = (1.f - ly) * ((1.f - lx) * to_f32<TIN>(p[(long long)y0 * W + x0]) + lx * to_f32<TIN>(p[(long long)y0 * W + x1])) +
                        ly * ((1.f - lx) * to_f32<TIN>(p[(long long)y1 * W + x0]) + lx * to_f32<TIN>(p[(long long)y1 * W + x1]));
        const int c = (nc % C) & 3;
        const float s = c == 0 ? sub.x : c == 1 ? sub.y : c == 2 ? sub.z : sub.w;
        const float m = c == 0 ? mul.x : c == 1 ? mul.y : c == 2 ? mul.z : mul.w;
        y[i] = (v - s) * m;
    }
}

cudaError_t launch_resize_bilinear_nchw(const void *x, int x_dtype, float *y, int n, int c, int h, int w, int oh, int ow, int align,
                                        const float *sub4, const float *mul4, cudaStream_t st) {
    const long long total = (long long)n * c * oh * ow;
    if (total == 0) return cudaSuccess;
    long long blocks = (total + 255) / 256;
    if (blocks > (long long)sm_count() * 16) blocks = (long long)sm_count() * 16;
    const float4 s = make_float4(sub4[0], sub4[1], sub4[2], sub4[3]), m = make_float4(mul4[0], mul4[1], mul4[2], mul4[3]);
    if (x_dtype == kF32)
        resize_bilinear_nchw_kernel<float><<<(unsigned)blocks, 256, 0, st>>>((const float *)x, y, n * c, c, h, w, oh, ow, align, s, m);
    else
        resize_bilinear_nchw_kernel<bf16><<<(unsigned)blocks, 256, 0, st>>>((const bf16 *)x, y, n * c, c, h, w, oh, ow, align, s, m);
    return cudaGetLastError();
}

// F.avg_pool2d(2, 2, count_include_pad=False) on NCHW fp32       models/spynet_arch.py:56-57
__global__ void __launch_bounds__(256) avg_pool2_kernel(const float *__restrict__ x, float *__restrict__ y, int NC, int H, int W) {
    const int OH = H / 2, OW = W / 2;
    const long long total = (long long)NC * OH * OW;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ox = (int)(i % OW), oy = (int)((i / OW) % OH), nc = (int)(i / ((long long)OW * OH));
        const float *p = x + ((long long)nc * H + 2 * oy) * W + 2 * ox;
        y[i] = (p[0] + p[1] + p[W] + p[W + 1]) * 0.25f;
    }
}
cudaError_t launch_avg_pool2(const float *x, float *y, int n, int c, int h, int w, cudaStream_t st) {
    const long long total = (long long)n * c * (h / 2) * (w / 2);
    if (total == 0) return cudaSuccess;
    long long blocks = (total + 255) / 256;
    if (blocks > (long long)sm_count() * 16) blocks = (long long)sm_count() * 16;
    avg_pool2_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, y, n * c, h, w);
    return cudaGetLastError();
}

// One SPyNet pyramid level's network input, fused (models/spynet_arch.py:64-78):
//   up   = interpolate(flow_prev, x2, bilinear, align_corners=True) * 2   (+ replicate pad by one row/col if the level is odd)
//   warp = flow_warp(supp, up, padding 'border')
//   out  = NHWC [ref(3) | warp(3) | up(2) | zero pad to CS channels],  up also kept NCHW fp32 for the residual
// flow_prev == nullptr means the zero flow of the coarsest level.
template <typename TOUT>
__global__ void __launch_bounds__(256) spynet_level_input_kernel(const float *__restrict__ ref, const float *__restrict__ supp,
                                                                 const float *__restrict__ flow_prev, TOUT *__restrict__ out,
                                                                 float *__restrict__ up, int N, int H, int W, int PH, int PW, int CS) {
    const long long P = (long long)N * H * W;
    const float sh = resize_scale(PH, 2 * PH, true), sw = resize_scale(PW, 2 * PW, true);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (long long)gridDim.x * blockDim.x) {
        const int xw = (int)(i % W), yh = (int)((i / W) % H), n = (int)(i / ((long long)W * H));
        float fu = 0.f, fv = 0.f;
        if (flow_prev) {
            const int uy = min(yh, 2 * PH - 1), ux = min(xw, 2 * PW - 1);  // replicate pad of the x2 upsampled flow
            int y0, y1, x0, x1;
            float ly, lx;
            src_index(sh, uy, PH, true, y0, y1, ly);
            src_index(sw, ux, PW, true, x0, x1, lx);
            const float *p = flow_prev + (long long)n * 2 * PH * PW;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const float *q = p + (long long)c * PH * PW;
                const float v = (1.f - ly) * ((1.f - lx) * q[y0 * PW + x0] + lx * q[y0 * PW + x1]) +
                                ly * ((1.f - lx) * q[y1 * PW + x0] + lx * q[y1 * PW + x1]);
                (c == 0 ? fu : fv) = v * 2.0f;
            }
        }
        // flow_warp(supp, up, border): same fp32 normalise / un-normalise round trip as the reference
        const float dw = (float)max(W - 1, 1), dh = (float)max(H - 1, 1);
        float ix = ((2.0f * ((float)xw + fu) / dw - 1.0f + 1.f) / 2.f) * (float)(W - 1);
        float iy = ((2.0f * ((float)yh + fv) / dh - 1.0f + 1.f) / 2.f) * (float)(H - 1);
        ix = fminf((float)(W - 1), fmaxf(ix, 0.f));
        iy = fminf((float)(H - 1), fmaxf(iy, 0.f));
        const float fx0 = floorf(ix), fy0 = floorf(iy);
        const int x0 = (int)fx0, y0 = (int)fy0;
        const float ex = (fx0 + 1.f) - ix, wx = ix - fx0, ey = (fy0 + 1.f) - iy, wy = iy - fy0;
        const bool vx1 = x0 + 1 < W, vy1 = y0 + 1 < H;
        TOUT *o = out + i * CS;
        const long long plane = (long long)H * W;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            o[c] = from_f32<TOUT>(ref[((long long)n * 3 + c) * plane + (long long)yh * W + xw]);
            const float *s = supp + ((long long)n * 3 + c) * plane;
            float v = s[(long long)y0 * W + x0] * (ex * ey);
            if (vx1) v += s[(long long)y0 * W + x0 + 1] * (wx * ey);
            if (vy1) v += s[(long long)(y0 + 1) * W + x0] * (ex * wy);
            if (vx1 && vy1) v += s[(long long)(y0 + 1) * W + x0 + 1] * (wx * wy);
            o[3 + c] = from_f32<TOUT>(v);
        }
        o[6] = from_f32<TOUT>(fu);
        o[7] = from_f32<TOUT>(fv);
        for (int c = 8; c < CS; ++c) o[c] = from_f32<TOUT>(0.f);
        up[((long long)n * 2) * plane + (long long)yh * W + xw] = fu;
        up[((long long)n * 2 + 1) * plane + (long long)yh * W + xw] = fv;
    }
}
cudaError_t launch_spynet_level_input(const float *ref, const float *supp, const float *flow_prev, void *out, int out_dtype, float *up,
                                      int n, int h, int w, int ph, int pw, int cs, cudaStream_t st) {
    const long long P = (long long)n * h * w;
    if (P == 0) return cudaSuccess;
    long long blocks = (P + 255) / 256;
    if (blocks > (long long)sm_count() * 16) blocks = (long long)sm_count() * 16;
    if (out_dtype == kF32)
        spynet_level_input_kernel<float><<<(unsigned)blocks, 256, 0, st>>>(ref, supp, flow_prev, (float *)out, up, n, h, w, ph, pw, cs);
    else
        spynet_level_input_kernel<bf16><<<(unsigned)blocks, 256, 0, st>>>(ref, supp, flow_prev, (bf16 *)out, up, n, h, w, ph, pw, cs);
    return cudaGetLastError();
}

// y_nchw[n,c,h,w] = a_nhwc[n,h,w,c] (first C of CS channels) + b_nchw[n,c,h,w]      (flow = BasicModule(...) + up, :72-78)
__global__ void __launch_bounds__(256) nhwc_plus_nchw_kernel(const float *__restrict__ a, const float *__restrict__ b,
                                                             float *__restrict__ y, int N, int C, int H, int W, int CS) {
    const long long total = (long long)N * C * H * W;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int xw = (int)(i % W), yh = (int)((i / W) % H), c = (int)((i / ((long long)W * H)) % C), n = (int)(i / ((long long)W * H * C));
        y[i] = a[(((long long)n * H + yh) * W + xw) * CS + c] + (b ? b[i] : 0.f);
    }
}
cudaError_t launch_nhwc_plus_nchw(const float *a, const float *b, float *y, int n, int c, int h, int w, int cs, cudaStream_t st) {
    const long long total = (long long)n * c * h * w;
    if (total == 0) return cudaSuccess;
    long long blocks = (total + 255) / 256;
    if (blocks > (long long)sm_count() * 16) blocks = (long long)sm_count() * 16;
    nhwc_plus_nchw_kernel<<<(unsigned)blocks, 256, 0, st>>>(a, b, y, n, c, h, w, cs);
    return cudaGetLastError();
}

// NCHW image (3 channels) -> channels [co, co+3) of an NHWC tensor with CS channels     (torch.cat([x_i, feat_prop]), :69,81)
template <typename TIN, typename TOUT>
__global__ void __launch_bounds__(256) nchw3_to_nhwc_kernel(const TIN *__restrict__ x, long long x_nstride, TOUT *__restrict__ y, int N,
                                                            int H, int W, int CS, int co) {
    const long long P = (long long)N * H * W;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (long long)gridDim.x * blockDim.x) {
        const long long hw = i % ((long long)H * W);
        const int n = (int)(i / ((long long)H * W));
#pragma unroll
        for (int c = 0; c < 3; ++c) y[i * CS + co + c] = from_f32<TOUT>(to_f32<TIN>(x[n * x_nstride + (long long)c * H * W + hw]));
    }
}
cudaError_t launch_nchw3_to_nhwc(const void *x, int x_dtype, long long x_nstride, void *y, int y_dtype, int n, int h, int w, int cs, int co,
                                 cudaStream_t st) {
    const long long P = (long long)n * h * w;
    if (P == 0) return cudaSuccess;
    long long blocks = (P + 255) / 256;
    if (blocks > (long long)sm_count() * 16) blocks = (long long)sm_count() * 16;
    if (x_dtype == kF32 && y_dtype == kF32)
        nchw3_to_nhwc_kernel<float, float><<<(unsigned)blocks, 256, 0, st>>>((const float *)x, x_nstride, (float *)y, n, h, w, cs, co);
    else if (x_dtype == kF32 && y_dtype == kBF16)
        nchw3_to_nhwc_kernel<float, bf16><<<(unsigned)blocks, 256, 0, st>>>((const float *)x, x_nstride, (bf16 *)y, n, h, w, cs, co);
    else if (x_dtype == kBF16 && y_dtype == kBF16)
        nchw3_to_nhwc_kernel<bf16, bf16><<<(unsigned)blocks, 256, 0, st>>>((const bf16 *)x, x_nstride, (bf16 *)y, n, h, w, cs, co);
    else
        return cudaErrorInvalidValue;
    return cudaGetLastError();
}

// out_nchw[n,c,Y,X] = a_nhwc[n,Y,X,c] + bilinear_x4(img)[n,c,Y,X]   (conv_last + F.interpolate(x_i, scale 4, align False), :90-92)
// PS: a is the LR tensor (n,h,w,CS) of 3 x 16 channels and PixelShuffle(4) is applied on the fly: a_hr[n,c,Y,X] = a[n,Y/4,X/4,16c + 4(Y%4) + X%4]
// (models/naive_multi_model_easy.py:141-144: shuf(decode(x)) + base)
template <typename TA, typename TIMG, bool PS>
__global__ void __launch_bounds__(256) vsr_base_add_kernel(const TA *__restrict__ a, int CS, const TIMG *__restrict__ img, long long img_nstride,
                                                           float *__restrict__ y, long long y_nstride, int N, int h, int w) {
    const int OH = 4 * h, OW = 4 * w;
    const long long total = (long long)N * 3 * OH * OW;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ox = (int)(i % OW), oy = (int)((i / OW) % OH), c = (int)((i / ((long long)OW * OH)) % 3), n = (int)(i / ((long long)OW * OH * 3));
        int y0, y1, x0, x1;
        float ly, lx;
        src_index(0.25f, oy, h, false, y0, y1, ly);
        src_index(0.25f, ox, w, false, x0, x1, lx);
        const TIMG *p = img + n * img_nstride + (long long)c * h * w;
        const float base = (1.f - ly) * ((1.f - lx) * to_f32<TIMG>(p[y0 * w + x0]) + lx * to_f32<TIMG>(p[y0 * w + x1])) +
                           ly * ((1.f - lx) * to_f32<TIMG>(p[y1 * w + x0]) + lx * to_f32<TIMG>(p[y1 * w + x1]));
        const long long ai = PS ? (((long long)n * h + (oy >> 2)) * w + (ox >> 2)) * CS + 16 * c + 4 * (oy & 3) + (ox & 3)
                                : (((long long)n * OH + oy) * OW + ox) * CS + c;
        y[n * y_nstride + ((long long)c * OH + oy) * OW + ox] = to_f32<TA>(a[ai]) + base;
    }
}
cudaError_t launch_vsr_base_add(const void *a, int a_dtype, int cs, const void *img, int img_dtype, long long img_nstride, float *y,
                                long long y_nstride, int n, int h, int w, cudaStream_t st, bool shuffle4) {
    const long long total = (long long)n * 3 * 16 * h * w;
    if (total == 0) return cudaSuccess;
    long long blocks = (total + 255) / 256;
    if (blocks > (long long)sm_count() * 16) blocks = (long long)sm_count() * 16;
#define B200SR_BA_CASE(TD, ID, TA, TI)                                                                                                      \
    if (a_dtype == TD && img_dtype == ID) {                                                                                                 \
        if (shuffle4) vsr_base_add_kernel<TA, TI, true><<<(unsigned)blocks, 256, 0, st>>>((const TA *)a, cs, (const TI *)img, img_nstride, y, y_nstride, n, h, w); \
        else vsr_base_add_kernel<TA, TI, false><<<(unsigned)blocks, 256, 0, st>>>((const TA *)a, cs, (const TI *)img, img_nstride, y, y_nstride, n, h, w);      \
        return cudaGetLastError();                                                                                                          \
    }
    B200SR_BA_CASE(kF32, kF32, float, float)
    B200SR_BA_CASE(kBF16, kF32, bf16, float)
    B200SR_BA_CASE(kBF16, kBF16, bf16, bf16)
#undef B200SR_BA_CASE
    return cudaErrorInvalidValue;
}

// Tail of the fork's BasicVSR / MotionVectorVSR after conv_last = ConvTranspose2d(2nf, 3, 5, stride 4) evaluated as a 3x3 convolution with
// 48 output channels on the (h+1) x (w+1) zero-extended features (video.py _transposed_s4k5_as_conv3x3):
//   hr[n,c,yy,xx]  = t[n, yy/4, xx/4, 16 c + 4 (yy%4) + xx%4]           (PixelShuffle(4), cropped to (4h+1) x (4w+1))
//   y[n,c,Y,X]     = bilinear(hr -> (OH,OW))[Y,X] + bilinear(x_i -> (OH,OW))[Y,X]      models/mvvsr_arch.py:100-104, models/basicvsr_arch.py:98-102
// (both F.interpolate calls use align_corners=False) -- shuffle, crop, both resizes and the add in one pass, no intermediate tensor.
template <typename TT, typename TIMG>
__global__ void __launch_bounds__(256) deconv_tail_resize_add_kernel(const TT *__restrict__ t, int CS, const TIMG *__restrict__ img, long long img_nstride,
                                                                     float *__restrict__ y, long long y_nstride, int N, int h, int w, int OH, int OW) {
    const int HH = 4 * h + 1, WH = 4 * w + 1, TWp = w + 1;
    const float sh = resize_scale(HH, OH, false), sw = resize_scale(WH, OW, false), bh = resize_scale(h, OH, false), bw = resize_scale(w, OW, false);
    const long long total = (long long)N * 3 * OH * OW;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ox = (int)(i % OW), oy = (int)((i / OW) % OH), c = (int)((i / ((long long)OW * OH)) % 3), n = (int)(i / ((long long)OW * OH * 3));
        int y0, y1, x0, x1;
        float ly, lx;
        src_index(sh, oy, HH, false, y0, y1, ly);
        src_index(sw, ox, WH, false, x0, x1, lx);
        const TT *tn = t + (long long)n * (h + 1) * TWp * CS + 16 * c;
        auto hr = [&](int yy, int xx) { return to_f32<TT>(tn[((long long)(yy >> 2) * TWp + (xx >> 2)) * CS + 4 * (yy & 3) + (xx & 3)]); };
        const float res = (1.f - ly) * ((1.f - lx) * hr(y0, x0) + lx * hr(y0, x1)) + ly * ((1.f - lx) * hr(y1, x0) + lx * hr(y1, x1));
        src_index(bh, oy, h, false, y0, y1, ly);
        src_index(bw, ox, w, false, x0, x1, lx);
        const TIMG *p = img + n * img_nstride + (long long)c * h * w;
        const float base = (1.f - ly) * ((1.f - lx) * to_f32<TIMG>(p[y0 * w + x0]) + lx * to_f32<TIMG>(p[y0 * w + x1])) +
                           ly * ((1.f - lx) * to_f32<TIMG>(p[y1 * w + x0]) + lx * to_f32<TIMG>(p[y1 * w + x1]));
        y[n * y_nstride + ((long long)c * OH + oy) * OW + ox] = res + base;
    }
}
cudaError_t launch_deconv_tail_resize_add(const void *t, int t_dtype, int cs, const void *img, int img_dtype, long long img_nstride, float *y,
                                          long long y_nstride, int n, int h, int w, int oh, int ow, cudaStream_t st) {
    const long long total = (long long)n * 3 * oh * ow;
    if (total == 0) return cudaSuccess;
    long long blocks = (total + 255) / 256;
    if (blocks > (long long)sm_count() * 16) blocks = (long long)sm_count() * 16;
#define B200SR_DT_CASE(TD, ID, TT, TI) \
    if (t_dtype == TD && img_dtype == ID) { deconv_tail_resize_add_kernel<TT, TI><<<(unsigned)blocks, 256, 0, st>>>((const TT *)t, cs, (const TI *)img, img_nstride, y, y_nstride, n, h, w, oh, ow); return cudaGetLastError(); }
    B200SR_DT_CASE(kF32, kF32, float, float)
    B200SR_DT_CASE(kF32, kBF16, float, bf16)
    B200SR_DT_CASE(kBF16, kF32, bf16, float)
    B200SR_DT_CASE(kBF16, kBF16, bf16, bf16)
#undef B200SR_DT_CASE
    return cudaErrorInvalidValue;
}

// ---- 8-bit frame glue (utils/estimate.py:23-133, common/metrics.py:10-19, datasets/_isr.py:74-75) ---------------------------
// y = x / 255: torchvision's to_tensor on an 8-bit frame, on the device (the H2D copy carries a quarter of the float32 bytes)
template <typename T>
__global__ void __launch_bounds__(256) u8_to_unit_kernel(const uint8_t *__restrict__ x, T *__restrict__ y, long long total) {
    for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < total; i += (long long)gridDim.x * blockDim.x * 4) {
        if (i + 3 < total && (reinterpret_cast<uintptr_t>(x + i) & 3) == 0) {
            const uchar4 v = *reinterpret_cast<const uchar4 *>(x + i);
            y[i] = from_f32<T>((float)v.x / 255.f), y[i + 1] = from_f32<T>((float)v.y / 255.f);
            y[i + 2] = from_f32<T>((float)v.z / 255.f), y[i + 3] = from_f32<T>((float)v.w / 255.f);
        } else {
            for (long long j = i; j < total && j < i + 4; ++j) y[j] = from_f32<T>((float)x[j] / 255.f);
        }
    }
}
cudaError_t launch_u8_to_unit(const uint8_t *x, void *y, int y_dtype, long long total, cudaStream_t st) {
    if (total == 0) return cudaSuccess;
    long long blocks = (total / 4 + 255) / 256 + 1;
    if (blocks > (long long)sm_count() * 16) blocks = (long long)sm_count() * 16;
    if (y_dtype == kF32) u8_to_unit_kernel<float><<<(unsigned)blocks, 256, 0, st>>>(x, (float *)y, total);
    else if (y_dtype == kBF16) u8_to_unit_kernel<bf16><<<(unsigned)blocks, 256, 0, st>>>(x, (bf16 *)y, total);
    else return cudaErrorInvalidValue;
    return cudaGetLastError();
}

// Sum of squared differences of two 8-bit frames per image over the shaved window: the integer core of common/metrics.py:10-19
// (psnr = -10 log10( sum / (255^2 * count) ) per image; both sides are already the quantised frames).  Exact: uint64 accumulation.
__global__ void __launch_bounds__(256) ssd_u8_kernel(const uint8_t *__restrict__ a, const uint8_t *__restrict__ b, unsigned long long *__restrict__ out,
                                                     int C, int H, int W, int shave, int blocks_per_image) {
    const int n = blockIdx.x / blocks_per_image, blk = blockIdx.x % blocks_per_image;
    const int hh = H - 2 * shave, ww = W - 2 * shave;
    const long long count = (long long)C * hh * ww;
    unsigned long long acc = 0;
    for (long long i = (long long)blk * blockDim.x + threadIdx.x; i < count; i += (long long)blocks_per_image * blockDim.x) {
        const int x = (int)(i % ww) + shave, y = (int)((i / ww) % hh) + shave, c = (int)(i / ((long long)ww * hh));
        const long long o = (((long long)n * C + c) * H + y) * W + x;
        const int d = (int)a[o] - (int)b[o];
        acc += (unsigned long long)(d * d);
    }
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, off);
    __shared__ unsigned long long part[8];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int i = 0; i < 8; ++i) t += part[i];
        atomicAdd(out + n, t);
    }
}
cudaError_t launch_ssd_u8(const uint8_t *a, const uint8_t *b, unsigned long long *out, int n, int c, int h, int w, int shave, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(out, 0, sizeof(unsigned long long) * n, st);
    if (e != cudaSuccess) return e;
    const long long count = (long long)c * (h - 2 * shave) * (w - 2 * shave);
    int bpi = (int)((count + 256 * 16 - 1) / (256 * 16));
    bpi = bpi < 1 ? 1 : bpi > 1024 ? 1024 : bpi;
    ssd_u8_kernel<<<n * bpi, 256, 0, st>>>(a, b, out, c, h, w, shave, bpi);
    return cudaGetLastError();
}

// Trunk layout conversion between the WDSR head / tail kernels' trunk (NHWC or planar-8, cp channels) and the NCHW tensors of the
// fork's Split_Block kernel (c channels): one thread moves the 8 channels of one pixel and chunk.  layout: 0 = NHWC [n][h][w][cp],
// 1 = planar-8 [n][cp/8][h][w][8], 2 = NCHW [n][c][h][w] (channels >= c of the padded side read as / are written as zero).
template <typename T>
__global__ void __launch_bounds__(256) trunk_convert_kernel(const T *__restrict__ src, int sl, T *__restrict__ dst, int dl, int N, int C, int CP, int H, int W) {
    const int nch = CP / 8;
    const long long total = (long long)N * H * W * nch, plane = (long long)H * W;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long px = i % plane;                 // consecutive threads = consecutive pixels (coalesced on the planar / NCHW side)
        const int ch = (int)((i / plane) % nch);
        const long long n = i / (plane * nch);
        T v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = 8 * ch + j;
            if (sl == 2) v[j] = c < C ? src[(n * C + c) * plane + px] : from_f32<T>(0.f);
            else if (sl == 1) v[j] = src[((n * nch + ch) * plane + px) * 8 + j];
            else v[j] = src[(n * plane + px) * CP + c];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = 8 * ch + j;
            if (dl == 2) { if (c < C) dst[(n * C + c) * plane + px] = v[j]; }
            else if (dl == 1) dst[((n * nch + ch) * plane + px) * 8 + j] = v[j];
            else dst[(n * plane + px) * CP + c] = v[j];
        }
    }
}
// bf16, planar-8 <-> NCHW with every channel real (the fork NAS_MODEL's two layout changes per forward): a thread moves an 8-channel x
// 8-pixel block -- eight 16-byte loads (one pixel's 8 channels each, or one channel's 8 pixels), an 8 x 8 transpose of 16-bit values with
// byte permutes, eight 16-byte stores.  (The generic kernel below moves 2-byte elements: 94 us for 8 x 360p frames, 31 % of the HBM rate.)
template <bool TO_NCHW>
__global__ void __launch_bounds__(256) trunk_convert8_bf16_kernel(const uint4 *__restrict__ src, uint4 *__restrict__ dst, long long plane8, int nplanes) {
    // plane8 = H * W / 8 pixel groups per channel plane; nplanes = N * C / 8 planar-8 planes; item = (planar plane, pixel group)
    const long long total = plane8 * nplanes;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long g = i % plane8, pl = i / plane8;
        // planar-8 side: 8 consecutive pixels x 16 bytes; NCHW side: channel 8 pl + c at ((8 pl + c) * plane8 + g) x 16 bytes
        const uint4 *sp = TO_NCHW ? src + (pl * plane8 + g) * 8 : src + pl * 8 * plane8 + g;
        const long long ss = TO_NCHW ? 1 : plane8;
        uint32_t a[8][4];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const uint4 v = __ldg(sp + r * ss);
            a[r][0] = v.x, a[r][1] = v.y, a[r][2] = v.z, a[r][3] = v.w;
        }
        uint4 *dp = TO_NCHW ? dst + pl * 8 * plane8 + g : dst + (pl * plane8 + g) * 8;
        const long long ds = TO_NCHW ? plane8 : 1;
#pragma unroll
        for (int c = 0; c < 8; ++c) {       // output row c holds element c of every input row: 16-bit transpose
            uint32_t o[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) o[k] = __byte_perm(a[2 * k][c >> 1], a[2 * k + 1][c >> 1], (c & 1) ? 0x7632 : 0x5410);
            dp[c * ds] = make_uint4(o[0], o[1], o[2], o[3]);
        }
    }
}

cudaError_t launch_trunk_convert(const void *src, int src_layout, void *dst, int dst_layout, int dtype, int n, int c, int cp, int h, int w, cudaStream_t st) {
    if (cp % 8 != 0 || c > cp || src_layout == dst_layout) return cudaErrorInvalidValue;
    const long long total = (long long)n * h * w * (cp / 8);
    if (total == 0) return cudaSuccess;
    if (dtype == kBF16 && c == cp && ((long long)h * w) % 8 == 0 && ((src_layout == 1 && dst_layout == 2) || (src_layout == 2 && dst_layout == 1)) &&
        (reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
        const long long plane8 = (long long)h * w / 8, items = plane8 * n * (cp / 8);
        long long blocks = (items + 255) / 256;
        if (blocks > (long long)sm_count() * 16) blocks = (long long)sm_count() * 16;
        if (dst_layout == 2) trunk_convert8_bf16_kernel<true><<<(unsigned)blocks, 256, 0, st>>>((const uint4 *)src, (uint4 *)dst, plane8, n * (cp / 8));
        else trunk_convert8_bf16_kernel<false><<<(unsigned)blocks, 256, 0, st>>>((const uint4 *)src, (uint4 *)dst, plane8, n * (cp / 8));
        return cudaGetLastError();
    }
    long long blocks = (total + 255) / 256;
    if (blocks > (long long)sm_count() * 16) blocks = (long long)sm_count() * 16;
    if (dtype == kF32) trunk_convert_kernel<float><<<(unsigned)blocks, 256, 0, st>>>((const float *)src, src_layout, (float *)dst, dst_layout, n, c, cp, h, w);
    else if (dtype == kBF16) trunk_convert_kernel<bf16><<<(unsigned)blocks, 256, 0, st>>>((const bf16 *)src, src_layout, (bf16 *)dst, dst_layout, n, c, cp, h, w);
    else return cudaErrorInvalidValue;
    return cudaGetLastError();
}

}  // namespace b200sr
