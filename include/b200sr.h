/*
 * b200sr.h -- C ABI of the B200-native super-resolution forward path.
 *
 * The reference (zhuzhui-2000/mobilesuperresolution) is pure Python/PyTorch and defines NO FFI:
 * its boundary for this path is the nn.Module surface.  These entry points are what
 * a binding for that surface calls; each cites the reference interface it stands behind
 * (paths relative to /root/reference).  The Python mirror of the nn.Module surface lives in
 * mobilesuperresolution_b200/ and reaches this library through ctypes.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no torch / C++ types.
 *   - Pointers named *_dev are CUDA device pointers, *_host are host pointers.
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).  Device entry points
 *     never synchronise and never allocate: the caller passes a workspace.  They are CUDA-graph
 *     capturable.
 *   - Return value: 0 = ok, <0 = invalid argument (B200SR_E_*), >0 = cudaError_t.
 *     b200sr_last_error() returns a thread-local message for the last non-zero return.
 *   - There is NO CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef B200SR_H_
#define B200SR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200SR_VERSION 1

#if defined(__GNUC__)
#define B200SR_API __attribute__((visibility("default")))
#else
#define B200SR_API
#endif

/* element types of activation tensors, and arithmetic precisions */
#define B200SR_F32 0  /* float32 storage; true-fp32 FMA arithmetic                                  */
#define B200SR_BF16 1 /* bfloat16 storage; bf16 tensor-core operands, fp32 accumulate/bias/residual */
#define B200SR_U8 2   /* y_dtype of b200sr_wdsr_forward* / b200sr_wdsr_tail only (bf16 precision, tcgen05 tail): the 8-bit frame
                       * (sr * 255).round().clamp(0, 255) of common/metrics.py:12 written by the tail epilogue -- a quarter of the
                       * float32 output bytes over PCIe */

/* flow_warp padding modes (models/spynet_arch.py:98 `padding_mode`) */
#define B200SR_PAD_ZEROS 0
#define B200SR_PAD_BORDER 1

/* activations for b200sr_conv2d_nhwc */
#define B200SR_ACT_NONE 0
#define B200SR_ACT_RELU 1
#define B200SR_ACT_LRELU01 2 /* LeakyReLU(0.1), models/basicvsr_arch_origin.py:40 */

#define B200SR_E_INVAL (-1)
#define B200SR_E_STATE (-2)
#define B200SR_E_WORKSPACE (-3)
#define B200SR_E_UNSUPPORTED (-4)

B200SR_API int b200sr_version(void);
B200SR_API const char *b200sr_last_error(void);
/* number of CUDA devices visible; <=0 means the library cannot compute */
B200SR_API int b200sr_device_count(void);

/* ------------------------------------------------------------------------------------------------
 * WDSR-B image path.
 *   BASIC_MODEL.__init__/forward   models/basic_wdsr_b.py:16-93
 *   Block                          models/basic_wdsr_b.py:96-144 == models/wdsr_b.py:253-319
 *   pruned Model                   export_onnx.py:6-88  (per-block (IN,M1,M2), no "+ image_mean")
 *   BinaryConv2d masks / depth gate are resolved by the host into (IN,M1,M2) and kept-block lists
 *   before a plan is created (models/ops.py:7-43, models/wdsr_b.py:358-365).
 * ---------------------------------------------------------------------------------------------- */
typedef struct b200sr_wdsr b200sr_wdsr_t;

typedef struct {
    int32_t scale;       /* PixelShuffle factor s (2, 3 or 4); output channels = 3*s*s          */
    int32_t num_blocks;  /* kept residual blocks (>= 0)                                         */
    int32_t c_trunk;     /* trunk width IN (num_residual_units; 24 dense, 8..24 pruned)         */
    int32_t add_mean;    /* 1: "+ image_mean" after the shuffle (BASIC_MODEL/NAS_MODEL), 0: export_onnx.Model */
    float image_mean;    /* params.image_mean (0.5)                                             */
    const int32_t *m1;   /* [num_blocks] expand widths  (144 dense)                             */
    const int32_t *m2;   /* [num_blocks] reduce widths  (20 dense)                              */
} b200sr_wdsr_desc;

B200SR_API int b200sr_wdsr_create(const b200sr_wdsr_desc *desc, b200sr_wdsr_t **out);
B200SR_API void b200sr_wdsr_destroy(b200sr_wdsr_t *plan);

/* Weight-norm-FOLDED float32 filters in PyTorch OIHW order, host pointers; copied by the call.
 *   head  w[IN][3][3][3]            b[IN]                  models/basic_wdsr_b.py:32-42
 *   block w1[M1][IN] b1[M1]  w2[M2][M1] b2[M2]  w3[IN][M2][3][3] b3[IN]      :108-138
 *   tail  wt[3ss][IN][3][3] bt[3ss]   skip ws[3ss][3][5][5] bs[3ss]           :55-78        */
B200SR_API int b200sr_wdsr_set_head(b200sr_wdsr_t *plan, const float *w_host, const float *b_host);
B200SR_API int b200sr_wdsr_set_block(b200sr_wdsr_t *plan, int block, const float *w1_host, const float *b1_host,
                          const float *w2_host, const float *b2_host, const float *w3_host, const float *b3_host);
B200SR_API int b200sr_wdsr_set_tail(b200sr_wdsr_t *plan, const float *wt_host, const float *bt_host, const float *ws_host,
                         const float *bs_host);
/* Pack (pad to tile multiples, round to bf16 once) and upload to the current device.  Synchronous. */
B200SR_API int b200sr_wdsr_commit(b200sr_wdsr_t *plan);

/* bytes of device workspace b200sr_wdsr_forward needs for an N x 3 x H x W input */
B200SR_API size_t b200sr_wdsr_workspace_bytes(const b200sr_wdsr_t *plan, int n, int h, int w, int precision);

/* y[N,3,sH,sW] = model(x[N,3,H,W]); x and y are contiguous NCHW (the reference's tensors,
 * models/basic_wdsr_b.py:85-93).  x_dtype / y_dtype: element type of x / y; precision: arithmetic. */
B200SR_API int b200sr_wdsr_forward(const b200sr_wdsr_t *plan, const void *x_dev, int x_dtype, void *y_dev, int y_dtype, int n,
                        int h, int w, int precision, void *workspace_dev, size_t workspace_bytes, void *stream);

/* Same call with HOST buffers (pinned recommended): H2D copy, forward, D2H copy, all enqueued on
 * `stream`; x_stage_dev / y_stage_dev are caller-provided device staging buffers of the tensors' sizes. */
B200SR_API int b200sr_wdsr_forward_host(const b200sr_wdsr_t *plan, const void *x_host, int x_dtype, void *y_host, int y_dtype,
                             int n, int h, int w, int precision, void *x_stage_dev, void *y_stage_dev,
                             void *workspace_dev, size_t workspace_bytes, void *stream);

/* Stage-level entry points (parity tests).  Trunk tensors have b200sr_wdsr_trunk_channels() channels (IN padded up to a
 * multiple of 8, or to 24 where that lets a pruned net take the tcgen05 kernels; padding channels are zero), element type =
 * precision, in the layout b200sr_wdsr_trunk_layout() reports for that precision: NHWC, or -- on the bf16 tcgen05 path --
 * planar-8, [n][channels/8][h][w][8].  The layout is internal to the forward; only these three calls expose it. */
#define B200SR_TRUNK_NHWC 0
#define B200SR_TRUNK_PLANAR8 1
B200SR_API int b200sr_wdsr_trunk_channels(const b200sr_wdsr_t *plan);
B200SR_API int b200sr_wdsr_trunk_layout(const b200sr_wdsr_t *plan, int precision);
B200SR_API int b200sr_wdsr_head(const b200sr_wdsr_t *plan, const void *x_dev, int x_dtype, void *trunk_dev, int n, int h, int w,
                     int precision, void *stream);
B200SR_API int b200sr_wdsr_block(const b200sr_wdsr_t *plan, int block, const void *trunk_in_dev, void *trunk_out_dev, int n,
                      int h, int w, int precision, void *stream);
B200SR_API int b200sr_wdsr_tail(const b200sr_wdsr_t *plan, const void *trunk_dev, const void *x_dev, int x_dtype, void *y_dev,
                     int y_dtype, int n, int h, int w, int precision, void *stream);
/* kernels launched by the last b200sr_wdsr_forward* call on this plan (bench "gpu_launches") */
B200SR_API int b200sr_wdsr_launches_per_forward(const b200sr_wdsr_t *plan);
/* Host-only introspection (no device needed): the shared-memory operand image the row-streaming tcgen05 block kernel consumes
 * for one Block (models/basic_wdsr_b.py:96-144) with folded filters w1 (m1, c), w2 (m2, m1), w3 (c, m2, 3, 3) -- the layout of
 * csrc/wdsr_rs_layout.cuh, including the A-operand slice table of the 3x3.  tests/ replays the kernel's data flow on it.
 * *bytes receives the image size; the image is written when img_host != NULL and cap >= *bytes. */
B200SR_API int b200sr_wdsr_pack_block_image(int c, int m1, int m2, const float *w1, const float *b1, const float *w2, const float *b2,
                                 const float *w3, const float *b3, void *img_host, size_t cap, size_t *bytes);

/* ------------------------------------------------------------------------------------------------
 * Split_Block.forward_body (the fork's searchable block, models/wdsr_b.py:406-496), one fused kernel:
 *   x1 = e*x; x2 = x - x1; x3 = x2 + sum_k p_k relu(PW_k(relu(DW_k(x1)))) + x1; y = x2 + e*x3,   k = 3, 5, 7
 * Host arrays (float32): dw_k = weight-norm-folded depthwise filters [C][k*k] (Conv_sep.body[0], :382), dw_bias [3][C],
 * pw = folded 1x1 filters [3][C out][C in] (Conv_sep.body[2], :387), pw_bias [3][C], mask_eff [C] = the BinaryConv2d forward
 * weight w - (w - rounding(w, 0)) of `split` (:424, models/ops.py:18-26), prob [3] = softmax(alpha) (:487).
 * x / y: (n, C, h, w) NCHW of `dtype` (arithmetic is fp32 FMA for both).  C in {8, 16, 24, 32}.  y must not alias x
 * (B200SR_E_INVAL): a CTA reads a 3-pixel halo of x that neighbouring CTAs would overwrite.
 * ---------------------------------------------------------------------------------------------- */
typedef struct b200sr_split b200sr_split_t;
B200SR_API int b200sr_split_create(int channels, const float *dw3_host, const float *dw5_host, const float *dw7_host,
                                   const float *dw_bias_host, const float *pw_host, const float *pw_bias_host,
                                   const float *mask_eff_host, const float *prob_host, b200sr_split_t **out);
B200SR_API void b200sr_split_destroy(b200sr_split_t *blk);
B200SR_API int b200sr_split_forward(const b200sr_split_t *blk, const void *x_dev, void *y_dev, int n, int h, int w, int dtype,
                                    void *stream);
/* Per-channel multiplier applied to x before everything else (host array [C]; NULL = all ones): NAS_MODEL.forward's
 * `y = self.mask(y)` in front of every block (models/wdsr_b.py:116-119), fused into the block kernel. */
B200SR_API int b200sr_split_set_premask(b200sr_split_t *blk, const float *premask_host);

/* The fork's NAS_MODEL.forward (models/wdsr_b.py:105-137) in one call: head -> the KEPT MyAggregationLayer blocks (depth gates
 * resolved by the host, models/wdsr_b.py:539-546; each with the global mask as its pre-mask) -> tail + skip + PixelShuffle (+ mean).
 * `plan` is a b200sr_wdsr_t created with num_blocks = 0 whose head / tail filters carry the global mask (head: nothing to fold,
 * the first block's pre-mask applies it; tail: input channels scaled by the mask).  x / y as in b200sr_wdsr_forward. */
B200SR_API size_t b200sr_nas_workspace_bytes(const b200sr_wdsr_t *plan, int n, int h, int w, int precision);
B200SR_API int b200sr_nas_forward(const b200sr_wdsr_t *plan, const b200sr_split_t *const *blocks, int num_blocks, const void *x_dev,
                                  int x_dtype, void *y_dev, int y_dtype, int n, int h, int w, int precision, void *workspace_dev,
                                  size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------------------------------
 * flow_warp(x, flow, 'bilinear', padding_mode, align_corners=True)   models/spynet_arch.py:98-129
 * (mmedit twin used at models/basicvsr_arch.py:74,85; basicvsr_arch_origin.py:68,79).
 * x, y: (n,c,h,w) NCHW float32 contiguous.  flow: float32, logical shape (n,h,w,2), addressed with
 * ELEMENT strides so the caller's `.permute(0,2,3,1)` view of an (n,2,h,w) tensor is consumed as is.
 * ---------------------------------------------------------------------------------------------- */
B200SR_API int b200sr_flow_warp_nchw(const float *x_dev, const float *flow_dev, int64_t fs_n, int64_t fs_h, int64_t fs_w,
                          int64_t fs_c, float *y_dev, int n, int c, int h, int w, int padding_mode, void *stream);
/* NHWC variant used inside the video path: x,y (n,h,w,c) of `dtype`, c % 8 == 0 for bf16 / % 4 for f32;
 * flow (n,2,h,w) float32 planar. */
B200SR_API int b200sr_flow_warp_nhwc(const void *x_dev, const float *flow_nchw_dev, void *y_dev, int n, int c, int h, int w,
                          int padding_mode, int dtype, void *stream);

/* The same warp written into channels [y_coff, y_coff + c) of a wider NHWC tensor of y_cstride channels: the warped features land
 * directly inside the trunk's input (torch.cat([x_i, feat_prop], 1), models/basicvsr_arch_origin.py:69,81) -- no copy. */
B200SR_API int b200sr_flow_warp_nhwc_into(const void *x_dev, const float *flow_nchw_dev, void *y_dev, int y_cstride, int y_coff, int n, int c,
                                          int h, int w, int padding_mode, int dtype, void *stream);
/* ... and x read from channels [x_coff, x_coff + c) of a wider tensor (n,h,w,x_cstride) as well */
B200SR_API int b200sr_flow_warp_nhwc_windows(const void *x_dev, int x_cstride, int x_coff, const float *flow_nchw_dev, void *y_dev, int y_cstride,
                                             int y_coff, int n, int c, int h, int w, int padding_mode, int dtype, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Video path building blocks (SPyNet + BasicVSR).  The Python mirror (mobilesuperresolution_b200/video.py)
 * sequences these exactly as the reference's forward does:
 *   SpyNet.process / forward          models/spynet_arch.py:49-96
 *   BasicVSR_origin.forward           models/basicvsr_arch_origin.py:53-96
 * ---------------------------------------------------------------------------------------------- */
typedef struct b200sr_conv b200sr_conv_t;

/* nn.Conv2d(cin, cout, k, stride 1, padding k/2) with k in {1,3,7}; w_host is PyTorch OIHW float32, bias may be NULL.
 * (models/spynet_arch.py:17-22, models/basicvsr_arch_origin.py:31-35,110-131) */
B200SR_API int b200sr_conv_create(int cin, int cout, int k, const float *w_host, const float *bias_host, b200sr_conv_t **out);
B200SR_API void b200sr_conv_destroy(b200sr_conv_t *conv);
/* y = act(conv(x)) (+ residual).  x / y / residual are NHWC tensors of n x h x w pixels whose used channels start at
 * *_coff inside a pixel of *_cstride channels (so concatenations are views).  shuffle = 2 folds PixelShuffle(2) into the
 * store (y is then n x 2h x 2w x cout/4).  act: B200SR_ACT_*.  precision F32: in/out float32; BF16: in bf16|f32, out bf16|f32. */
B200SR_API int b200sr_conv_forward(const b200sr_conv_t *conv, const void *x_dev, int x_cstride, int x_coff, void *y_dev, int y_cstride,
                                   int y_coff, const void *residual_dev, int r_cstride, int r_coff, int n, int h, int w, int act,
                                   int shuffle, int in_dtype, int out_dtype, int precision, void *stream);

/* The same convolution with explicit activation layouts (B200SR_TRUNK_NHWC | B200SR_TRUNK_PLANAR8).  Planar-8,
 * [n][channels/8][h][w][8] bf16 (channel count a multiple of 8, no channel window), is what the BasicVSR propagation trunks
 * (ConvResidualBlocks, models/basicvsr_arch_origin.py:98-137) and SPyNet's BasicModule (models/spynet_arch.py:17-22) keep
 * their private tensors in: only the tcgen05 kernels take it -- 3x3 (64..80) -> 64 k (k <= 4, PixelShuffle(2) store for the
 * upsampler convs), 1x1 128 -> 64 k (NHWC input) and 7x7 (8|16 -> 32, 32 -> 64, 64 -> 32, 32 -> 16) -- in bf16 precision with bf16 tensors; anything else
 * returns B200SR_E_UNSUPPORTED.  The residual is laid out like x.  b200sr_conv_tcgen05_ok() != 0 says such a kernel exists
 * for this conv's shape (developer switch B200SR_CONV_IMPL=mma turns them off). */
B200SR_API int b200sr_conv_forward_layout(const b200sr_conv_t *conv, const void *x_dev, int x_layout, int x_cstride, int x_coff, void *y_dev,
                                          int y_layout, int y_cstride, int y_coff, const void *residual_dev, int r_cstride, int r_coff, int n,
                                          int h, int w, int act, int shuffle, int in_dtype, int out_dtype, int precision, void *stream);
B200SR_API int b200sr_conv_tcgen05_ok(const b200sr_conv_t *conv);
/* Grid cap of this conv's tcgen05 launches (persistent kernels, one CTA per SM by default; 0 restores that).  BasicVSR's two
 * propagation directions (models/basicvsr_arch_origin.py:61-82) are independent chains of one-frame launches: run on two streams
 * with half-size grids they share the SMs instead of queueing behind each other (14.7 -> 12.4 ms per 15-frame clip). */
B200SR_API int b200sr_conv_set_max_ctas(b200sr_conv_t *conv, int max_ctas);

/* ConvResidualBlocks.forward (models/basicvsr_arch_origin.py:98-137) as one host call on the tcgen05 kernels, bf16:
 *   t = lrelu(first(buf));  for k: o = relu(conv1_k(t)); t = t + conv2_k(o);  out = t
 * blocks = {conv1_0, conv2_0, conv1_1, ...} (2 * num_block handles, 3x3 64 -> 64); first: 3x3 (64..80) -> 64 on the NHWC trunk input
 * `buf` (n,h,w,buf_cstride); t, o: planar-8 scratch tensors of n*h*w*64 bf16 each; out: NHWC (n,h,w,64).  2 * num_block + 1 launches
 * on `stream`, no synchronisation.  Fails (B200SR_E_UNSUPPORTED) if a conv is not served by the tcgen05 kernel. */
B200SR_API int b200sr_vsr_trunk_forward(const b200sr_conv_t *first, const b200sr_conv_t *const *blocks, int num_block, const void *buf_dev,
                                        int buf_cstride, void *t_dev, void *o_dev, void *out_dev, int n, int h, int w, void *stream);
/* the same with the NHWC features written into channels [out_coff, out_coff + 64) of a wider tensor (n,h,w,out_cstride): the backward and
 * forward trunks of a frame fill the two halves of ONE tensor, `torch.cat([out_l[i], feat_prop], dim=1)` (models/basicvsr_arch_origin.py:84)
 * without the copy */
B200SR_API int b200sr_vsr_trunk_forward_into(const b200sr_conv_t *first, const b200sr_conv_t *const *blocks, int num_block, const void *buf_dev,
                                             int buf_cstride, void *t_dev, void *o_dev, void *out_dev, int out_cstride, int out_coff, int n, int h,
                                             int w, void *stream);

/* conv_last + base of BasicVSR_origin's reconstruction in ONE kernel (models/basicvsr_arch_origin.py:90-92):
 *   y[n,c,Y,X] = conv3x3(x)[n,c,Y,X] + F.interpolate(base, scale_factor=4, mode='bilinear', align_corners=False)[n,c,Y,X]
 * conv = a 3x3 64 -> 3 conv; x: n x H x W x 64 bf16 (NHWC channel window or planar-8); base: float32 NCHW (3, H/4, W/4) per image,
 * image n at base + n * base_nstride elements; y: float32 NCHW (3, H, W) per image at y + n * y_nstride.  tcgen05 only. */
B200SR_API int b200sr_vsr_conv_last_base(const b200sr_conv_t *conv, const void *x_dev, int x_layout, int x_cstride, int x_coff,
                                         const float *base_dev, int64_t base_nstride, float *y_dev, int64_t y_nstride, int n, int H, int W,
                                         void *stream);

/* SpyNet.forward (models/spynet_arch.py:49-96) in one call: bilinear resize to multiples of 32 + ImageNet normalisation, the 5 average
 * pools, six pyramid levels (x2 upsample of the flow -> warp(border) -> 7x7 convs 8-32-64-32-16-2 -> + flow), final resize + rescale.
 * convs: 30 handles, level-major (basic_module.L.basic_module.{0,2,4,6,8}, L = 0..5).  ref / supp: NCHW (n,3,h,w) of img_dtype;
 * flow_out: float32 NCHW (n,2,h,w).  mean4 / inv_std4: host arrays {r,g,b,0} / {1/r,1/g,1/b,1} (the module's `mean` / `std` buffers).
 * precision B200SR_BF16 runs the convolutions on bf16 tensor-core operands; flows, warps and the pyramid stay float32. */
B200SR_API size_t b200sr_spynet_workspace_bytes(int n, int h, int w, int precision);
B200SR_API int b200sr_spynet_forward(const b200sr_conv_t *const *convs, const void *ref_dev, const void *supp_dev, int img_dtype,
                                     float *flow_out_dev, int n, int h, int w, int precision, const float *mean4_host,
                                     const float *inv_std4_host, void *workspace_dev, size_t workspace_bytes, void *stream);

/* 8-bit frame glue around the forward.
 *   b200sr_u8_to_unit: y = x / 255, torchvision's to_tensor on an 8-bit frame (datasets/_isr.py:74-75), any shape, `count` elements.
 *   b200sr_ssd_u8:     out[i] = sum over (c, shaved h, shaved w) of (a - b)^2 of image i, uint64 -- the integer core of
 *                      common/metrics.py:10-19 for two quantised frames: psnr_i = -10 log10( out[i] / (255^2 * c*(h-2s)*(w-2s)) ). */
B200SR_API int b200sr_u8_to_unit(const uint8_t *x_dev, void *y_dev, int y_dtype, int64_t count, void *stream);
B200SR_API int b200sr_ssd_u8(const uint8_t *a_dev, const uint8_t *b_dev, uint64_t *out_dev, int n, int c, int h, int w, int shave, void *stream);

/* Tail of the fork's BasicVSR / MotionVectorVSR (models/basicvsr_arch.py:96-102, models/mvvsr_arch.py:98-104) behind conv_last =
 * ConvTranspose2d(2nf, 3, 5, stride 4) evaluated as a 3x3 convolution with 3 x 16 output channels on the (h+1) x (w+1) zero-extended
 * features: t_dev NHWC (n, h+1, w+1, t_cstride >= 48).  PixelShuffle(4) + crop to (4h+1) x (4w+1) + bilinear resize to (oh, ow) +
 * bilinear base of the 3-channel NCHW frame img_dev (n, 3, h, w), all align_corners=False, in one pass; y float32 NCHW (n, 3, oh, ow). */
B200SR_API int b200sr_vsr_deconv_tail(const void *t_dev, int t_dtype, int t_cstride, const void *img_dev, int img_dtype, int64_t img_nstride,
                                      float *y_dev, int64_t y_nstride, int n, int h, int w, int oh, int ow, void *stream);

/* F.interpolate(x, size=(oh,ow), mode='bilinear', align_corners) on NCHW, then (v - sub[c%4]) * mul[c%4]; y float32.
 * (models/spynet_arch.py:88-94 pre/post resize, normalisation :45-47; models/basicvsr_arch_origin.py:93) */
B200SR_API int b200sr_resize_bilinear_nchw(const void *x_dev, int x_dtype, float *y_dev, int n, int c, int h, int w, int oh, int ow,
                                           int align_corners, const float *sub4_host, const float *mul4_host, void *stream);
/* F.avg_pool2d(x, 2, 2, count_include_pad=False), NCHW float32   (models/spynet_arch.py:56-57) */
B200SR_API int b200sr_avg_pool2_nchw(const float *x_dev, float *y_dev, int n, int c, int h, int w, void *stream);
/* One SPyNet level's network input, fused (models/spynet_arch.py:64-78): up = 2*interpolate(flow_prev, x2, align_corners=True)
 * (replicate-padded to h x w), warped = flow_warp(supp, up, 'border'); out NHWC [ref|warped|up|0..] with cs channels;
 * up_dev (n,2,h,w) float32.  flow_prev_dev NULL = the coarsest level's zero flow; (ph,pw) = its size. */
B200SR_API int b200sr_spynet_level_input(const float *ref_dev, const float *supp_dev, const float *flow_prev_dev, void *out_dev,
                                         int out_dtype, float *up_dev, int n, int h, int w, int ph, int pw, int cs, void *stream);
/* y[n,c,h,w] = a[n,h,w,c] + b[n,c,h,w]  (a NHWC float32 with cs channels, b may be NULL)   (models/spynet_arch.py:72-78) */
B200SR_API int b200sr_nhwc_plus_nchw(const float *a_dev, const float *b_dev, float *y_dev, int n, int c, int h, int w, int cs, void *stream);
/* zero-fill `bytes` bytes of device memory on `stream` (cudaMemsetAsync: a memset node under graph capture, no kernel): the
 * zero features of the first propagation step and the pad channels of the trunk input
 * (`feat_prop = x.new_zeros(b, self.num_feat, h, w)`, models/basicvsr_arch_origin.py:62,75) */
B200SR_API int b200sr_zero_async(void *dst_dev, size_t bytes, void *stream);
/* dst (n, h+1, w+1, px_bytes) = src (n, h, w, px_bytes) with one zero row below and one zero column to the right
 * (torch.nn.functional.pad(o, (0, 0, 0, 1, 0, 1)) of an NHWC tensor: the domain on which ConvTranspose2d(2 nf, 3, 5, stride=4),
 * models/mvvsr_arch.py:38,100, is evaluated as a 3x3 convolution).  cudaMemcpy2DAsync + two memsets per image: copy / memset
 * nodes under graph capture, no kernel */
B200SR_API int b200sr_pad_bottom_right_async(const void *src_dev, void *dst_dev, int n, int h, int w, int px_bytes, void *stream);
/* copy a 3-channel NCHW image (image n at x + n*x_nstride elements) into channels [co,co+3) of an NHWC tensor
 * (torch.cat([x_i, feat_prop], 1), models/basicvsr_arch_origin.py:69,81) */
B200SR_API int b200sr_nchw3_to_nhwc(const void *x_dev, int x_dtype, int64_t x_nstride, void *y_dev, int y_dtype, int n, int h, int w,
                                    int cs, int co, void *stream);
/* out[n,c,4h,4w] = a[n,4h,4w,c] + F.interpolate(img, scale_factor=4, 'bilinear', align_corners=False)
 * (models/basicvsr_arch_origin.py:90-92); a NHWC with cs channels, out float32 with image stride y_nstride elements */
B200SR_API int b200sr_vsr_base_add(const void *a_dev, int a_dtype, int cs, const void *img_dev, int img_dtype, int64_t img_nstride,
                                   float *y_dev, int64_t y_nstride, int n, int h, int w, void *stream);

/* out[n,c,4h,4w] = PixelShuffle(4)(a)[n,c,4h,4w] + F.interpolate(img, scale_factor=4, 'bilinear', align_corners=False): the tail of
 * Naive_model (models/naive_multi_model_easy.py:141-144, `self.shuf(self.decode(x_)) + base`); a NHWC (n,h,w,cs >= 48), channel 16c + 4i + j
 * of a low-resolution pixel is output (c, 4y + i, 4x + j); out float32 with image stride y_nstride elements */
B200SR_API int b200sr_vsr_shuffle4_base_add(const void *a_dev, int a_dtype, int cs, const void *img_dev, int img_dtype, int64_t img_nstride,
                                            float *y_dev, int64_t y_nstride, int n, int h, int w, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* B200SR_H_ */
