// launch.h -- internal launcher interface between the C-ABI (b200sr.cu) and the kernel translation units.
#pragma once
#include <vector>
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200sr {

// precision / dtype codes mirror include/b200sr.h
constexpr int kF32 = 0, kBF16 = 1, kU8 = 2;   // kU8: output of the tcgen05 tail only

int sm_count();  // cached multiprocessor count of the current device

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute: the opt-in is tracked per (kernel instantiation, device),
// so a host thread that drives a second GPU opts in there too.  One static instance per call site (template instantiation).
struct SmemOptIn {
    size_t set[64] = {};
    template <typename K> cudaError_t ensure(K kern, size_t smem) {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64 && set[dev] >= smem) return cudaSuccess;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess && dev >= 0 && dev < 64) set[dev] = smem;
        return e;
    }
};

// tile geometry chosen by the launchers (exposed for tests/bench reporting)
struct TileInfo { int tw, th, tiles, ctas; };

cudaError_t launch_head(int CP, int x_dtype, int trunk_dtype, const void *x, void *trunk, const float *wpack, int N, int H,
                        int W, float mean, cudaStream_t st);
cudaError_t launch_block_f32(int CP, int M2P, const float *in, float *out, const float *wpack, int M1P, int N, int H, int W,
                             cudaStream_t st);
cudaError_t launch_block_bf16(int CP, int M2P, const void *in, void *out, const uint8_t *wimg, int M1P, int N, int H, int W,
                              cudaStream_t st);
// tcgen05 form of the fused block (CP == 24, M2 <= 24): variant 0 = sequential reference form, 1 = pipelined
// true when launch_block_tc5(variant 1) runs the packed 3x3 for this reduce width (M2 = 17..20; B200SR_G3_NOPACK=1 switches it off): the operand
// image must then carry the packed w3
bool block_tc5_g3_packed(int M2);
cudaError_t launch_block_tc5(int variant, const void *in, void *out, const uint8_t *wimg, int M1P, int M2, int N, int H, int W,
                             cudaStream_t st);
// every block of a run of same-shaped blocks in one persistent cooperative launch (wdsr_tc5c.cuh); gsync: 4 bytes of device memory
cudaError_t launch_block_chain_tc5(void *buf_a, void *buf_b, const uint8_t *const *wimgs, int nlayers, unsigned *gsync, int M1P, int M2, int N,
                                   int H, int W, cudaStream_t st);
// row-streaming tcgen05 form of the fused block (wdsr_rs.cuh; planar-8 trunk, CP == 24, M2 <= 24, M1P <= 144)
bool block_rs_eligible(int N, int H, int W);
cudaError_t launch_block_rs(const void *in, void *out, const uint8_t *wimg, int M1P, int M2, int N, int H, int W, cudaStream_t st);
// the same with the reduce 1x1 on mma.sync out of registers (wdsr_rh.cuh; same operand image, same eligibility)
cudaError_t launch_block_rh(const void *in, void *out, const uint8_t *wimg, int M1P, int M2, int N, int H, int W, cudaStream_t st);
// tcgen05 form of the head (bf16 trunk padded to 24 channels)
cudaError_t launch_head_tc5(int x_dtype, const void *x, void *trunk, const uint8_t *wimg, int N, int H, int W, float mean, cudaStream_t st);
// the same head on mma.sync (wdsr_head_mma.cu; the default of the tcgen05 path): wh = the fp32 head image [27][24] | bias[24]
cudaError_t launch_head_mma(int x_dtype, const void *x, void *trunk, const float *wh, int N, int H, int W, float mean, cudaStream_t st);
// tcgen05 form of the fused tail (trunk padded to 24 channels)
cudaError_t launch_tail_tc5(int S, int x_dtype, int y_dtype, const void *trunk, const void *x, void *y, const uint8_t *wimg, int N, int H,
                            int W, float mean, float out_add, cudaStream_t st);
cudaError_t launch_tail_f32(int CP, int S, int x_dtype, int y_dtype, const float *trunk, const void *x, void *y,
                            const float *wpack, int N, int H, int W, float mean, float out_add, cudaStream_t st);
cudaError_t launch_tail_bf16(int CP, int S, int x_dtype, int y_dtype, const void *trunk, const void *x, void *y,
                             const uint8_t *wimg, int N, int H, int W, float mean, float out_add, cudaStream_t st);

cudaError_t launch_flow_warp_nchw(const float *x, const float *flow, long long fs_n, long long fs_h, long long fs_w,
                                  long long fs_c, float *y, int n, int c, int h, int w, int border, cudaStream_t st);
cudaError_t launch_flow_warp_nhwc(const void *x, const float *flow_nchw, void *y, int n, int c, int h, int w, int border,
                                  int dtype, cudaStream_t st, int y_cs = 0, int y_co = 0, int x_cs = 0, int x_co = 0);   // y_cs > 0: y is a channel window

struct ConvArgs;
// generic NHWC convolution (conv.cuh); nt = output-channel n-tiles per CTA of the bf16 kernel (1,2,4,8)
cudaError_t launch_conv(const ConvArgs &a, int k, int nt, int in_dtype, int out_dtype, int precision, cudaStream_t st);
// tcgen05 3x3 64 -> 64 bf16 NHWC convolution (conv_tc5.cuh): wimg = its operand image; eligibility = channel windows 16-byte aligned
bool conv_tc5_eligible(const ConvArgs &a);
cudaError_t launch_conv3x3_c64_tc5(const ConvArgs &a, const uint8_t *wimg, cudaStream_t st);
// tcgen05 7x7 bf16 convolution (conv7_tc5.cuh; SPyNet layers): wimg = [7 tap rows][stage image]
int conv7_tc5_nch(int cin);
bool conv7_tc5_shape_ok(int cin, int cout);
bool conv7_tc5_eligible(const ConvArgs &a);
cudaError_t launch_conv7x7_tc5(const ConvArgs &a, const uint8_t *wimg, cudaStream_t st);
// fused Split_Block body (split_block.cu): params = HOST pointer to the packed float image of split_param_floats(C) floats (it
// travels as a kernel argument), C in {8,16,24,32}
int split_param_floats(int C);
cudaError_t launch_split_block(int C, int dtype, const void *x, void *y, const float *params, int N, int H, int W, cudaStream_t st);
// bf16 arm with the 1x1 convolutions on mma.sync (split_block_tc.cu): image = HOST pointer to the kernel-argument image that
// split_tc_pack derives from the packed float image; eligible: bf16 tensors, W % 8 == 0, 16-byte aligned pointers
void split_tc_pack(int C, const float *packed, std::vector<uint8_t> &out);
bool split_tc_eligible(int dtype, const void *x, const void *y, int W);
cudaError_t launch_split_block_tc(int C, const void *x, void *y, const uint8_t *image, int N, int H, int W, cudaStream_t st);
// trunk layout conversion (video_glue.cu): layouts 0 = NHWC, 1 = planar-8, 2 = NCHW (c of cp channels)
cudaError_t launch_trunk_convert(const void *src, int src_layout, void *dst, int dst_layout, int dtype, int n, int c, int cp, int h, int w, cudaStream_t st);
cudaError_t launch_deconv_tail_resize_add(const void *t, int t_dtype, int cs, const void *img, int img_dtype, long long img_nstride, float *y,
                                          long long y_nstride, int n, int h, int w, int oh, int ow, cudaStream_t st);
// 8-bit frame glue (video_glue.cu)
cudaError_t launch_u8_to_unit(const uint8_t *x, void *y, int y_dtype, long long total, cudaStream_t st);
cudaError_t launch_ssd_u8(const uint8_t *a, const uint8_t *b, unsigned long long *out, int n, int c, int h, int w, int shave, cudaStream_t st);
// video glue (video_glue.cu)
cudaError_t launch_resize_bilinear_nchw(const void *x, int x_dtype, float *y, int n, int c, int h, int w, int oh, int ow, int align,
                                        const float *sub4, const float *mul4, cudaStream_t st);
cudaError_t launch_avg_pool2(const float *x, float *y, int n, int c, int h, int w, cudaStream_t st);
cudaError_t launch_spynet_level_input(const float *ref, const float *supp, const float *flow_prev, void *out, int out_dtype, float *up,
                                      int n, int h, int w, int ph, int pw, int cs, cudaStream_t st);
cudaError_t launch_nhwc_plus_nchw(const float *a, const float *b, float *y, int n, int c, int h, int w, int cs, cudaStream_t st);
cudaError_t launch_nchw3_to_nhwc(const void *x, int x_dtype, long long x_nstride, void *y, int y_dtype, int n, int h, int w, int cs, int co,
                                 cudaStream_t st);
cudaError_t launch_vsr_base_add(const void *a, int a_dtype, int cs, const void *img, int img_dtype, long long img_nstride, float *y,
                                long long y_nstride, int n, int h, int w, cudaStream_t st, bool shuffle4 = false);

}  // namespace b200sr
