// launch.h -- internal launcher interface between the C-ABI (b200sr.cu) and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200sr {

// precision / dtype codes mirror include/b200sr.h
constexpr int kF32 = 0, kBF16 = 1;

int sm_count();  // cached multiprocessor count of the current device

// tile geometry chosen by the launchers (exposed for tests/bench reporting)
struct TileInfo { int tw, th, tiles, ctas; };

cudaError_t launch_head(int CP, int x_dtype, int trunk_dtype, const void *x, void *trunk, const float *wpack, int N, int H,
                        int W, float mean, cudaStream_t st);
cudaError_t launch_block_f32(int CP, int M2P, const float *in, float *out, const float *wpack, int M1P, int N, int H, int W,
                             cudaStream_t st);
cudaError_t launch_block_bf16(int CP, int M2P, const void *in, void *out, const uint8_t *wimg, int M1P, int N, int H, int W,
                              cudaStream_t st);
// tcgen05 form of the fused block (CP == 24, M2 <= 24): variant 0 = sequential reference form, 1 = pipelined
cudaError_t launch_block_tc5(int variant, const void *in, void *out, const uint8_t *wimg, int M1P, int N, int H, int W,
                             cudaStream_t st);
cudaError_t launch_tail_f32(int CP, int S, int x_dtype, int y_dtype, const float *trunk, const void *x, void *y,
                            const float *wpack, int N, int H, int W, float mean, float out_add, cudaStream_t st);
cudaError_t launch_tail_bf16(int CP, int S, int x_dtype, int y_dtype, const void *trunk, const void *x, void *y,
                             const uint8_t *wimg, int N, int H, int W, float mean, float out_add, cudaStream_t st);

cudaError_t launch_flow_warp_nchw(const float *x, const float *flow, long long fs_n, long long fs_h, long long fs_w,
                                  long long fs_c, float *y, int n, int c, int h, int w, int border, cudaStream_t st);
cudaError_t launch_flow_warp_nhwc(const void *x, const float *flow_nchw, void *y, int n, int c, int h, int w, int border,
                                  int dtype, cudaStream_t st);

}  // namespace b200sr
