"""The tcgen05 3x3 64->64 convolution launches profiled in profiles/r01_conv_tc5_ncu.md: conv_hr shape 1 x 720 x 1280 (NHWC and planar-8)
and one BasicVSR trunk frame 1 x 180 x 320 (planar-8), bf16, LeakyReLU / residual epilogues."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn
from mobilesuperresolution_b200 import video
torch.set_grad_enabled(False)
dev = torch.device("cuda")
hd = video._ConvHandle(nn.Conv2d(64, 64, 3, 1, 1), dev)
for (n, h, w) in [(1, 720, 1280), (1, 180, 320)]:
    x = torch.randn(n, h, w, 64, device=dev).bfloat16()
    r = torch.randn(n, h, w, 64, device=dev).bfloat16()
    xp, rp = x.view(n, 8, h, w, 8), r.view(n, 8, h, w, 8)
    for _ in range(3):
        hd(x, "bf16", video.ACT_LRELU)                                   # NHWC -> NHWC (conv_hr)
        hd(xp, "bf16", video.ACT_NONE, residual=rp, x_planar=True, y_planar=True)   # planar-8 trunk conv2 (+ residual)
torch.cuda.synchronize()
# SPyNet's two big 7x7 layers at cfg4's finest level, both directions batched (28 x 192 x 320), planar-8
h32 = video._ConvHandle(nn.Conv2d(32, 64, 7, 1, 3), dev)
h64 = video._ConvHandle(nn.Conv2d(64, 32, 7, 1, 3), dev)
x32 = torch.randn(28, 4, 192, 320, 8, device=dev).bfloat16()
for _ in range(3):
    t = h32(x32, "bf16", video.ACT_RELU, x_planar=True, y_planar=True)
    h64(t, "bf16", video.ACT_RELU, x_planar=True, y_planar=True)
torch.cuda.synchronize()
# one trunk frame as the propagation launches it: planar-8, + residual, 74-CTA grid (two directions share the GPU), 30x4-output tiles
hd.set_max_ctas(74)
x = torch.randn(1, 8, 180, 320, 8, device=dev).bfloat16(); r = torch.randn_like(x.float()).bfloat16()
for _ in range(3):
    hd(x, "bf16", video.ACT_NONE, residual=r, x_planar=True, y_planar=True)
torch.cuda.synchronize()
