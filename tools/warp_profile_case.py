"""The three flow_warp launches profiled in profiles/r01_flow_warp_ncu.md (cfg4 trunk shape: 14 x 64 x 180 x 320)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mobilesuperresolution_b200 import video
dev = torch.device("cuda")
n, c, h, w = 14, 64, 180, 320
x = torch.randn(n, c, h, w, device=dev)
fl = ((torch.rand(n, h, w, 2, device=dev) - 0.5) * 6).float()
xn = x.permute(0, 2, 3, 1).contiguous(); fn = fl.permute(0, 3, 1, 2).contiguous()
xb = xn.bfloat16()
for _ in range(3):
    video.flow_warp(x, fl); video.flow_warp_nhwc(xn, fn); video.flow_warp_nhwc(xb, fn)
torch.cuda.synchronize()
