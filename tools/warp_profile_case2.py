"""The bf16 NHWC flow_warp launches profiled in profiles/r02_flow_warp_ncu.md: cfg4 trunk shape (14 x 180 x 320 x 64) and the 720p shape
(8 x 720 x 1280 x 64)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mobilesuperresolution_b200 import video
dev = torch.device("cuda")
for (n, c, h, w) in [(14, 64, 180, 320), (8, 64, 720, 1280)]:
    x = torch.randn(n, h, w, c, device=dev).bfloat16()
    fl = ((torch.rand(n, 2, h, w, device=dev) - 0.5) * 6).float()
    for _ in range(3):
        video.flow_warp_nhwc(x, fl)
    torch.cuda.synchronize()
