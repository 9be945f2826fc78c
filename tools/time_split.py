"""Split_Block (fork NAS block body) fused kernel at the north-star frame size: 1 x 24 x 360 x 640 (and batch 8), fp32 / bf16 storage."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mobilesuperresolution_b200 as sr
torch.set_grad_enabled(False)
m = sr.Split_Block(num_residual_units=24, kernel_size=3).eval().cuda()
for n in (1, 8):
    for dt in (torch.float32, torch.bfloat16):
        x = torch.randn(n, 24, 360, 640, device="cuda", dtype=dt)
        for _ in range(3): m(x)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20): m(x)
        b.record(); torch.cuda.synchronize()
        us = a.elapsed_time(b) / 20 * 1e3
        px = n * 360 * 640
        print(f"Split_Block {n}x24x360x640 {str(dt)[6:]:8s} {us:8.1f} us  {px * 3720 * 2 / us / 1e6:6.1f} TFLOP/s (3,720 MAC/px)  "
              f"{px * 24 * 2 * x.element_size() / us / 1e3:7.1f} GB/s (read + write once)")
