import os, sys
sys.path.insert(0, "/root/repo")
import torch
from mobilesuperresolution_b200 import video
torch.set_grad_enabled(False)
def timeit(fn, reps=30, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
dev = torch.device("cuda")
for (n, c, h, w) in [(14, 64, 180, 320), (4, 64, 720, 1280)]:
    x = torch.randn(n, c, h, w, device=dev)
    fl = ((torch.rand(n, h, w, 2, device=dev) - 0.5) * 6).float()
    gb = n * h * w * (2 * c * 4 + 8) / 1e9
    us = timeit(lambda: video.flow_warp(x, fl))
    print(f"NCHW fp32 {n}x{c}x{h}x{w}: {us:9.1f} us  {gb / us * 1e6:8.1f} GB/s", flush=True)
