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
dt = torch.float32 if os.environ.get("WARP_FP32") else torch.bfloat16
for (n, c, h, w) in [(14, 64, 180, 320), (8, 64, 720, 1280), (8, 32, 720, 1280), (4, 128, 360, 640)]:
    x = torch.randn(n, h, w, c, device=dev, dtype=dt)
    fl = ((torch.rand(n, 2, h, w, device=dev) - 0.5) * 6).float()
    gb = n * h * w * (2 * c * x.element_size() + 8) / 1e9
    us = timeit(lambda: video.flow_warp_nhwc(x, fl))
    print(f"NHWC {str(dt)[6:]} {n}x{h}x{w}x{c}: {us:9.1f} us  {gb / us * 1e6:8.1f} GB/s", flush=True)
