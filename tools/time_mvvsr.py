"""MotionVectorVSR(64,15) bf16, clip 15 x 180 x 320 -> 720 x 1280: eager (host time / total) and as one CUDA graph (sr.Graphed)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mobilesuperresolution_b200 import video
import mobilesuperresolution_b200 as sr
torch.set_grad_enabled(False)
dev = torch.device("cuda")
m = video.MotionVectorVSR(64, 15).to(dev).eval().set_precision("bf16")
xm = torch.rand(1, 15, 5, 180, 320, device=dev)
xm[:, :, 3:] = (xm[:, :, 3:] - 0.5) * 8
for _ in range(3): m(xm, 720, 1280)
torch.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter(); m(xm, 720, 1280); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"eager: host {1e3*(t1-t0):.2f} ms, total {1e3*(t2-t0):.2f} ms")
g = sr.Graphed(m, xm, 720, 1280)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5): g(xm)
b.record(); torch.cuda.synchronize()
print(f"graph: {a.elapsed_time(b)/5:.2f} ms/clip")
