"""Timing of the 3x3 64->64 bf16 NHWC convolution (BasicVSR trunk shape): tcgen05 kernel vs the mma.sync kernel it replaces.
A "chain" is 60 dependent launches ping-ponging two buffers (one trunk pass of a frame), as the recurrence issues them."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn
from mobilesuperresolution_b200 import video
torch.set_grad_enabled(False)
dev = torch.device("cuda")
conv = nn.Conv2d(64, 64, 3, 1, 1)
hd = video._ConvHandle(conv, dev)


def timeit(fn, reps=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3  # us


for (n, h, w) in [(1, 180, 320), (2, 180, 320), (8, 180, 320), (1, 720, 1280)]:
    x = torch.randn(n, h, w, 64, device=dev).bfloat16()
    y = torch.empty_like(x); z = torch.empty_like(x)
    flop = n * h * w * 64 * 64 * 9 * 2
    for impl in ("tc5-planar8", "tc5", "mma"):
        os.environ["B200SR_CONV_IMPL"] = impl[:3]
        pl = impl.endswith("planar8")
        if pl:
            x, y, z = (t.view(n, 8, h, w, 8) for t in (x, y, z))   # same bytes, read as planar-8
        else:
            x, y, z = (t.view(n, h, w, 64) for t in (x, y, z))
        def chain():
            a, b = x, y
            for k in range(30):
                hd(a, "bf16", video.ACT_RELU, out=z, x_planar=pl, y_planar=pl)
                hd(z, "bf16", video.ACT_NONE, out=b, residual=a, x_planar=pl, y_planar=pl)
                a, b = b, (y if b is x else x)
        us = timeit(chain, reps=5, warm=2) / 60
        g = torch.cuda.CUDAGraph()
        st = torch.cuda.Stream(); st.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(st):
            chain(); st.synchronize()
            with torch.cuda.graph(g, stream=st):
                chain()
            g.replay(); st.synchronize()
            a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a_.record(st)
            for _ in range(5): g.replay()
            b_.record(st); st.synchronize()
        usg = a_.elapsed_time(b_) / 5 * 1e3 / 60
        print(f"{impl} {n}x{h}x{w}: eager {us:8.2f} us/conv = {flop / us / 1e6:7.1f} TFLOP/s ; graph {usg:8.2f} us/conv = {flop / usg / 1e6:7.1f} TFLOP/s", flush=True)
    x.normal_()


print("SPyNet 7x7 layers, 28 pairs at the finest cfg4 level (192x320), bf16 NHWC")
for (cin, cout) in [(8, 32), (32, 64), (64, 32), (32, 16)]:
    hd7 = video._ConvHandle(nn.Conv2d(cin, cout, 7, 1, 3), dev)
    x = torch.randn(28, 192, 320, max(cin, 16), device=dev).bfloat16()
    flop = 28 * 192 * 320 * cin * cout * 49 * 2
    for impl in ("tc5", "mma"):
        os.environ["B200SR_CONV_IMPL"] = impl
        us = timeit(lambda: hd7(x, "bf16", video.ACT_RELU), reps=5, warm=2)
        print(f"{impl} {cin}->{cout}: {us:9.1f} us = {flop / us / 1e6:7.1f} TFLOP/s", flush=True)
os.environ.pop("B200SR_CONV_IMPL", None)
