"""Device-resident timing of the video path (cfg4 shapes): flow_warp (HBM-bound), SPyNet pair batch, BasicVSR_origin clip."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mobilesuperresolution_b200 as sr
from mobilesuperresolution_b200 import video
torch.set_grad_enabled(False)


def timeit(fn, reps=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3  # us


dev = torch.device("cuda")
print("flow_warp (reference signature: x NCHW, flow (n,h,w,2)); algorithmic bytes = read C + write C + 8 B flow per pixel")
for (n, c, h, w, dt) in [(1, 64, 180, 320, torch.float32), (14, 64, 180, 320, torch.float32), (14, 64, 180, 320, torch.bfloat16), (8, 64, 720, 1280, torch.bfloat16)]:
    x = torch.randn(n, c, h, w, device=dev, dtype=dt)
    fl = ((torch.rand(n, h, w, 2, device=dev) - 0.5) * 6).float()
    e = x.element_size()
    gb = n * h * w * (2 * c * e + 8) / 1e9
    if dt == torch.float32:
        us = timeit(lambda: video.flow_warp(x, fl))
        print(f"  NCHW {n}x{c}x{h}x{w} {str(dt)[6:]:9s} {us:9.1f} us  {gb / us * 1e6:8.1f} GB/s")
    xn = x.permute(0, 2, 3, 1).contiguous()
    fn = fl.permute(0, 3, 1, 2).contiguous()
    us = timeit(lambda: video.flow_warp_nhwc(xn, fn))
    print(f"  NHWC {n}x{h}x{w}x{c} {str(dt)[6:]:9s} {us:9.1f} us  {gb / us * 1e6:8.1f} GB/s")

for prec in ("fp32", "bf16"):
    sp = video.SpyNet(None).to(dev).eval().set_precision(prec)
    ref = torch.rand(14, 3, 180, 320, device=dev); sup = torch.rand(14, 3, 180, 320, device=dev)
    us = timeit(lambda: sp(ref, sup), reps=5, warm=2)
    print(f"SPyNet {prec}: 14 pairs 180x320: {us:10.1f} us  = {14 * 39.3e9 / us / 1e6:8.1f} TFLOP/s (39.3 GFLOP/pair)")

for prec in ("fp32", "bf16"):
    m = video.BasicVSR_origin(64, 30).to(dev).eval().set_precision(prec)
    x = torch.rand(1, 15, 3, 180, 320, device=dev)
    us = timeit(lambda: m(x, 720, 1280), reps=3, warm=1)
    print(f"BasicVSR_origin(64,30) {prec}: clip 15x180x320 -> 720x1280: {us / 1e3:9.2f} ms/clip = {15 / us * 1e6:8.1f} frames/s, {11.23e12 / us / 1e6:7.1f} TFLOP/s (11.23 TFLOP/clip)")
    try:   # the same forward as ONE CUDA graph (no Python / ctypes launch overhead between the ~4,000 kernels of a clip)
        st = torch.cuda.Stream()
        st.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(st):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=st):
                y = m(x, 720, 1280)
            g.replay(); st.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st)
            for _ in range(3): g.replay()
            b.record(st); st.synchronize()
        us = a.elapsed_time(b) / 3 * 1e3
        print(f"    as one CUDA graph: {us / 1e3:9.2f} ms/clip = {15 / us * 1e6:8.1f} frames/s, {11.23e12 / us / 1e6:7.1f} TFLOP/s")
    except Exception as e:
        print("    CUDA-graph capture of the clip forward failed:", repr(e)[:200])

# MotionVectorVSR (models/mvvsr_arch.py): flows are codec motion vectors in input channels 3:5 -- no SPyNet, ConvTranspose2d tail
m = video.MotionVectorVSR(64, 15).to(dev).eval().set_precision("bf16")
xm = torch.rand(1, 15, 5, 180, 320, device=dev)
xm[:, :, 3:] = (xm[:, :, 3:] - 0.5) * 8
us = timeit(lambda: m(xm, 720, 1280), reps=5, warm=3)
print(f"MotionVectorVSR(64,15) bf16: clip 15x180x320 -> 720x1280: {us / 1e3:9.2f} ms/clip = {15 / us * 1e6:8.1f} frames/s")
