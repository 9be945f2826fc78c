"""Where a BasicVSR_origin(64,30) bf16 clip (15 x 180 x 320 -> 720 x 1280) spends its time: each part captured as one CUDA graph."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mobilesuperresolution_b200 import video
torch.set_grad_enabled(False)
dev = torch.device("cuda")
m = video.BasicVSR_origin(64, 30).to(dev).eval().set_precision("bf16")
x = torch.rand(1, 15, 3, 180, 320, device=dev)


def graph_time(fn, reps=5):
    st = torch.cuda.Stream(); st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        fn(); st.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            out = fn()
        g.replay(); st.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        for _ in range(reps): g.replay()
        b.record(st); st.synchronize()
    return a.elapsed_time(b) / reps, out


t_flow, (ff, fb) = graph_time(lambda: m.get_flow(x))
t_prop, (back, fwd) = graph_time(lambda: m.propagate(x, ff, fb))
one = os.environ.get("B200SR_ONE_STREAM")
t_all, _ = graph_time(lambda: m(x, 720, 1280))
print(f"get_flow (SPyNet, 28 pairs) {t_flow:7.2f} ms | propagate (2 x 15 x 61 convs + warps) {t_prop:7.2f} ms | "
      f"reconstruction (rest) {t_all - t_flow - t_prop:7.2f} ms | clip {t_all:7.2f} ms")
