"""Head and tail kernels of cfg2 timed as CUDA-graph replays of 20 launches each (no host launch gaps), L2 warm."""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.set_grad_enabled(False)
import mobilesuperresolution_b200 as sr
from mobilesuperresolution_b200 import _lib
P = types.SimpleNamespace(image_mean=0.5, num_channels=3, scale=4, num_blocks=16, num_residual_units=24, width_search=False, pretrained=False)
torch.manual_seed(0)
m = sr.BASIC_MODEL(P).eval().cuda().set_precision("bf16")
plan = m.prepare()
dev = torch.device("cuda")
x = torch.rand(64, 3, 96, 96, device=dev).bfloat16()
trunk = plan.head_internal(x, "bf16")
y = torch.empty(64, 3, 384, 384, device=dev, dtype=torch.bfloat16)
L = _lib.lib()
st = torch.cuda.Stream()
def tail(): _lib.check(L.b200sr_wdsr_tail(plan.handle, trunk.data_ptr(), x.data_ptr(), _lib.BF16, y.data_ptr(), _lib.BF16, 64, 96, 96, _lib.BF16, _lib.current_stream_ptr(dev)))
trunk2 = torch.empty_like(trunk)
def head(): _lib.check(L.b200sr_wdsr_head(plan.handle, x.data_ptr(), _lib.BF16, trunk2.data_ptr(), 64, 96, 96, _lib.BF16, _lib.current_stream_ptr(dev)))
for name, fn in (("tail", tail), ("head", head)):
    with torch.cuda.stream(st):
        for _ in range(3): fn()
        st.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for _ in range(20): fn()
        for _ in range(3): g.replay()
        st.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        for _ in range(10): g.replay()
        b.record(st); st.synchronize()
        print(f"{name}: {a.elapsed_time(b) / 200 * 1e3:7.2f} us per launch (graph of 20, 10 replays)")
