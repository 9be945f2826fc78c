"""Warm device timing of the WDSR head and tail kernels alone at the bench workload (cfg2)."""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mobilesuperresolution_b200 as sr
torch.set_grad_enabled(False)
P = types.SimpleNamespace(image_mean=0.5, num_channels=3, scale=4, num_blocks=16, num_residual_units=24, width_search=False, pretrained=False)
dev = torch.device("cuda")
m = sr.BASIC_MODEL(P).eval().to(dev).set_precision("bf16")
plan = m.prepare(dev)
x = torch.rand(64, 3, 96, 96, device=dev).bfloat16()
flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def t(fn, reps=20, flush=False):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        if flush: flush_buf.fill_(0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / reps * 1e3


from mobilesuperresolution_b200 import _lib
trunk = plan.head_internal(x, "bf16")   # the kernels' own trunk layout: no conversion inside the timed calls
y = torch.empty(64, 3, 384, 384, device=dev, dtype=torch.bfloat16)
L, st = _lib.lib(), _lib.current_stream_ptr(dev)
def tail():
    _lib.check(L.b200sr_wdsr_tail(plan.handle, trunk.data_ptr(), x.data_ptr(), _lib.BF16, y.data_ptr(), _lib.BF16, 64, 96, 96, _lib.BF16, st))
for fl in (False, True):
    print(f"L2 {'flushed' if fl else 'warm   '}: head {t(lambda: plan.head_internal(x, 'bf16'), flush=fl):7.1f} us   tail {t(tail, flush=fl):7.1f} us   "
          f"forward {t(lambda: plan.forward(x, 'bf16'), flush=fl):7.1f} us")
