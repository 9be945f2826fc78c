"""Chained block launch (B200SR_BLOCK_IMPL=chain, csrc/wdsr_tc5c.cuh) vs one launch per block: bit-equality and timing."""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.set_grad_enabled(False)
import mobilesuperresolution_b200 as sr

def P(scale, nb): return types.SimpleNamespace(image_mean=0.5, num_channels=3, scale=scale, num_blocks=nb, num_residual_units=24, width_search=False, pretrained=False)

def model(impl, scale, nb, seed=0):
    os.environ["B200SR_BLOCK_IMPL"] = impl
    torch.manual_seed(seed)
    m = sr.BASIC_MODEL(P(scale, nb)).eval().cuda().set_precision("bf16")
    m.prepare()
    return m

for nb, scale, shape in [(1, 4, (1, 3, 16, 32)), (2, 2, (2, 3, 37, 45)), (3, 4, (3, 3, 96, 96)), (16, 4, (64, 3, 96, 96)), (5, 2, (1, 3, 131, 200)), (16, 4, (1, 3, 360, 640))]:
    a, b = model("tc5", scale, nb), model("chain", scale, nb)
    x = torch.rand(*shape, device="cuda").bfloat16()
    ya, yb = a(x), b(x)
    torch.cuda.synchronize()
    same = torch.equal(ya, yb)
    print(f"nb={nb} x{scale} {shape}: identical={same} maxdiff={float((ya.float() - yb.float()).abs().max()):.3e}", flush=True)
    if shape[0] * shape[2] * shape[3] >= 96 * 96 * 3:
        gs = {"per-block": sr.Graphed(a, x), "chain": sr.Graphed(b, x)}
        for _ in range(200): gs["per-block"](x)          # clocks up
        torch.cuda.synchronize()
        for rnd in range(3):
            for name, g in gs.items():
                for _ in range(5): g(x)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(50): g(x)
                e1.record(); torch.cuda.synchronize()
                print(f"    round {rnd} {name:10s} forward (graph replay): {e0.elapsed_time(e1) / 50 * 1e3:8.1f} us", flush=True)
