"""Forward latency of the BASELINE configs through the public modules (CUDA-graph replay, device-resident)."""
import os, sys, types, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mobilesuperresolution_b200 as sr
torch.set_grad_enabled(False)
def P(scale, nb): return types.SimpleNamespace(image_mean=0.5, num_channels=3, scale=scale, num_blocks=nb, num_residual_units=24, width_search=False, pretrained=False)
def pruned(scale, widths):
    f = tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False); f.write(repr((list(range(len(widths))), [list(w) for w in widths])) + "\n"); f.close()
    m = sr.Model(scale, f.name); os.unlink(f.name); return m
P1 = [(9, 91, 14), (9, 94, 10), (9, 107, 12), (9, 110, 13), (9, 115, 12), (9, 94, 12), (9, 115, 17), (9, 116, 16)]
P2 = [(14, 123, 20), (14, 126, 20), (14, 137, 20), (14, 141, 20), (14, 144, 20)]
def graph_time(model, x, prec, reps=20):
    model = model.cuda().eval().set_precision(prec)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for _ in range(3): y = model(x)
        st.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            y = model(x)
        for _ in range(3): g.replay()
        st.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(reps): g.replay()
        e1.record(st); st.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
cases = [("cfg1 x4 1x64x64 fp32", sr.BASIC_MODEL(P(4, 16)), (1, 3, 64, 64), "fp32"),
         ("cfg1 x4 1x64x64 bf16", sr.BASIC_MODEL(P(4, 16)), (1, 3, 64, 64), "bf16"),
         ("cfg2 x4 64x96x96 bf16", sr.BASIC_MODEL(P(4, 16)), (64, 3, 96, 96), "bf16"),
         ("dense x4 360p bf16", sr.BASIC_MODEL(P(4, 16)), (1, 3, 360, 640), "bf16"),
         ("dense x4 360p bf16 B=8", sr.BASIC_MODEL(P(4, 16)), (8, 3, 360, 640), "bf16"),
         ("dense x4 360p fp32", sr.BASIC_MODEL(P(4, 16)), (1, 3, 360, 640), "fp32"),
         ("cfg3 P1 x4 360p bf16", pruned(4, P1), (1, 3, 360, 640), "bf16"),
         ("cfg3 P1 x4 360p bf16 B=8", pruned(4, P1), (8, 3, 360, 640), "bf16"),
         ("cfg3 P2 x4 360p bf16", pruned(4, P2), (1, 3, 360, 640), "bf16"),
         ("cfg5 x2 1080p bf16 B=1", sr.BASIC_MODEL(P(2, 16)), (1, 3, 1080, 1920), "bf16"),
         ("cfg5 x2 1080p bf16 B=4", sr.BASIC_MODEL(P(2, 16)), (4, 3, 1080, 1920), "bf16")]
print("pad24 =", os.environ.get("B200SR_TRUNK_PAD24"))
for name, m, shape, prec in cases:
    x = torch.rand(*shape, device="cuda")
    if prec == "bf16": x = x.bfloat16()
    us = graph_time(m, x, prec)
    n, _, h, w = shape; s = m.scale
    print(f"{name:28s} {us:10.1f} us/forward  {n / us * 1e6:10.1f} frames/s  {n * s * s * h * w / us:10.1f} Mpix/s out")
