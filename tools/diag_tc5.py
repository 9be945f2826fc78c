"""tcgen05 block kernel vs oracle + timing (run with B200SR_BLOCK_IMPL=tc5seq|tc5)."""
import sys, os, types, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mobilesuperresolution_b200 as sr
from mobilesuperresolution_b200 import _lib
from oracle import port, synth
torch.set_grad_enabled(False)
print("impl", os.environ.get("B200SR_BLOCK_IMPL"), torch.cuda.get_device_name(0))
def P(scale, nb): return types.SimpleNamespace(image_mean=0.5, num_channels=3, scale=scale, num_blocks=nb, num_residual_units=24, width_search=False, pretrained=False)
m = sr.BASIC_MODEL(P(4, 2)).eval()
shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
sd = {k: torch.from_numpy(v) for k, v in synth.synth_state_dict(shapes, 51).items()}
m.load_state_dict(sd); m = m.to("cuda").set_precision("bf16"); plan = m.prepare()
for shape in [(1, 24, 16, 32), (2, 24, 37, 45), (1, 24, 96, 96)]:
    g = torch.Generator().manual_seed(3)
    t = (torch.rand(shape, generator=g) - 0.5) * 2
    tin = t.permute(0, 2, 3, 1).contiguous().cuda().bfloat16()
    ref = port.wdsr_block(sd, "body.0.", tin.float().cpu().permute(0, 3, 1, 2))
    got = plan.block(0, tin, "bf16"); torch.cuda.synchronize()
    got = got.float().cpu().permute(0, 3, 1, 2)
    d = (got - ref).abs()
    print(shape, "block maxabs", float(d.max()), "psnr", port.psnr_db(got, ref), "argmax", [int(v) for v in torch.nonzero(d == d.max())[0]])
    if float(d.max()) > 0.1:
        bad = (d > 0.1)
        print("  bad fraction", float(bad.float().mean()), "bad per channel", bad.float().mean(dim=(0, 2, 3)).tolist()[:24])
        print("  bad rows", bad.float().mean(dim=(0, 1, 3)).tolist()[:20])
        print("  bad cols", bad.float().mean(dim=(0, 1, 2)).tolist()[:40])
x = torch.from_numpy(synth.synth_input((2, 3, 37, 45), 52))
y = m(x.cuda().bfloat16()).float().cpu(); ref = port.basic_model_forward(sd, x, 4)
print("model psnr", port.psnr_db(y, ref))
# timing at cfg2 size
torch.manual_seed(0)
mm = sr.BASIC_MODEL(P(4, 16)).eval().to("cuda").set_precision("bf16"); pl = mm.prepare()
tr = torch.randn(64, 96, 96, 24, device="cuda").bfloat16(); tb = torch.empty_like(tr)
for _ in range(3): pl.block(0, tr, "bf16")
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for i in range(16):
    _lib.check(_lib.lib().b200sr_wdsr_block(pl.handle, i, tr.data_ptr(), tb.data_ptr(), 64, 96, 96, _lib.BF16, _lib.current_stream_ptr(tr.device)))
b.record(); torch.cuda.synchronize()
print("block us/launch @cfg2 (eager launches)", a.elapsed_time(b) / 16 * 1e3)
def graph_time(shape):
    tr = torch.randn(*shape, device="cuda").bfloat16(); tb = torch.empty_like(tr)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for _ in range(2): pl.block(0, tr, "bf16")
        st.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            src, dst = tr, tb
            for i in range(16):
                _lib.check(_lib.lib().b200sr_wdsr_block(pl.handle, i, src.data_ptr(), dst.data_ptr(), shape[0], shape[1], shape[2], _lib.BF16, _lib.current_stream_ptr(tr.device)))
                src, dst = dst, src
        for _ in range(3): g.replay()
        st.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(10): g.replay()
        e1.record(st); st.synchronize()
    return e0.elapsed_time(e1) / 160 * 1e3
for nm, shape in [("cfg2", (64, 96, 96, 24)), ("360p", (1, 360, 640, 24)), ("1080p", (1, 1080, 1920, 24))]:
    print("block us/launch (CUDA graph) @", nm, graph_time(shape))
for nm, shape in [("360p", (1, 360, 640, 24)), ("1080p", (1, 1080, 1920, 24))]:
    tr = torch.randn(*shape, device="cuda").bfloat16(); tb = torch.empty_like(tr)
    for _ in range(2): pl.block(0, tr, "bf16")
    torch.cuda.synchronize(); a.record()
    for i in range(16):
        _lib.check(_lib.lib().b200sr_wdsr_block(pl.handle, i, tr.data_ptr(), tb.data_ptr(), shape[0], shape[1], shape[2], _lib.BF16, _lib.current_stream_ptr(tr.device)))
    b.record(); torch.cuda.synchronize()
    print("block us/launch @", nm, a.elapsed_time(b) / 16 * 1e3)
