"""cfg3 (searched widths P1, x4, 360p frames) per block-kernel form: B200SR_BLOCK_IMPL=tc5 | tc5q | rs | rh, CUDA-graph replay."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.set_grad_enabled(False)
import mobilesuperresolution_b200 as sr
P1 = [(9, 91, 14), (9, 94, 10), (9, 107, 12), (9, 110, 13), (9, 115, 12), (9, 94, 12), (9, 115, 17), (9, 116, 16)]
f = tempfile.NamedTemporaryFile("w", suffix="_block_index.txt", delete=False)
f.write(repr((list(range(len(P1))), [list(w) for w in P1])) + "\n"); f.close()
outs = {}
for impl in sys.argv[1:] or ["tc5", "tc5q"]:
    os.environ["B200SR_BLOCK_IMPL"] = impl
    torch.manual_seed(0)
    m = sr.Model(4, f.name).eval().cuda().set_precision("bf16")
    for b in (1, 8):
        torch.manual_seed(1)
        x = torch.rand(b, 3, 360, 640, device="cuda").bfloat16()
        y = m(x)
        outs.setdefault(b, {})[impl] = y.float().cpu()
        g = sr.Graphed(m, x)
        for _ in range(20): g(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(30): g(x)
        e1.record(); torch.cuda.synchronize()
        print(f"{impl:5s} P1 x4 batch {b}: {e0.elapsed_time(e1) / 30 / b * 1e3:7.1f} us per 360p frame", flush=True)
for b, d in outs.items():
    ks = list(d)
    for k in ks[1:]:
        print(f"batch {b}: max |{k} - {ks[0]}| = {float((d[k] - d[ks[0]]).abs().max()):.4f}")
os.unlink(f.name)
