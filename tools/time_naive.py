import sys, tempfile, os
sys.path.insert(0, "/root/repo")
import torch
torch.set_grad_enabled(False)
import mobilesuperresolution_b200 as sr
f = tempfile.NamedTemporaryFile("w", suffix="_idx.txt", delete=False); f.write(repr(([0,1,2,3], [[32,0,3]]*4)) + "\n"); f.close()
torch.manual_seed(0)
m = sr.Naive_model(4, f.name).eval().cuda().set_precision("bf16")
x = torch.rand(1, 5, 3, 180, 320, device="cuda")
y = m(x).clone()
g = sr.Graphed(m, x)
assert torch.equal(g(x), y), "graph replay differs"
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
print(f"Naive_model(4 blocks of 32 ch) 5 x 180x320 -> 720x1280 bf16: eager {t(lambda: m(x)):.2f} ms, graph {t(lambda: g(x)):.2f} ms per clip")
