"""One BasicVSR_origin(64,30) bf16 forward of a short clip (3 x 180 x 320 -> 720 x 1280) for the ncu launch list
(profiles/r01_launches_basicvsr_clip3.csv): which kernels a clip spends its time in."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mobilesuperresolution_b200 import video
torch.set_grad_enabled(False)
dev = torch.device("cuda")
m = video.BasicVSR_origin(64, 30).to(dev).eval().set_precision("bf16")
nf = int(sys.argv[1]) if len(sys.argv) > 1 else 3
x = torch.rand(1, nf, 3, 180, 320, device=dev)
y = m(x, 720, 1280)
torch.cuda.synchronize()
print(tuple(y.shape), float(y.abs().mean()))
