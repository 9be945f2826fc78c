"""One MotionVectorVSR(64,15) bf16 forward of a 15-frame 180x320 clip for the ncu launch list (which kernels the fork's video model spends its time in)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mobilesuperresolution_b200 import video
torch.set_grad_enabled(False)
dev = torch.device("cuda")
m = video.MotionVectorVSR(64, 15).to(dev).eval().set_precision("bf16")
xm = torch.rand(1, 15, 5, 180, 320, device=dev)
xm[:, :, 3:] = (xm[:, :, 3:] - 0.5) * 8
y = m(xm, 720, 1280)
torch.cuda.synchronize()
y = m(xm, 720, 1280)
torch.cuda.synchronize()
print(tuple(y.shape))
