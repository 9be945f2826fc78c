"""One forward of the fork's NAS_MODEL (16 kept Split_Blocks, x4) on 8 x 360p bf16 frames for the ncu launch list."""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mobilesuperresolution_b200 as sr
torch.set_grad_enabled(False)
p = types.SimpleNamespace(image_mean=0.5, num_channels=3, scale=4, num_blocks=16, num_residual_units=24, width_search=True, pretrained=False)
torch.manual_seed(0)
m = sr.NAS_MODEL(p).eval().cuda().set_precision("bf16")
x = torch.rand(8, 3, 360, 640, device="cuda").bfloat16()
for _ in range(2):
    y, s = m(x)
torch.cuda.synchronize()
print(tuple(y.shape), float(s))
