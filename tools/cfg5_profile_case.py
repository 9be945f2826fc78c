import os, sys, types
sys.path.insert(0, "/root/repo")
import torch
import mobilesuperresolution_b200 as sr
torch.set_grad_enabled(False)
P = lambda s: types.SimpleNamespace(image_mean=0.5, num_channels=3, scale=s, num_blocks=16, num_residual_units=24, width_search=False, pretrained=False)
m = sr.BASIC_MODEL(P(2)).eval().cuda().set_precision("bf16")
x = torch.rand(1, 3, 1080, 1920, device="cuda").bfloat16()
for _ in range(2): y = m(x)
torch.cuda.synchronize()
m4 = sr.BASIC_MODEL(P(4)).eval().cuda().set_precision("bf16")
x = torch.rand(1, 3, 360, 640, device="cuda").bfloat16()
for _ in range(2): y = m4(x)
torch.cuda.synchronize()
