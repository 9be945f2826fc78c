"""Random-shape sweep of the Split_Block bf16 tensor-core arm against the fp32 oracle and the FFMA arm (one-off robustness check)."""
import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mobilesuperresolution_b200 as sr
from oracle import synth, port
torch.set_grad_enabled(False)
random.seed(1)
worst = 1e9
for trial in range(60):
    c = random.choice([8, 16, 24, 32]); n = random.randint(1, 3); h = random.randint(1, 40); w = 8 * random.randint(1, 12)
    m = sr.Split_Block(num_residual_units=c, kernel_size=3).eval()
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    sd = {k: torch.from_numpy(v) for k, v in synth.synth_state_dict(shapes, trial).items()}
    m.load_state_dict(sd); m = m.cuda()
    x = torch.from_numpy(synth.synth_input((n, c, h, w), 100 + trial, -1.0, 1.0)).bfloat16()
    y = m(x.cuda()).float().cpu()
    os.environ["B200SR_SPLIT_IMPL"] = "ffma"
    m2 = sr.Split_Block(num_residual_units=c, kernel_size=3).eval(); m2.load_state_dict(sd); y2 = m2.cuda()(x.cuda()).float().cpu()
    del os.environ["B200SR_SPLIT_IMPL"]
    ref = port.split_block(sd, "", x.float())
    p, p2 = port.psnr_db(y, ref), port.psnr_db(y2, ref)
    worst = min(worst, p)
    if p < 50 or not torch.isfinite(y).all(): print("FAIL", c, n, h, w, p, p2)
print("worst PSNR tc arm vs oracle over 60 random shapes:", worst)
