"""PSNR margin of the bf16 MotionVectorVSR forward against the reference-generated goldens (the tail's 3x3 runs as two tcgen05 convs with a
bf16 partial sum in between)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from oracle import port
from mobilesuperresolution_b200 import video as V
import test_gpu_video as T
torch.set_grad_enabled(False)
for name in ("mvvsr_nf64", "mvvsr_nf16"):
    try:
        meta, ref, sd, x = T._mvvsr_case(name)
    except Exception as e:
        print(name, "unavailable:", e); continue
    m = V.MotionVectorVSR(meta["num_feat"], meta["num_block"]).eval()
    m.load_state_dict(sd, strict=True)
    y = m.cuda().set_precision("bf16")(x.cuda(), *meta["size"]).cpu()
    print(name, "bf16 PSNR vs reference golden: %.2f dB" % port.psnr_db(y, ref), "halves" if getattr(m, "_tail_halves", None) else "generic tail")
