"""First-contact GPU diagnostics: per-stage errors vs the oracle, printed (not asserted)."""
import sys, os, types, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
import mobilesuperresolution_b200 as sr
from oracle import port, synth

def P(scale, nb): return types.SimpleNamespace(image_mean=0.5, num_channels=3, scale=scale, num_blocks=nb, num_residual_units=24, width_search=False, pretrained=False)
torch.set_grad_enabled(False)
print(torch.cuda.get_device_name(0))
for prec in ("fp32", "bf16"):
    try:
        m = sr.BASIC_MODEL(P(4, 2)).eval()
        shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        sd = {k: torch.from_numpy(v) for k, v in synth.synth_state_dict(shapes, 51).items()}
        m.load_state_dict(sd); m = m.to("cuda").set_precision(prec); plan = m.prepare()
        x = torch.from_numpy(synth.synth_input((2, 3, 37, 45), 52)); xd = x.cuda(); x0 = x - 0.5
        rh = F.conv2d(x0, port.weight_norm_fold(sd["head.weight_g"], sd["head.weight_v"]), sd["head.bias"], padding=1)
        t = plan.head(xd, prec); torch.cuda.synchronize()
        print(prec, "head maxabs", float((t.float().cpu().permute(0,3,1,2) - rh).abs().max()))
        tin = rh.permute(0,2,3,1).contiguous().cuda()
        if prec == "bf16": tin = tin.bfloat16()
        rb = port.wdsr_block(sd, "body.0.", tin.float().cpu().permute(0,3,1,2))
        gb = plan.block(0, tin, prec).float().cpu().permute(0,3,1,2); torch.cuda.synchronize()
        d = (gb - rb).abs()
        print(prec, "block maxabs", float(d.max()), "psnr", port.psnr_db(gb, rb), "interior maxabs", float(d[:, :, 2:-2, 2:-2].max()))
        rt = F.pixel_shuffle(F.conv2d(rb, port.weight_norm_fold(sd["tail.weight_g"], sd["tail.weight_v"]), sd["tail.bias"], padding=1) +
                             F.conv2d(x0, port.weight_norm_fold(sd["skip.0.weight_g"], sd["skip.0.weight_v"]), sd["skip.0.bias"], padding=2), 4) + 0.5
        tt = rb.permute(0,2,3,1).contiguous().cuda(); xin = xd
        if prec == "bf16": tt, xin = tt.bfloat16(), xd.bfloat16()
        gt = plan.tail(tt, xin, prec).float().cpu(); torch.cuda.synchronize()
        d = (gt - rt).abs()
        print(prec, "tail maxabs", float(d.max()), "psnr", port.psnr_db(gt, rt), "interior", float(d[:, :, 12:-12, 12:-12].max()))
        y = m(xd if prec == "fp32" else xd.bfloat16()).float().cpu()
        ref = port.basic_model_forward(sd, x, 4)
        print(prec, "model maxabs", float((y - ref).abs().max()), "psnr", port.psnr_db(y, ref))
    except Exception:
        traceback.print_exc()
