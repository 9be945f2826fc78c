"""Shared helpers for the -m gpu parity tests (CUDA path vs oracle, through the C ABI)."""
import tempfile
import types

import numpy as np
import torch

import mobilesuperresolution_b200 as sr
from oracle import port, synth

DEV = "cuda:0"


def params(scale, nb, nru=24, width_search=False):
    return types.SimpleNamespace(image_mean=0.5, num_channels=3, scale=scale, num_blocks=nb, num_residual_units=nru,
                                 width_search=width_search, pretrained=False, model_type="BASIC_MODEL")


def load_np_state(module, sd_np):
    module.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd_np.items()}, strict=True)
    return module


def synth_load(module, seed):
    shapes = {k: tuple(v.shape) for k, v in module.state_dict().items()}
    sd = synth.synth_state_dict(shapes, seed)
    load_np_state(module, sd)
    return {k: torch.from_numpy(v) for k, v in sd.items()}


def block_index_file(widths):
    f = tempfile.NamedTemporaryFile("w", suffix="_block_index.txt", delete=False)
    f.write("('header line', [])\n")
    f.write(repr((list(range(len(widths))), [list(w) for w in widths])) + "\n")
    f.close()
    return f.name


def run(model, x_cpu, precision):
    model = model.to(DEV).eval().set_precision(precision)
    x = x_cpu.to(DEV)
    if precision == "bf16":
        x = x.bfloat16()
    with torch.no_grad():
        y = model(x)
    if isinstance(y, tuple):
        y = y[0]
    torch.cuda.synchronize()
    return y.float().cpu()


def maxabs(a, b):
    return float((a - b).abs().max())


psnr = port.psnr_db
