"""The oracle port against the LIVE reference (runs wherever /root/reference exists -- the build container; skipped on the GPU box).

oracle/make_golden*.py assert the same equalities when fixtures are generated; this keeps them asserted on every CPU test run, so an
edit of oracle/port.py cannot drift from the reference unnoticed.  Bit-exact: both sides are the same torch CPU ops in the same order.
"""
import types
import warnings

import numpy as np
import pytest
import torch

from oracle import port, ref_import, synth

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="reference tree not present (GPU box)")
warnings.filterwarnings("ignore")


def _load_synth(module, seed):
    shapes = {k: tuple(v.shape) for k, v in module.state_dict().items()}
    sd = {k: torch.from_numpy(v) for k, v in synth.synth_state_dict(shapes, seed).items()}
    module.load_state_dict(sd, strict=True)
    return sd


def _params(scale, nb, ws=False):
    return types.SimpleNamespace(image_mean=0.5, num_channels=3, scale=scale, num_blocks=nb, num_residual_units=24, width_search=ws,
                                 pretrained=False)


@torch.no_grad()
def test_basic_model_port_is_the_reference():
    R = ref_import.modules()
    for scale, nb, shape in [(4, 2, (1, 3, 17, 23)), (2, 3, (2, 3, 12, 9))]:
        m = R.BASIC_MODEL(_params(scale, nb)).eval()
        sd = _load_synth(m, 100 + nb)
        x = torch.from_numpy(synth.synth_input(shape, 7))
        assert torch.equal(m(x), port.basic_model_forward(sd, x, scale))


@torch.no_grad()
def test_seeded_kat1_reproduces():
    """the reference's own seeded init, one 64x64 patch."""
    import json
    import os
    R = ref_import.modules()
    kat = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "kat.json")))["KAT1"]
    torch.manual_seed(0)
    m = R.BASIC_MODEL(_params(4, 16)).eval()
    x = torch.rand(1, 3, 64, 64, generator=torch.Generator().manual_seed(1234))
    y = m(x)
    assert abs(float(y.double().sum()) - kat["sum"]) < 1e-6
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    assert torch.equal(y, port.basic_model_forward(sd, x, 4))


@torch.no_grad()
def test_masked_block_flow_warp_spynet_ports():
    R = ref_import.modules()
    b = R.wdsr_b.Block(num_residual_units=24, kernel_size=3, res_scale=0.25, width_search=True).eval()
    sd = _load_synth(b, 11)
    x = torch.from_numpy(synth.synth_input((1, 24, 9, 11), 12, -1.0, 1.0))
    assert torch.equal(b(x), port.wdsr_block_masked(sd, "", x))
    f = torch.from_numpy(synth.synth_input((2, 5, 12, 17), 21, -1.0, 1.0))
    fl = torch.from_numpy(synth.synth_input((2, 12, 17, 2), 22, -9.0, 9.0))
    for pad in ("zeros", "border"):
        assert torch.equal(R.spynet_arch.flow_warp(f, fl, padding_mode=pad), port.flow_warp(f, fl, padding_mode=pad))
    sp = R.spynet_arch.SpyNet().eval()
    sd = _load_synth(sp, 31)
    a, c = torch.from_numpy(synth.synth_input((1, 3, 40, 72), 32)), torch.from_numpy(synth.synth_input((1, 3, 40, 72), 33))
    assert torch.equal(sp(a, c), port.spynet_forward(sd, a, c))


@torch.no_grad()
def test_split_block_and_fork_nas_model_ports():
    R = ref_import.modules()
    m = R.wdsr_b.Split_Block(num_residual_units=24, kernel_size=3).eval()
    sd = _load_synth(m, 21)
    x = torch.from_numpy(synth.synth_input((1, 24, 13, 15), 121, -1.0, 1.0))
    assert torch.equal(m(x), port.split_block(sd, "", x))
    from oracle.make_golden_r2 import nas_reference
    p = _params(2, 2, ws=True)
    nas, real_cuda = nas_reference(p)
    try:
        est = {k: v for k, v in nas.state_dict().items() if k.startswith("speed_estimator.")}
        sd = _load_synth(nas, 5)
        sd.update(est)
        nas.load_state_dict(sd)
        x = torch.from_numpy(synth.synth_input((1, 3, 10, 14), 6))
        out, speed = nas(x)
        pout, pspeed = port.nas_fork_forward(sd, x, 2)
        assert torch.equal(out, pout) and torch.equal(speed, pspeed)
    finally:
        torch.Tensor.cuda = real_cuda


@torch.no_grad()
def test_basicvsr_ports():
    R = ref_import.modules()
    vs = R.basicvsr_origin.BasicVSR_origin(num_feat=8, num_block=1).eval()
    sd = _load_synth(vs, 41)
    clip = torch.from_numpy(synth.synth_input((1, 2, 3, 64, 64), 42))
    assert torch.equal(vs(clip, 100, 90), port.basicvsr_origin_forward(sd, clip, 100, 90))
    fk = R.basicvsr_fork.BasicVSR(num_feat=3, num_block=1).eval()
    sd = _load_synth(fk, 43)
    assert torch.equal(fk(clip, 70, 120), port.basicvsr_fork_forward(sd, clip, 70, 120))
    with pytest.raises(RuntimeError):
        R.basicvsr_fork.BasicVSR(num_feat=8, num_block=1).eval()(clip, 70, 120)     # broken as committed for num_feat != 3


@torch.no_grad()
def test_naive_model_port(tmp_path):
    """oracle.port.naive_model_forward == the live reference Naive_model (harness shims: stub modules for the training stack its file imports,
    Tensor.to('cuda') as a no-op on this CPU-only box -- the reference hard-codes it at models/naive_multi_model_easy.py:127)."""
    Naive = ref_import.naive_model()
    blocks = [[8, 0, 3], [8, 0, 5]]
    f = tmp_path / "naive_index.txt"
    f.write_text(repr(([0, 1], blocks)) + "\n")
    real_to = torch.Tensor.to
    torch.Tensor.to = lambda self, *a, **k: self if (a and a[0] == "cuda") else real_to(self, *a, **k)
    try:
        m = Naive(4, str(f)).eval()
        sd = _load_synth(m, 45)
        clip = torch.from_numpy(synth.synth_input((2, 3, 3, 64, 64), 46))
        assert torch.equal(m(clip), port.naive_model_forward(sd, clip, len(blocks)))
    finally:
        torch.Tensor.to = real_to

