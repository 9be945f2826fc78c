"""GPU parity: flow_warp kernels against the golden vectors / oracle."""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def V():
    from mobilesuperresolution_b200 import video
    assert torch.cuda.is_available()
    return video


@pytest.mark.parametrize("mode", ["zeros", "border"])
def test_flow_warp_golden(V, mode):
    from oracle import synth
    meta, arrs = load_golden("flow_warp_small")
    x = torch.from_numpy(synth.synth_input(meta["xshape"], meta["xseed"], meta["xlo"], meta["xhi"])).cuda()
    fl = torch.from_numpy(synth.synth_input(meta["fshape"], meta["fseed"], meta["flo"], meta["fhi"])).cuda()
    y = V.flow_warp(x, fl, padding_mode=mode).cpu().numpy()
    assert np.abs(y - arrs[mode]).max() <= 1e-5


def test_flow_warp_kat4(V, kat):
    g = torch.Generator().manual_seed(9)
    feat = torch.rand(1, 8, 45, 80, generator=g)
    fl = (torch.rand(1, 45, 80, 2, generator=g) - 0.5) * 20
    wz = V.flow_warp(feat.cuda(), fl.cuda()).cpu()
    wb = V.flow_warp(feat.cuda(), fl.cuda(), padding_mode="border").cpu()
    assert abs(float(wz.double().sum()) - kat["KAT4"]["zeros_sum"]) < 5e-2
    assert abs(float(wb.double().sum()) - kat["KAT4"]["border_sum"]) < 5e-2
    assert float(wz[0, 0, 0, 0]) == 0.0 and abs(float(wz[0, 3, 22, 40]) - kat["KAT4"]["w[0,3,22,40]"]) < 1e-5


@pytest.mark.parametrize("mode", ["zeros", "border"])
def test_flow_warp_permuted_view_and_oob(V, mode):
    """The callers pass flow.permute(0,2,3,1) of an (n,2,h,w) tensor (basicvsr_arch_origin.py:68): consumed as a view.
    Flows of +-40 px on a 31x50 image push most samples out of the image."""
    from oracle import port
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 7, 31, 50, generator=g)
    fl = (torch.rand(2, 2, 31, 50, generator=g) - 0.5) * 80
    ref = port.flow_warp(x, fl.permute(0, 2, 3, 1), padding_mode=mode)
    y = V.flow_warp(x.cuda(), fl.cuda().permute(0, 2, 3, 1), padding_mode=mode).cpu()
    assert float((y - ref).abs().max()) <= 1e-5


def test_flow_warp_asserts_like_reference(V):
    with pytest.raises(AssertionError):
        V.flow_warp(torch.zeros(1, 2, 4, 4).cuda(), torch.zeros(1, 5, 4, 2).cuda())
    with pytest.raises(RuntimeError):
        V.flow_warp(torch.zeros(1, 2, 4, 4), torch.zeros(1, 4, 4, 2))          # CPU tensors: no fallback


@pytest.mark.parametrize("dtype,c", [(torch.float32, 64), (torch.bfloat16, 64), (torch.bfloat16, 24), (torch.float32, 8)])
@pytest.mark.parametrize("mode", ["zeros", "border"])
def test_flow_warp_nhwc(V, dtype, c, mode):
    from oracle import port
    g = torch.Generator().manual_seed(4)
    x = torch.randn(2, c, 45, 77, generator=g)
    fl = (torch.rand(2, 2, 45, 77, generator=g) - 0.5) * 30
    xq = x.to(dtype).float()
    ref = port.flow_warp(xq, fl.permute(0, 2, 3, 1), padding_mode=mode)
    y = V.flow_warp_nhwc(xq.permute(0, 2, 3, 1).contiguous().to(dtype).cuda(), fl.cuda(), padding_mode=mode)
    y = y.float().cpu().permute(0, 3, 1, 2)
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    assert float((y - ref).abs().max()) <= tol


def test_flow_warp_full_size_identity_and_shift(V):
    """cfg4 size (64ch, 180x320) against the oracle, plus two size-independent properties: zero flow is the identity
    and an integer flow a pure shift -- both only up to the bilinear leakage of the reference's own fp32
    normalise/un-normalise round trip (position error ~2e-5 px at x~300 times the local gradient), which the kernel
    replays rather than "fixes"."""
    from oracle import port
    xc = torch.randn(1, 64, 180, 320, generator=torch.Generator().manual_seed(8))
    x = xc.cuda()
    z = torch.zeros(1, 180, 320, 2, device="cuda")
    y0 = V.flow_warp(x, z)
    assert float((y0 - x).abs().max()) <= 5e-4
    assert float((y0.cpu() - port.flow_warp(xc, z.cpu())).abs().max()) <= 2e-5
    s = z.clone()
    s[..., 0], s[..., 1] = 3.0, -2.0
    y = V.flow_warp(x, s)
    assert float((y[:, :, 2:, :-3] - x[:, :, :-2, 3:]).abs().max()) <= 5e-4
    assert float(y[:, :, :1].abs().max()) <= 5e-4 and float(y[:, :, :, -2:].abs().max()) <= 5e-4
    fl = (torch.rand(1, 180, 320, 2, generator=torch.Generator().manual_seed(9)) - 0.5) * 30
    for mode in ("zeros", "border"):
        ref = port.flow_warp(xc, fl, padding_mode=mode)
        assert float((V.flow_warp(x, fl.cuda(), padding_mode=mode).cpu() - ref).abs().max()) <= 2e-5


@pytest.mark.parametrize("dtype,c,xc,xo,yc,yo", [(torch.bfloat16, 64, 128, 64, 80, 16), (torch.bfloat16, 32, 64, 0, 64, 32), (torch.bfloat16, 24, 48, 24, 32, 8),
                                                (torch.float32, 16, 24, 8, 20, 4), (torch.bfloat16, 16, 48, 32, 16, 0)])
def test_flow_warp_nhwc_channel_windows(V, dtype, c, xc, xo, yc, yo):
    """x read from channels [xo, xo + c) of a wider tensor and y written into channels [yo, yo + c) of another one (the in-place concatenations
    of the BasicVSR propagation): equal to the contiguous call bit for bit, and the other channels of y untouched.  Covers the lean bf16 kernel
    (c = 64, 32, 16 with 32-byte aligned windows), its fallback (a 16-byte aligned window) and the general kernel (c = 24, fp32)."""
    g = torch.Generator().manual_seed(14)
    wide = torch.randn(2, 37, 53, xc, generator=g).to(dtype).cuda()
    fl = ((torch.rand(2, 2, 37, 53, generator=g) - 0.5) * 12).cuda()
    xwin = wide[..., xo:xo + c]
    ref = V.flow_warp_nhwc(xwin.contiguous(), fl)
    out = torch.full((2, 37, 53, yc), 7.0, dtype=dtype, device="cuda")
    V.flow_warp_nhwc(xwin, fl, out=out, out_coff=yo)
    torch.cuda.synchronize()
    assert torch.equal(out[..., yo:yo + c], ref)
    rest = torch.cat([out[..., :yo], out[..., yo + c:]], -1)
    assert rest.numel() == 0 or bool((rest == 7.0).all())

