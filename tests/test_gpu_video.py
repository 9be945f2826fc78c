"""GPU parity: generic conv, SPyNet and BasicVSR (video path) against golden vectors / the oracle."""
import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from conftest import golden_case, load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def V():
    from mobilesuperresolution_b200 import video
    assert torch.cuda.is_available()
    return video


def _t(sd):
    return {k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}


@pytest.mark.parametrize("cin,cout,k", [(8, 32, 7), (32, 64, 7), (16, 2, 7), (67, 64, 3), (64, 64, 3), (128, 64, 1), (64, 256, 3), (64, 3, 3), (5, 9, 3), (16, 48, 5), (34, 16, 5)])
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_conv_vs_torch(V, cin, cout, k, precision):
    g = torch.Generator().manual_seed(cin * 100 + cout + k)
    conv = nn.Conv2d(cin, cout, k, 1, k // 2)
    x = torch.randn(2, cin, 19, 37, generator=g)
    res = torch.randn(2, cout, 19, 37, generator=g)
    dt = torch.float32 if precision == "fp32" else torch.bfloat16
    xq = x.to(dt).float()
    resq = res.to(dt).float()
    h = V._ConvHandle(conv, torch.device("cuda:0"))
    xn = xq.permute(0, 2, 3, 1).contiguous().to(dt).cuda()
    with torch.no_grad():
        ref = F.leaky_relu(F.conv2d(xq, conv.weight, conv.bias, padding=k // 2), 0.1) + resq
    y = h(xn, precision, V.ACT_LRELU, residual=resq.permute(0, 2, 3, 1).contiguous().to(dt).cuda()).float().cpu().permute(0, 3, 1, 2)
    err = float((y - ref).abs().max())
    assert err <= (2e-5 if precision == "fp32" else 0.05), err
    if cout % 4 == 0:   # PixelShuffle(2) folded into the store
        with torch.no_grad():
            ref2 = F.pixel_shuffle(F.relu(F.conv2d(xq, conv.weight, conv.bias, padding=k // 2)), 2)
        y2 = h(xn, precision, V.ACT_RELU, shuffle=2).float().cpu().permute(0, 3, 1, 2)
        assert float((y2 - ref2).abs().max()) <= (2e-5 if precision == "fp32" else 0.05)


def test_conv_channel_windows(V):
    """x and y may be channel windows of wider NHWC tensors (concatenation without copies)."""
    conv = nn.Conv2d(6, 10, 3, 1, 1)
    x = torch.randn(1, 16, 9, 11)
    h = V._ConvHandle(conv, torch.device("cuda:0"))
    xn = x.permute(0, 2, 3, 1).contiguous().cuda()
    out = torch.full((1, 9, 11, 24), 7.0, device="cuda")
    h(xn, "fp32", V.ACT_NONE, x_coff=5, out=out, y_coff=3)
    with torch.no_grad():
        ref = F.conv2d(x[:, 5:11], conv.weight, conv.bias, padding=1)
    o = out.cpu()
    assert float((o[..., 3:13].permute(0, 3, 1, 2) - ref).abs().max()) <= 2e-5
    assert float((o[..., :3] - 7).abs().max()) == 0 and float((o[..., 13:] - 7).abs().max()) == 0


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_spynet_golden(V, precision):
    from oracle import synth
    meta, arrs = load_golden("spynet_small")
    sp = V.SpyNet().eval()
    sp.load_state_dict(_t(synth.synth_state_dict(meta["shapes"], meta["wseed"])))
    a, b = synth.synth_input(meta["shape"], meta["aseed"]), synth.synth_input(meta["shape"], meta["bseed"])
    sp = sp.cuda().set_precision(precision)
    f = sp(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()).cpu().numpy()
    err = np.abs(f - arrs["flow"]).max()
    if precision == "fp32":
        assert err <= 1e-4, err
    else:
        assert err <= 0.1, err          # flows of up to 3.9 px; bf16 conv operands, fp32 flow arithmetic


def test_spynet_kat3_180x320(V, kat):
    """reference-seeded SpyNet on 2 x 180x320 pairs (the cfg4 frame size), fp32."""
    meta, arrs = load_golden("kat3_spynet_seed0")
    torch.manual_seed(0)
    sp = V.SpyNet().eval()
    assert abs(float(sum(v.double().sum() for v in sp.state_dict().values())) - kat["KAT3"]["weights_sum"]) < 1e-6
    g = torch.Generator().manual_seed(7)
    a = torch.rand(2, 3, 180, 320, generator=g)
    b = torch.rand(2, 3, 180, 320, generator=g)
    f = sp.cuda()(a.cuda(), b.cuda()).cpu()
    s = meta["stride"]
    assert float((f[:, :, ::s, ::s] - torch.from_numpy(arrs["f_strided"])).abs().max()) <= 1e-4
    assert abs(float(f.double().sum()) - kat["KAT3"]["sum"]) < 1.0
    assert abs(float(f[0, 0, 90, 160]) - kat["KAT3"]["f[0,0,90,160]"]) < 1e-4


def test_spynet_mmedit_key_remap(V):
    sd = {"basic_module.3.basic_module.2.conv.weight": 1, "basic_module.0.basic_module.4.conv.bias": 2, "mean": 3}
    out = V.SpyNet.remap_mmedit_state_dict(sd)
    assert set(out) == {"basic_module.3.basic_module.4.weight", "basic_module.0.basic_module.8.bias", "mean"}


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_basicvsr_origin_golden(V, precision):
    from oracle import port
    meta, arrs, sd, x = golden_case("basicvsr_origin_small")
    m = V.BasicVSR_origin(meta["num_feat"], meta["num_block"]).eval()
    m.load_state_dict(_t(sd))
    m = m.cuda().set_precision(precision)
    h, w = meta["out_hw"]
    y = m(torch.from_numpy(x).cuda(), h, w).cpu()
    s = meta["stride"]
    ref_s, ref_c = torch.from_numpy(arrs["y_strided"]), torch.from_numpy(arrs["y_corner"])
    if precision == "fp32":
        assert float((y[..., ::s, ::s] - ref_s).abs().max()) <= 1e-4
        assert float((y[..., :16, :16] - ref_c).abs().max()) <= 1e-4
    else:
        assert port.psnr_db(y[..., ::s, ::s], ref_s) >= 50.0


def test_basicvsr_origin_resized_output(V):
    """height/weight different from 4h x 4w goes through the final bilinear resize (basicvsr_arch_origin.py:93)."""
    from oracle import port
    meta, arrs, sd, x = golden_case("basicvsr_origin_small")
    m = V.BasicVSR_origin(meta["num_feat"], meta["num_block"]).eval()
    m.load_state_dict(_t(sd))
    xt = torch.from_numpy(x)[:, :2]
    ref = port.basicvsr_origin_forward(_t(sd), xt, 100, 200)
    y = m.cuda()(xt.cuda(), 100, 200).cpu()
    assert float((y - ref).abs().max()) <= 1e-4


def test_fork_basicvsr_flow_propagation_and_faithful_error(V):
    from oracle import port, synth
    m = V.BasicVSR(num_feat=8, num_block=1).eval()
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    sd = _t(synth.synth_state_dict(shapes, 61))
    m.load_state_dict(sd)
    x = torch.from_numpy(synth.synth_input((1, 3, 3, 36, 68), 62))
    m = m.cuda()
    ff, fb = m.get_flow(x.cuda())
    rff, rfb = port.vsr_get_flow(sd, x)
    assert float((ff.cpu() - rff).abs().max()) <= 1e-4 and float((fb.cpu() - rfb).abs().max()) <= 1e-4
    back, fwd = m.propagate(x.cuda(), ff, fb)
    rback, rfwd = port.vsr_propagate(sd, x, rff, rfb, 8)
    for a, r in zip(back + fwd, rback + rfwd):
        assert float((a.float().cpu().permute(0, 3, 1, 2) - r).abs().max()) <= 1e-4
    with pytest.raises(RuntimeError, match="must match the size of tensor b"):
        m(x.cuda(), 144, 272)


def test_cfg4_basicvsr_clip_properties(V):
    """cfg4 size: 15-frame 180x320 clip through BasicVSR_origin(64, 30) in bf16.  Size-independent properties:
    output shape; clips in a batch are independent (b=2 equals two b=1 runs); fp32 and bf16 paths agree to >= 50 dB."""
    from oracle import port
    torch.manual_seed(0)
    m = V.BasicVSR_origin(64, 30).eval().cuda()
    x = torch.rand(2, 15, 3, 180, 320, generator=torch.Generator().manual_seed(1234))[:, :5]   # 5 frames keep the fp32 arm short
    xb = x.cuda()
    y2 = m.set_precision("bf16")(xb, 720, 1280)
    assert tuple(y2.shape) == (2, 5, 3, 720, 1280)
    y1 = m(xb[1:2], 720, 1280)
    assert torch.equal(y1, y2[1:2])
    yf = m.set_precision("fp32")(xb[:1], 720, 1280)
    assert port.psnr_db(y2[:1].cpu(), yf.cpu()) >= 50.0


@pytest.mark.parametrize("mt", ["1", "2"])
@pytest.mark.parametrize("n,h,w", [(1, 8, 30), (2, 19, 37), (1, 180, 320), (3, 33, 61)])
@pytest.mark.parametrize("act,with_res", [(0, True), (1, False), (2, True)])
def test_conv3x3_c64_tcgen05(V, n, h, w, act, with_res, mt, monkeypatch):
    """The tcgen05 form of the BasicVSR trunk convolution (3x3, 64 -> 64, bf16 NHWC; models/basicvsr_arch_origin.py:115-137):
    against torch fp32 on the bf16-rounded operands, and against the mma.sync kernel it replaces (B200SR_CONV_IMPL=mma).
    Exercises partial tiles in x and y, several images per launch and the image border (TMA zero fill = the conv's zero padding)."""
    monkeypatch.setenv("B200SR_CONV_MT", mt)      # both tile heights (30 x 4 and 30 x 8 outputs); the launcher picks by tiles per CTA
    g = torch.Generator().manual_seed(n * 1000 + h * 10 + w + act)
    conv = nn.Conv2d(64, 64, 3, 1, 1)
    with torch.no_grad():
        conv.weight.copy_(conv.weight.bfloat16().float())
    x = torch.randn(n, 64, h, w, generator=g).bfloat16()
    res = torch.randn(n, 64, h, w, generator=g).bfloat16()
    hd = V._ConvHandle(conv, torch.device("cuda:0"))
    xn = x.permute(0, 2, 3, 1).contiguous().cuda()
    rn = res.permute(0, 2, 3, 1).contiguous().cuda() if with_res else None
    with torch.no_grad():
        ref = F.conv2d(x.double(), conv.weight.double(), conv.bias.double(), padding=1)
        ref = F.relu(ref) if act == 1 else F.leaky_relu(ref, 0.1) if act == 2 else ref
        if with_res:
            ref = ref + res.double()
    y = hd(xn, "bf16", act, residual=rn)
    torch.cuda.synchronize()
    monkeypatch.setenv("B200SR_CONV_IMPL", "mma")
    y_mma = hd(xn, "bf16", act, residual=rn)
    torch.cuda.synchronize()
    yf = y.float().cpu().permute(0, 3, 1, 2).double()
    # fp32 accumulation of exact bf16 products, one rounding to bf16 at the store: half an ulp of the output magnitude
    tol = 2.0 ** -8 * ref.abs().clamp_min(1.0) + 1e-3
    assert bool(((yf - ref).abs() <= tol).all()), float((yf - ref).abs().max())
    assert float((y.float() - y_mma.float()).abs().max()) <= 2.0 ** -6 * float(ref.abs().max())


def test_conv3x3_c64_tcgen05_channel_windows(V):
    """x / y / residual as 16-byte aligned channel windows of wider NHWC tensors."""
    conv = nn.Conv2d(64, 64, 3, 1, 1)
    g = torch.Generator().manual_seed(77)
    xw = torch.randn(1, 21, 45, 144, generator=g).bfloat16()
    rw = torch.randn(1, 21, 45, 72, generator=g).bfloat16()
    hd = V._ConvHandle(conv, torch.device("cuda:0"))
    out = torch.full((1, 21, 45, 80), 3.0, dtype=torch.bfloat16, device="cuda")
    import ctypes
    from mobilesuperresolution_b200 import _lib
    L = _lib.lib()
    xd, rd = xw.cuda(), rw.cuda()
    _lib.check(L.b200sr_conv_forward(hd._h, ctypes.c_void_p(xd.data_ptr()), 144, 72, ctypes.c_void_p(out.data_ptr()), 80, 8,
                                     ctypes.c_void_p(rd.data_ptr()), 72, 8, 1, 21, 45, 0, 1, _lib.BF16, _lib.BF16, _lib.precision_code("bf16"),
                                     _lib.current_stream_ptr(xd.device)))
    torch.cuda.synchronize()
    with torch.no_grad():
        ref = F.conv2d(xw[..., 72:136].permute(0, 3, 1, 2).float(), conv.weight.bfloat16().float(), conv.bias, padding=1) + \
            rw[..., 8:72].permute(0, 3, 1, 2).float()
    o = out.float().cpu()
    assert float((o[..., 8:72].permute(0, 3, 1, 2) - ref).abs().max()) <= 0.05
    assert float((o[..., :8] - 3).abs().max()) == 0 and float((o[..., 72:] - 3).abs().max()) == 0


def _to_planar8(t):   # (n,h,w,64) -> (n,8,h,w,8)
    n, h, w, _ = t.shape
    return t.view(n, h, w, 8, 8).permute(0, 3, 1, 2, 4).contiguous()


def _from_planar8(t):
    n, _, h, w, _ = t.shape
    return t.permute(0, 2, 3, 1, 4).reshape(n, h, w, 64)


@pytest.mark.parametrize("xp,yp", [(True, True), (True, False), (False, True)])
def test_conv3x3_c64_tcgen05_planar8(V, xp, yp):
    """Planar-8 [n][c/8][h][w][8] input (+ residual) / output of the tcgen05 conv: bit-identical to its NHWC form."""
    g = torch.Generator().manual_seed(5)
    conv = nn.Conv2d(64, 64, 3, 1, 1)
    hd = V._ConvHandle(conv, torch.device("cuda:0"))
    assert hd.tcgen05_ok()
    x = torch.randn(2, 27, 65, 64, generator=g).bfloat16().cuda()
    r = torch.randn(2, 27, 65, 64, generator=g).bfloat16().cuda()
    ref = hd(x, "bf16", V.ACT_LRELU, residual=r)
    y = hd(_to_planar8(x) if xp else x, "bf16", V.ACT_LRELU, residual=_to_planar8(r) if xp else r, x_planar=xp, y_planar=yp)
    y = _from_planar8(y) if yp else y
    assert torch.equal(y, ref)


def test_conv3x3_c67_tcgen05_first_trunk_conv(V, monkeypatch):
    """The trunk's first conv (67 -> 64 on [x_i | warped features | pad], models/basicvsr_arch_origin.py:69-70,104): channels past
    cin are clipped by the tensor map, so whatever the pad channels hold (NaN here) never reaches the MMA."""
    g = torch.Generator().manual_seed(9)
    conv = nn.Conv2d(67, 64, 3, 1, 1)
    hd = V._ConvHandle(conv, torch.device("cuda:0"))
    assert hd.tcgen05_ok()
    x = torch.randn(2, 23, 41, 80, generator=g).bfloat16()
    x[..., 67:] = float("nan")
    xd = x.cuda()
    y = hd(xd, "bf16", V.ACT_LRELU)
    yp = _from_planar8(hd(xd, "bf16", V.ACT_LRELU, y_planar=True))
    with torch.no_grad():
        ref = F.leaky_relu(F.conv2d(x[..., :67].permute(0, 3, 1, 2).double(), conv.weight.bfloat16().double(), conv.bias.double(), padding=1), 0.1)
    yf = y.float().cpu().permute(0, 3, 1, 2).double()
    tol = 2.0 ** -8 * ref.abs().clamp_min(1.0) + 1e-3
    assert bool(((yf - ref).abs() <= tol).all()), float((yf - ref).abs().max())
    assert torch.equal(y, yp)


def test_conv_planar8_needs_tcgen05(V, monkeypatch):
    """Planar-8 is served by the tcgen05 kernel only: asking the other kernels for it fails loudly (no silent re-layout)."""
    conv = nn.Conv2d(64, 64, 3, 1, 1)
    hd = V._ConvHandle(conv, torch.device("cuda:0"))
    x = torch.zeros(1, 8, 9, 9, 8, dtype=torch.bfloat16, device="cuda")
    monkeypatch.setenv("B200SR_CONV_IMPL", "mma")
    assert not hd.tcgen05_ok()
    with pytest.raises(RuntimeError):
        hd(x, "bf16", V.ACT_NONE, x_planar=True, y_planar=True)


@pytest.mark.parametrize("cout,shuffle", [(256, 2), (256, 1), (128, 1)])
def test_conv3x3_tcgen05_output_channel_groups(V, cout, shuffle, monkeypatch):
    """upconv1 / upconv2 (64 -> 256 + PixelShuffle(2) + LeakyReLU, models/basicvsr_arch_origin.py:87-88): one CTA per group of 64
    output channels, the shuffle folded into the store.  Against torch fp64 on the bf16 operands and the mma.sync kernel."""
    g = torch.Generator().manual_seed(cout + shuffle)
    conv = nn.Conv2d(64, cout, 3, 1, 1)
    hd = V._ConvHandle(conv, torch.device("cuda:0"))
    assert hd.tcgen05_ok()
    x = torch.randn(2, 21, 47, 64, generator=g).bfloat16()
    res = torch.randn(2, 21, 47, cout, generator=g).bfloat16() if shuffle == 1 else None
    with torch.no_grad():
        ref = F.leaky_relu(F.conv2d(x.permute(0, 3, 1, 2).double(), conv.weight.bfloat16().double(), conv.bias.double(), padding=1), 0.1)
        ref = F.pixel_shuffle(ref, 2) if shuffle == 2 else ref + res.permute(0, 3, 1, 2).double()
    xd, rd = x.cuda(), (res.cuda() if res is not None else None)
    y = hd(xd, "bf16", V.ACT_LRELU, shuffle=shuffle, residual=rd)
    torch.cuda.synchronize()
    monkeypatch.setenv("B200SR_CONV_IMPL", "mma")
    y_mma = hd(xd, "bf16", V.ACT_LRELU, shuffle=shuffle, residual=rd)
    torch.cuda.synchronize()
    yf = y.float().cpu().permute(0, 3, 1, 2).double()
    tol = 2.0 ** -8 * ref.abs().clamp_min(1.0) + 1e-3
    assert bool(((yf - ref).abs() <= tol).all()), float((yf - ref).abs().max())
    assert float((y.float() - y_mma.float()).abs().max()) <= 2.0 ** -6 * float(ref.abs().max())


def _to_planar(t):   # (n,h,w,c) -> (n,c/8,h,w,8)
    n, h, w, c = t.shape
    return t.view(n, h, w, c // 8, 8).permute(0, 3, 1, 2, 4).contiguous()


def _from_planar(t):
    n, q, h, w, _ = t.shape
    return t.permute(0, 2, 3, 1, 4).reshape(n, h, w, q * 8)


@pytest.mark.parametrize("cin,cout", [(8, 32), (32, 64), (64, 32), (32, 16)])
@pytest.mark.parametrize("n,h,w", [(1, 6, 10), (2, 31, 53), (1, 192, 320)])
def test_conv7x7_tcgen05_spynet_layers(V, cin, cout, n, h, w, monkeypatch):
    """SPyNet BasicModule layers (7x7, models/spynet_arch.py:17-22) on the tcgen05 kernel with streamed filter tap rows: against torch
    fp64 on the bf16 operands and the mma.sync kernel; NHWC and planar-8 forms agree bit for bit.  6x10 is SPyNet's coarsest level
    (the whole image is halo), 31x53 has partial tiles in both directions, 192x320 is cfg4's finest level (several tiles per CTA)."""
    g = torch.Generator().manual_seed(cin * 7 + cout + h)
    conv = nn.Conv2d(cin, cout, 7, 1, 3)
    hd = V._ConvHandle(conv, torch.device("cuda:0"))
    assert hd.tcgen05_ok()
    cs = max(cin, 16)   # the level input carries 8 channels in 16-channel pixels
    x = torch.randn(n, h, w, cs, generator=g).bfloat16()
    with torch.no_grad():
        ref = F.relu(F.conv2d(x[..., :cin].permute(0, 3, 1, 2).double(), conv.weight.bfloat16().double(), conv.bias.double(), padding=3))
    xd = x.cuda()
    y = hd(xd, "bf16", V.ACT_RELU)
    torch.cuda.synchronize()
    if cin % 16 == 0:
        yp = hd(_to_planar(xd), "bf16", V.ACT_RELU, x_planar=True, y_planar=True)
        assert torch.equal(_from_planar(yp), y)
    monkeypatch.setenv("B200SR_CONV_IMPL", "mma")
    y_mma = hd(xd, "bf16", V.ACT_RELU)
    torch.cuda.synchronize()
    yf = y.float().cpu().permute(0, 3, 1, 2).double()
    tol = 2.0 ** -8 * ref.abs().clamp_min(1.0) + 2e-3
    assert bool(((yf - ref).abs() <= tol).all()), float((yf - ref).abs().max())
    assert float((y.float() - y_mma.float()).abs().max()) <= 2.0 ** -6 * max(1.0, float(ref.abs().max()))


def test_basicvsr_origin_cuda_graph_two_streams(V):
    """The whole clip forward (SPyNet batch, two propagation directions forked onto two streams, reconstruction) captures into ONE
    CUDA graph and replays to the same result as the eager call -- no host sync, no allocation outside the capture pool."""
    torch.manual_seed(3)
    m = V.BasicVSR_origin(64, 2).cuda().eval().set_precision("bf16")
    x = torch.rand(1, 4, 3, 64, 96, device="cuda")
    with torch.no_grad():
        eager = m(x, 256, 384).clone()
        st = torch.cuda.Stream()
        st.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(st):
            m(x, 256, 384)
            st.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=st):
                y = m(x, 256, 384)
            y.zero_()
            g.replay()
            st.synchronize()
        assert torch.equal(y, eager)
        x.copy_(torch.rand_like(x))            # new frames in the captured input buffer
        g.replay()
        torch.cuda.synchronize()
        assert float((y - m(x, 256, 384)).abs().max()) == 0.0


def _mvvsr_case(name):
    from oracle import synth
    meta, arrs = load_golden(name)
    sd = {k: torch.from_numpy(v) for k, v in synth.synth_state_dict(meta["shapes"], meta["seed"]).items()}
    return meta, torch.from_numpy(arrs["y"]), sd, torch.from_numpy(synth.synth_mv_clip(meta["shape"], meta["input_seed"]))


@pytest.mark.parametrize("name", ["mvvsr_nf64", "mvvsr_nf16"])
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_mvvsr_golden(V, name, precision):
    """MotionVectorVSR (models/mvvsr_arch.py:56-109) against the reference's output: codec motion vectors as flows, BasicVSR trunks,
    ConvTranspose2d(stride 4) tail as 3x3 conv + PixelShuffle(4), bilinear resize + base (size = x4 and not x4)."""
    from oracle import port
    meta, ref, sd, x = _mvvsr_case(name)
    m = V.MotionVectorVSR(meta["num_feat"], meta["num_block"]).eval()
    m.load_state_dict(sd, strict=True)
    m = m.cuda().set_precision(precision)
    with torch.no_grad():
        y = m(x.cuda(), *meta["size"]).cpu()
    assert y.shape == ref.shape
    if precision == "fp32":
        assert float((y - ref).abs().max()) <= 1e-4, float((y - ref).abs().max())
    else:
        assert port.psnr_db(y, ref) >= 50.0, port.psnr_db(y, ref)


def test_conv3x3_tcgen05_shuffle_store_planar8(V):
    """upconv2's PixelShuffle(2) store straight into the planar-8 layout conv_hr reads: same values as the NHWC store."""
    g = torch.Generator().manual_seed(12)
    conv = nn.Conv2d(64, 256, 3, 1, 1)
    hd = V._ConvHandle(conv, torch.device("cuda:0"))
    x = torch.randn(2, 19, 43, 64, generator=g).bfloat16().cuda()
    ref = hd(x, "bf16", V.ACT_LRELU, shuffle=2)                       # (2, 38, 86, 64)
    yp = hd(x, "bf16", V.ACT_LRELU, shuffle=2, y_planar=True)         # (2, 8, 38, 86, 8)
    assert tuple(yp.shape) == (2, 8, 38, 86, 8)
    assert torch.equal(_from_planar(yp), ref)


@pytest.mark.parametrize("planar", [False, True])
def test_conv_last_plus_base_fused_tcgen05(V, planar):
    """conv_last (3x3, 64 -> 3) + F.interpolate(x_i, scale_factor=4, bilinear, align_corners=False) in one launch
    (models/basicvsr_arch_origin.py:90-92): fp32 NCHW result written into a strided batch slot, against torch fp64."""
    import ctypes
    from mobilesuperresolution_b200 import _lib
    g = torch.Generator().manual_seed(44)
    conv = nn.Conv2d(64, 3, 3, 1, 1)
    hd = V._ConvHandle(conv, torch.device("cuda:0"))
    b, h, w = 2, 9, 13
    x = torch.randn(b, 4 * h, 4 * w, 64, generator=g).bfloat16()
    lr = torch.rand(b, 3, 3, h, w, generator=g)                         # (b, n, 3, h, w) clip: frame 1 is the base
    with torch.no_grad():
        ref = F.conv2d(x.permute(0, 3, 1, 2).double(), conv.weight.bfloat16().double(), conv.bias.double(), padding=1) + \
            F.interpolate(lr[:, 1].double(), scale_factor=4, mode="bilinear", align_corners=False)
    xd = x.cuda()
    xin = _to_planar(xd) if planar else xd
    lrd = lr.cuda()
    out = torch.full((b, 3, 3, 4 * h, 4 * w), 9.0, device="cuda")
    base = lrd[:, 1]
    slot = out[:, 2]
    _lib.check(_lib.lib().b200sr_vsr_conv_last_base(hd._h, ctypes.c_void_p(xin.data_ptr()), int(planar), 64, 0, ctypes.c_void_p(base.data_ptr()),
                                                    lrd.stride(0), ctypes.c_void_p(slot.data_ptr()), out.stride(0), b, 4 * h, 4 * w,
                                                    _lib.current_stream_ptr(xd.device)))
    torch.cuda.synchronize()
    o = out.cpu()
    assert float((o[:, 2].double() - ref).abs().max()) <= 2e-4 * max(1.0, float(ref.abs().max()))
    assert float((o[:, :2] - 9).abs().max()) == 0


def test_graphed_helper_wdsr_and_clip(V):
    """mobilesuperresolution_b200.Graphed: a forward replayed as one CUDA graph returns the eager result for new inputs."""
    import types
    import mobilesuperresolution_b200 as sr
    torch.manual_seed(5)
    p = types.SimpleNamespace(image_mean=0.5, num_channels=3, scale=4, num_blocks=3, num_residual_units=24, width_search=False, pretrained=False)
    m = sr.BASIC_MODEL(p).cuda().eval().set_precision("bf16")
    x = torch.rand(2, 3, 48, 64, device="cuda").bfloat16()
    with torch.no_grad():
        g = sr.Graphed(m, x)
        x2 = torch.rand_like(x.float()).bfloat16()
        assert torch.equal(g(x2), m(x2))
        vsr = V.BasicVSR_origin(64, 1).cuda().eval().set_precision("bf16")
        clip = torch.rand(1, 3, 3, 64, 64, device="cuda")
        gv = sr.Graphed(vsr, clip, 256, 256)
        clip2 = torch.rand_like(clip)
        y = gv(clip2, clone=True)
        assert torch.equal(y, vsr(clip2, 256, 256))
        with pytest.raises(RuntimeError):
            gv(torch.rand(1, 2, 3, 64, 64, device="cuda"))


@pytest.mark.parametrize("cout", [64, 128])
@pytest.mark.parametrize("n,h,w", [(1, 5, 9), (2, 37, 70), (1, 180, 320)])
def test_conv1x1_c128_tcgen05_fusion(V, cout, n, h, w, monkeypatch):
    """The fusion conv (1x1, 2 nf -> nf | 2 nf on cat([backward, forward]), models/basicvsr_arch_origin.py:84, mvvsr_arch.py:33) on the 1x1
    form of the tcgen05 kernel: against torch fp64 on the bf16 operands and the mma.sync kernel."""
    g = torch.Generator().manual_seed(cout + h)
    conv = nn.Conv2d(128, cout, 1, 1, 0)
    hd = V._ConvHandle(conv, torch.device("cuda:0"))
    assert hd.tcgen05_ok()
    x = torch.randn(n, h, w, 128, generator=g).bfloat16()
    with torch.no_grad():
        ref = F.leaky_relu(F.conv2d(x.permute(0, 3, 1, 2).double(), conv.weight.bfloat16().double(), conv.bias.double()), 0.1)
    xd = x.cuda()
    y = hd(xd, "bf16", V.ACT_LRELU)
    torch.cuda.synchronize()
    monkeypatch.setenv("B200SR_CONV_IMPL", "mma")
    y_mma = hd(xd, "bf16", V.ACT_LRELU)
    torch.cuda.synchronize()
    yf = y.float().cpu().permute(0, 3, 1, 2).double()
    tol = 2.0 ** -8 * ref.abs().clamp_min(1.0) + 1e-3
    assert bool(((yf - ref).abs() <= tol).all()), float((yf - ref).abs().max())
    assert float((y.float() - y_mma.float()).abs().max()) <= 2.0 ** -6 * max(1.0, float(ref.abs().max()))
