"""GPU parity added in round 2: the fork's NAS_MODEL as committed, the fork BasicVSR.forward (num_feat = 3), the nf = 64 BasicVSR_origin
chain end to end, cfg4-size frames against the oracle port, SPyNet bf16 in dB at the cfg4 frame size, and the row-streaming form of
the fused block against the tile form.  Gates as everywhere: fp32 raw max-abs <= 1e-4, bf16 PSNR >= 50 dB."""
import os
import types

import numpy as np
import pytest
import torch

from conftest import golden_case, load_golden

pytestmark = pytest.mark.gpu

FP32_TOL, BF16_PSNR = 1e-4, 50.0


def _t(sd):
    return {k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}


@pytest.fixture(scope="module")
def sr():
    import mobilesuperresolution_b200 as m
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return m


def _nas_params(scale, nb, ws=True):
    return types.SimpleNamespace(image_mean=0.5, num_channels=3, scale=scale, num_blocks=nb, num_residual_units=24, width_search=ws,
                                 pretrained=False)


# ------------------------------------------------------------------------------------------------ fork NAS_MODEL
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_fork_nas_model_golden(sr, precision):
    """models/wdsr_b.py:30-137 as committed: reference state_dict loads strict=True (the estimator MLP keeps its own init: it never
    enters the forward), one block gated off, 12 of 24 trunk channels masked; (sr, speed_accu) against the reference."""
    from oracle import port
    meta, arrs = load_golden("nas_fork")
    m = sr.NAS_MODEL(_nas_params(meta["scale"], meta["nb"])).eval()
    sd = {k[3:]: torch.from_numpy(v) for k, v in arrs.items() if k.startswith("sd.")}
    full = dict(m.state_dict())
    assert set(sd) == {k for k in full if not k.startswith("speed_estimator.")}
    full.update(sd)
    m.load_state_dict(full, strict=True)
    assert m.get_block_status() == meta["block_status"] and m.get_width_from_block_idx(m.get_block_status()) == meta["widths"]
    x = torch.from_numpy(np.asarray(golden_case("nas_fork")[3]))
    m = m.cuda().set_precision(precision)
    xd = x.cuda() if precision == "fp32" else x.cuda().bfloat16()
    with torch.no_grad():
        out, speed = m(xd)
    ref = torch.from_numpy(arrs["y"])
    assert torch.equal(speed.cpu(), torch.from_numpy(arrs["speed"]))            # scalar host arithmetic in the reference's order
    if precision == "fp32":
        assert float((out.float().cpu() - ref).abs().max()) <= FP32_TOL
    else:
        assert port.psnr_db(out.float().cpu(), ref) >= BF16_PSNR
    assert m.launches_per_forward() == 2 + len(meta["block_status"]) + 2          # head, 2 layout changes, kept blocks, tail


def test_fork_nas_model_without_width_search_raises_like_the_reference(sr):
    m = sr.NAS_MODEL(_nas_params(2, 1, ws=False)).eval().cuda()
    with pytest.raises(AttributeError):
        m(torch.rand(1, 3, 8, 8, device="cuda"))


def test_fork_nas_model_all_blocks_skipped_and_odd_shapes(sr):
    from oracle import port, synth
    m = sr.NAS_MODEL(_nas_params(4, 2)).eval()
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    sd = _t(synth.synth_state_dict(shapes, 33))
    for i in range(2):
        sd[f"body.{i}.alpha1"], sd[f"body.{i}.alpha2"] = torch.tensor([0.9]), torch.tensor([0.1])
    m.load_state_dict(sd)
    x = torch.from_numpy(synth.synth_input((1, 3, 13, 37), 34))
    ref, rspeed = port.nas_fork_forward(sd, x, 4)
    with torch.no_grad():
        out, speed = m.cuda()(x.cuda())
    assert float((out.cpu() - ref).abs().max()) <= FP32_TOL and torch.equal(speed.cpu(), rspeed)
    sd["body.1.alpha1"], sd["body.1.alpha2"] = torch.tensor([0.1]), torch.tensor([0.9])       # mutate: the plan must follow
    m.load_state_dict(sd)
    ref, _ = port.nas_fork_forward(sd, x, 4)
    with torch.no_grad():
        out, _ = m(x.cuda())
    assert float((out.cpu() - ref).abs().max()) <= FP32_TOL


@pytest.mark.parametrize("shape", [(2, 3, 24, 40), (1, 3, 13, 37), (1, 3, 10, 12)])
def test_fork_nas_model_bf16_layout_and_split_arms(sr, shape):
    """bf16 forward of the fork NAS_MODEL on shapes that take the vectorised planar-8 <-> NCHW conversion and the tensor-core Split_Block
    arm (W % 8 == 0), the generic conversion with the FFMA arm (13 x 37), and the vectorised conversion with the FFMA arm (10 x 12:
    H * W % 8 == 0, W % 8 != 0): >= 50 dB against the oracle port on the bf16-rounded input."""
    from oracle import port, synth
    m = sr.NAS_MODEL(_nas_params(4, 3)).eval()
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    sd = _t(synth.synth_state_dict(shapes, 35))
    m.load_state_dict(sd)
    x = torch.from_numpy(synth.synth_input(shape, 36)).bfloat16()
    ref, _ = port.nas_fork_forward(sd, x.float(), 4)
    with torch.no_grad():
        out, _ = m.cuda().set_precision("bf16")(x.cuda())
    assert port.psnr_db(out.float().cpu(), ref) >= BF16_PSNR


# ------------------------------------------------------------------------------------------------ video
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_fork_basicvsr_forward_nf3_golden(sr, precision):
    """models/basicvsr_arch.py:56-105 runs as committed only for num_feat == 3: SPyNet + propagation + ConvTranspose2d(6, 3, 5, stride 4)
    tail + resize + base."""
    from oracle import port
    meta, arrs, sd, x = golden_case("basicvsr_fork_nf3")
    m = sr.BasicVSR(meta["num_feat"], meta["num_block"]).eval()
    m.load_state_dict(_t(sd))
    m = m.cuda().set_precision(precision)
    h, w = meta["out_hw"]
    with torch.no_grad():
        y = m(torch.from_numpy(x).cuda(), h, w).cpu()
    assert tuple(y.shape) == (1, 3, 3, h, w)
    s = meta["stride"]
    ref_s, ref_c = torch.from_numpy(arrs["y_strided"]), torch.from_numpy(arrs["y_corner"])
    if precision == "fp32":
        assert float((y[..., ::s, ::s] - ref_s).abs().max()) <= FP32_TOL
        assert float((y[..., -16:, -16:] - ref_c).abs().max()) <= FP32_TOL
    else:
        assert port.psnr_db(y[..., ::s, ::s], ref_s) >= BF16_PSNR


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_basicvsr_origin_nf64_golden(sr, precision):
    """BasicVSR_origin(64, 2): the nf = 64 chain end to end against the reference -- tcgen05 trunks 67 -> 64 / 64 -> 64, fusion on the
    1x1 form, upconv1/2 with the PixelShuffle(2) store, planar conv_hr, fused conv_last + base (bf16), the FFMA arm (fp32)."""
    from oracle import port
    meta, arrs, sd, x = golden_case("basicvsr_origin_nf64")
    m = sr.BasicVSR_origin(meta["num_feat"], meta["num_block"]).eval()
    m.load_state_dict(_t(sd))
    m = m.cuda().set_precision(precision)
    h, w = meta["out_hw"]
    with torch.no_grad():
        y = m(torch.from_numpy(x).cuda(), h, w).cpu()
    s = meta["stride"]
    ref_s, ref_c = torch.from_numpy(arrs["y_strided"]), torch.from_numpy(arrs["y_corner"])
    if precision == "fp32":
        assert float((y[..., ::s, ::s] - ref_s).abs().max()) <= FP32_TOL
        assert float((y[..., :16, :16] - ref_c).abs().max()) <= FP32_TOL
    else:
        assert port.psnr_db(y[..., ::s, ::s], ref_s) >= BF16_PSNR


def test_cfg4_size_frames_against_the_oracle(sr):
    """BASELINE cfg4 as stated -- BasicVSR_origin(64, 30), 180x320 frames -> 720x1280 -- on 3 frames against the oracle port (the CPU
    forward of the full 15 frames takes ~40 s; the recurrence is the same code for any clip length): fp32 <= 1e-4, bf16 >= 50 dB."""
    from oracle import port, synth
    m = sr.BasicVSR_origin(64, 30).eval()
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    sd = _t(synth.synth_state_dict(shapes, 91))
    for k in sd:                                       # 60 residual convs deep: keep the synthetic trunk contractive
        if ".main.2." in k and k.endswith("weight"):
            sd[k] = sd[k] * 0.25
    m.load_state_dict(sd)
    x = torch.from_numpy(synth.synth_input((1, 3, 3, 180, 320), 92))
    ref = port.basicvsr_origin_forward(sd, x, 720, 1280)
    m = m.cuda()
    with torch.no_grad():
        yf = m.set_precision("fp32")(x.cuda(), 720, 1280).cpu()
        yb = m.set_precision("bf16")(x.cuda(), 720, 1280).cpu()
    assert float((yf - ref).abs().max()) <= FP32_TOL
    assert port.psnr_db(yb, ref) >= BF16_PSNR


def test_spynet_bf16_in_db_at_cfg4_size(sr):
    """SPyNet at 180x320 with flows up to 8.5 px: fp32 <= 1e-4 on the flow, bf16 convolutions >= 50 dB against the reference flow
    (peak = the reference flow's dynamic range) -- the same gate as every other bf16 arm, instead of a pixel tolerance."""
    from oracle import port, synth
    meta, arrs = load_golden("spynet_180x320")
    sp = sr.SpyNet().eval()
    sp.load_state_dict(_t(synth.synth_state_dict(meta["shapes"], meta["wseed"])))
    a = torch.from_numpy(synth.synth_input(meta["shape"], meta["aseed"])).cuda()
    b = torch.from_numpy(synth.synth_input(meta["shape"], meta["bseed"])).cuda()
    ref = torch.from_numpy(arrs["f_strided"])
    s = meta["stride"]
    sp = sp.cuda()
    with torch.no_grad():
        f32 = sp.set_precision("fp32")(a, b).cpu()
        f16 = sp.set_precision("bf16")(a, b).cpu()
    assert float((f32[..., ::s, ::s] - ref).abs().max()) <= FP32_TOL
    assert port.psnr_db(f16[..., ::s, ::s], ref) >= BF16_PSNR


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_naive_model_golden(sr, precision, tmp_path):
    """Naive_model (models/naive_multi_model_easy.py:110-147): SPyNet flows, warp of the previous frame's encoded features into the first
    block's input window, conv-ReLU-conv residual blocks (kernel sizes 3, 3, 5), 5x5 decode + PixelShuffle(4) + bilinear base -- against the
    vectors the unmodified reference produced (oracle/make_golden_r3.py)."""
    from oracle import port
    meta, arrs, sd, x = golden_case("naive_model")
    f = tmp_path / "naive_index.txt"
    f.write_text(repr((list(range(len(meta["blocks"]))), meta["blocks"])) + "\n")
    m = sr.Naive_model(meta["scale"], str(f)).eval()
    m.load_state_dict(_t(sd), strict=True)
    m = m.cuda().set_precision(precision)
    with torch.no_grad():
        y = m(torch.from_numpy(x).cuda()).cpu()
    assert tuple(y.shape) == (1, 3, 3, 144, 208)
    s = meta["stride"]
    ref_s, ref_c = torch.from_numpy(arrs["y_strided"]), torch.from_numpy(arrs["y_corner"])
    if precision == "fp32":
        assert float((y[..., ::s, ::s] - ref_s).abs().max()) <= FP32_TOL * max(1.0, float(ref_s.abs().max()))
        assert float((y[..., :16, :16] - ref_c).abs().max()) <= FP32_TOL * max(1.0, float(ref_s.abs().max()))
    else:
        assert port.psnr_db(y[..., ::s, ::s], ref_s, peak=float(ref_s.max() - ref_s.min())) >= BF16_PSNR
    # frame 0 alone takes the zero-flow / own-features branch; a width the warp kernel's window form does not serve takes the NCHW kernel
    with torch.no_grad():
        y0 = m(torch.from_numpy(x[:, :1]).cuda()).cpu()
    assert torch.equal(y0[:, 0], y[:, 0])
    with pytest.raises(RuntimeError):
        sr.Naive_model(2, str(f)).eval().cuda()(torch.from_numpy(x).cuda())     # x4 base on a x2 model: the reference's shape error


@pytest.mark.parametrize("IN,ks,precision", [(12, (3, 7), "fp32"), (8, (5,), "bf16"), (24, (3, 3), "bf16")])
def test_naive_model_widths_and_batches_against_the_oracle(sr, IN, ks, precision, tmp_path):
    """Naive_model at widths the golden does not cover, batch 2, against the oracle port on the same seeded weights: IN = 12 is not a
    multiple of the warp kernel's 16-byte channel slices in bf16 / takes the 3-slice window form in fp32, k = 7 blocks, a single 5x5 block
    (decode inherits 5)."""
    from oracle import port, synth
    blocks = [[IN, 0, k] for k in ks]
    f = tmp_path / "naive_index.txt"
    f.write_text(repr((list(range(len(blocks))), blocks)) + "\n")
    m = sr.Naive_model(4, str(f)).eval()
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    sd = _t(synth.synth_state_dict(shapes, 500 + IN))
    m.load_state_dict(sd, strict=True)
    m = m.cuda().set_precision(precision)
    x = torch.from_numpy(synth.synth_input((2, 3, 3, 40, 72), 501))
    ref = port.naive_model_forward(sd, x, len(blocks))
    with torch.no_grad():
        y = m(x.cuda()).cpu()
    if precision == "fp32":
        assert float((y - ref).abs().max()) <= FP32_TOL * max(1.0, float(ref.abs().max()))
    else:
        assert port.psnr_db(y, ref) >= BF16_PSNR


def test_naive_model_graph_replay(sr, tmp_path):
    """Naive_model's forward (SPyNet on its cached workspace, windowed warp, convs, tail) is CUDA-graph capturable and replays bit-identically."""
    f = tmp_path / "naive_index.txt"
    f.write_text(repr(([0, 1], [[16, 0, 3], [16, 0, 3]])) + "\n")
    torch.manual_seed(21)
    m = sr.Naive_model(4, str(f)).eval().cuda().set_precision("bf16")
    x = torch.rand(1, 3, 3, 64, 96, device="cuda")
    with torch.no_grad():
        y = m(x).clone()
        g = sr.Graphed(m, x)
        assert torch.equal(g(x), y) and torch.equal(g(x), y)


# ------------------------------------------------------------------------------------------------ row-streaming block
@pytest.mark.parametrize("shape", [(1, 24, 16, 32), (2, 24, 37, 45), (3, 24, 96, 96), (1, 24, 5, 300), (2, 24, 1, 1), (1, 24, 131, 7)])
@pytest.mark.parametrize("widths", [(24, 144, 20), (24, 144, 24), (20, 100, 13), (9, 91, 7)])
@pytest.mark.parametrize("form", ["rs", "rh", "tc5q"])
def test_row_streaming_block_equals_oracle_and_tile_form(sr, shape, widths, form, monkeypatch):
    """The row-streaming forms of the fused block (csrc/wdsr_rs.cuh: all tcgen05; csrc/wdsr_rh.cuh: reduce 1x1 on mma.sync out of
    registers; B200SR_BLOCK_IMPL=rs|rh) on ragged shapes and pruned widths:
    >= 50 dB against the oracle block, and within one bf16 ulp-flip budget of the tile form (same rounding points, different fp32
    summation order)."""
    import tempfile
    from oracle import port, synth
    c, m1, m2 = widths
    n, _, h, w = shape

    def build(impl):
        monkeypatch.setenv("B200SR_BLOCK_IMPL", impl)
        f = tempfile.NamedTemporaryFile("w", suffix="_block_index.txt", delete=False)
        f.write(repr(([0], [[c, m1, m2]])) + "\n")
        f.close()
        m = sr.Model(2, f.name).eval()
        os.unlink(f.name)
        shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        sd = _t(synth.synth_state_dict(shapes, 300 + m1))
        m.load_state_dict(sd)
        return m.cuda().set_precision("bf16"), sd

    g = torch.Generator().manual_seed(h * 1000 + w)
    t = (torch.rand(n, c, h, w, generator=g) - 0.5) * 2
    outs = {}
    for impl in (form, "tc5"):
        m, sd = build(impl)
        plan = m.prepare()
        cp = plan.trunk_channels
        tin = torch.zeros(n, h, w, cp)
        tin[..., :c] = t.permute(0, 2, 3, 1)
        tin = tin.cuda().bfloat16()
        with torch.no_grad():
            o = plan.block(0, tin, "bf16")
        torch.cuda.synchronize()
        outs[impl] = o.float().cpu()
    ref = port.wdsr_block(sd, "body.1.", tin.float().cpu()[..., :c].permute(0, 3, 1, 2))
    got = outs[form][..., :c].permute(0, 3, 1, 2)
    assert port.psnr_db(got, ref) >= BF16_PSNR
    assert float(outs[form][..., c:].abs().max()) == 0.0 if cp > c else True            # pad channels stay exactly zero
    d = (outs[form] - outs["tc5"]).abs()
    assert float(d.max()) <= 2 ** -6 * max(1.0, float(ref.abs().max()))                  # <= a couple of bf16 ulps at the output's scale
    assert float((d > 0).float().mean()) < 0.2


@pytest.mark.parametrize("form", ["rs", "rh", "tc5q", "chain"])
def test_row_streaming_model_graph_replay_is_deterministic(sr, form, monkeypatch):
    monkeypatch.setenv("B200SR_BLOCK_IMPL", form)
    torch.manual_seed(3)
    p = types.SimpleNamespace(image_mean=0.5, num_channels=3, scale=4, num_blocks=4, num_residual_units=24, width_search=False, pretrained=False)
    m = sr.BASIC_MODEL(p).eval().cuda().set_precision("bf16")
    x = torch.rand(3, 3, 70, 90, device="cuda").bfloat16()
    with torch.no_grad():
        ref = m(x).clone()
        g = sr.Graphed(m, x)
        for _ in range(3):
            assert torch.equal(g(x), ref)


@pytest.mark.parametrize("nb,scale,shape", [(1, 4, (1, 3, 16, 32)), (2, 2, (2, 3, 37, 45)), (5, 2, (1, 3, 131, 200)), (16, 4, (4, 3, 96, 96))])
def test_chained_block_launch_matches_the_per_block_forward(sr, nb, scale, shape, monkeypatch):
    """All residual blocks in one persistent cooperative launch with a grid-wide layer barrier (csrc/wdsr_tc5c.cuh, B200SR_BLOCK_IMPL=chain):
    the same tiles through the same pipeline as the one-launch-per-block forward -- odd and even block counts (the result lands in either
    ping-pong buffer), one-tile-per-CTA grids (every tile is a layer boundary), eager and graph replay (bit-identical to itself).  Against
    the per-block forward: equal up to the fp32 summation order of the 3x3 (the tile form packs its K into 12 MMAs per M-tile, the chained
    kernel keeps the 27-slice form), i.e. a couple of bf16 ulps after each block."""
    outs = {}
    for impl in ("tc5", "chain"):
        monkeypatch.setenv("B200SR_BLOCK_IMPL", impl)
        torch.manual_seed(11)
        m = sr.BASIC_MODEL(_nas_params(scale, nb, ws=False)).eval().cuda().set_precision("bf16")
        torch.manual_seed(12)
        x = torch.rand(*shape, device="cuda").bfloat16()
        with torch.no_grad():
            y = m(x).clone()
            g = sr.Graphed(m, x)
            assert torch.equal(g(x), y) and torch.equal(g(x), y)
        outs[impl] = y
    a, b = outs["tc5"].float(), outs["chain"].float()
    d = (a - b).abs()
    from oracle import port
    assert float(d.max()) <= 2 ** -5 * max(1.0, float(a.abs().max())) * max(1, nb // 4)
    assert port.psnr_db(b, a) >= 55.0          # ulp flips accumulate over the blocks; a misplaced 3x3 slice would be ~20 dB


# ------------------------------------------------------------------------------------------------ our own memcheck / racecheck
# compute-sanitizer is closed on this GPU pool (profiles/r02_compute_sanitizer_closed.txt), so the two properties it would check are
# tested directly: (1) no kernel of the WDSR path writes outside its output / workspace -- every buffer sits between red zones of a
# canary pattern that must survive; (2) the mbarrier / TMEM pipelines are free of observable races -- the same input gives
# bit-identical output 12 times in a row while a second stream keeps the SMs and L2 busy with unrelated work.
@pytest.mark.parametrize("impl", ["tc5", "rs", "rh", "tc5q", "chain"])
@pytest.mark.parametrize("shape,scale", [((2, 3, 37, 45), 4), ((1, 3, 130, 66), 2), ((3, 3, 96, 96), 4)])
def test_red_zones_survive_and_results_repeat_under_load(sr, impl, shape, scale, monkeypatch):
    from mobilesuperresolution_b200 import _lib
    monkeypatch.setenv("B200SR_BLOCK_IMPL", impl)
    torch.manual_seed(5)
    p = types.SimpleNamespace(image_mean=0.5, num_channels=3, scale=scale, num_blocks=3, num_residual_units=24, width_search=False, pretrained=False)
    m = sr.BASIC_MODEL(p).eval().cuda().set_precision("bf16")
    plan = m.prepare()
    n, _, h, w = shape
    L = _lib.lib()
    need = L.b200sr_wdsr_workspace_bytes(plan.handle, n, h, w, _lib.BF16)
    RZ = 1 << 16
    ob = n * 3 * scale * h * scale * w * 2
    arena = torch.full((RZ + need + RZ + ob + RZ,), 0xA5, dtype=torch.uint8, device="cuda")
    ws = arena[RZ:RZ + need]
    out = arena[RZ + need + RZ:RZ + need + RZ + ob]
    x = torch.rand(*shape, device="cuda").bfloat16()
    side = torch.cuda.Stream()
    noise = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
    ref = None
    for it in range(12):
        with torch.cuda.stream(side):                     # unrelated traffic on another stream: perturbs CTA placement and timing
            for _ in range(1 + it % 3):
                noise.fill_(it)
        _lib.check(L.b200sr_wdsr_forward(plan.handle, x.data_ptr(), _lib.BF16, out.data_ptr(), _lib.BF16, n, h, w, _lib.BF16, ws.data_ptr(), need,
                                         _lib.current_stream_ptr(x.device)))
        torch.cuda.synchronize()
        y = out.clone()
        if ref is None:
            ref = y
        assert torch.equal(y, ref), f"iteration {it}: output changed"
    for a, b in [(0, RZ), (RZ + need, RZ + need + RZ), (RZ + need + RZ + ob, RZ + need + RZ + ob + RZ)]:
        assert bool((arena[a:b] == 0xA5).all()), "a kernel wrote outside its buffers"
    yv = ref.view(torch.bfloat16).view(n, 3, scale * h, scale * w)
    with torch.no_grad():
        assert torch.equal(yv, m(x))


def test_u8_frames_and_integer_psnr(sr):
    """8-bit frames in and out: to_tensor on the device, the tail's 8-bit store, and common/metrics.py:10-19 as an exact
    integer sum of squared differences."""
    torch.manual_seed(7)
    p = types.SimpleNamespace(image_mean=0.5, num_channels=3, scale=2, num_blocks=2, num_residual_units=24, width_search=False, pretrained=False)
    m = sr.BASIC_MODEL(p).eval().cuda().set_precision("bf16")
    x8 = torch.randint(0, 256, (2, 3, 33, 47), dtype=torch.uint8, device="cuda")
    xf = sr.u8_to_unit(x8)
    assert torch.equal(xf.cpu(), x8.cpu().float() / 255.0)                     # == torchvision to_tensor
    with torch.no_grad():
        y8 = sr.forward_u8_frames(m, x8)
        assert torch.equal(y8, m.forward_u8(sr.u8_to_unit(x8, torch.bfloat16)))
    hr8 = torch.randint(0, 256, tuple(y8.shape), dtype=torch.uint8, device="cuda")
    for shave in (0, 4, 6):
        ssd = sr.ssd_u8(y8, hr8, shave).cpu()
        a, b = y8.cpu().long(), hr8.cpu().long()
        if shave:
            a, b = a[..., shave:-shave, shave:-shave], b[..., shave:-shave, shave:-shave]
        assert torch.equal(ssd, ((a - b) ** 2).sum(dim=(1, 2, 3)))              # exact
        # the reference's formula on the float tensors it would have seen
        srf, hrf = y8.cpu().float() / 255.0, hr8.cpu().float() / 255.0
        q = (srf * 255).round().clamp(0, 255) / 255
        d = q.clamp(0, 1) - hrf
        if shave:
            d = d[..., shave:-shave, shave:-shave]
        ref = (-10 * d.pow(2).mean([-3, -2, -1]).log10()).sum()
        assert abs(float(sr.psnr_u8(y8, hr8, shave)) - float(ref)) < 1e-3


def test_device_shards_runner(sr):
    """shard.DeviceShards: one process drives every device; with a single GPU two replicas on the same device exercise the slicing,
    the per-replica streams and the host gather.  Result == the plain forward, for an image model and a clip model."""
    from mobilesuperresolution_b200.shard import DeviceShards
    torch.manual_seed(9)
    p = types.SimpleNamespace(image_mean=0.5, num_channels=3, scale=2, num_blocks=2, num_residual_units=24, width_search=False, pretrained=False)
    m = sr.BASIC_MODEL(p).eval()
    ndev = torch.cuda.device_count()
    devs = list(range(ndev)) if ndev > 1 else [0, 0]
    run = DeviceShards(m, devices=devs, precision="bf16")
    x = torch.rand(5, 3, 24, 40).pin_memory()
    y = run(x)
    assert tuple(y.shape) == (5, 3, 48, 80) and not y.is_cuda
    with torch.no_grad():
        ref = m.cuda().set_precision("bf16")(x.cuda().bfloat16()).cpu()
    assert torch.equal(y, ref)
    assert run.slices(5) == ([(0, 3), (3, 5)] if len(devs) == 2 else run.slices(5))
    v = sr.BasicVSR_origin(16, 1).eval()
    runv = DeviceShards(v, devices=devs, precision="fp32")
    clips = torch.rand(3, 2, 3, 64, 64).pin_memory()
    yv = runv(clips, 256, 256)
    with torch.no_grad():
        refv = v.cuda().set_precision("fp32")(clips.cuda(), 256, 256).cpu()
    assert tuple(yv.shape) == (3, 2, 3, 256, 256) and torch.equal(yv, refv)
