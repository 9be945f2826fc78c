"""GPU parity: the CUDA WDSR-B path (through the C ABI) against the golden vectors / the oracle.

Gates: fp32 arithmetic raw max-abs <= 1e-4; bf16 arithmetic PSNR >= 50 dB with
peak = reference dynamic range for synthetic/random-init weights and peak = 1 for the shipped pretrained weights.
"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import GOLDEN, golden_case, load_golden

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-4
BF16_PSNR = 50.0


@pytest.fixture(scope="module")
def G():
    import util_gpu
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return util_gpu


def _t(sd):
    return {k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}


# ---------------------------------------------------------------- stage level ---------------------------
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_stages_against_oracle(G, precision):
    """head / one fused block / tail separately, with non-zero biases and an image not a multiple of the tile."""
    from oracle import port
    sr = G.sr
    m = sr.BASIC_MODEL(G.params(4, 2)).eval()
    sd = G.synth_load(m, 51)
    m = m.to(G.DEV).set_precision(precision)
    plan = m.prepare()
    x = torch.from_numpy(G.synth.synth_input((2, 3, 37, 45), 52))
    xd = x.to(G.DEV)
    # head
    x0 = x - 0.5
    ref_head = F.conv2d(x0, port.weight_norm_fold(sd["head.weight_g"], sd["head.weight_v"]), sd["head.bias"], padding=1)
    t = plan.head(xd, precision)
    torch.cuda.synchronize()
    got = t.float().cpu().permute(0, 3, 1, 2)
    assert G.maxabs(got, ref_head) <= (1e-5 if precision == "fp32" else 2e-2)
    # block 0 on the oracle's head output (so errors do not compound)
    tin = ref_head.permute(0, 2, 3, 1).contiguous().to(G.DEV)
    if precision == "bf16":
        tin = tin.bfloat16()
    ref_blk = port.wdsr_block(sd, "body.0.", tin.float().cpu().permute(0, 3, 1, 2))
    got = plan.block(0, tin, precision).float().cpu().permute(0, 3, 1, 2)
    if precision == "fp32":
        assert G.maxabs(got, ref_blk) <= 2e-5
    else:
        assert G.psnr(got, ref_blk) >= 50.0        # the same gate as end to end (measured ~69 dB)
    # tail
    ref_tail = F.pixel_shuffle(
        F.conv2d(ref_blk, port.weight_norm_fold(sd["tail.weight_g"], sd["tail.weight_v"]), sd["tail.bias"], padding=1) +
        F.conv2d(x0, port.weight_norm_fold(sd["skip.0.weight_g"], sd["skip.0.weight_v"]), sd["skip.0.bias"], padding=2), 4) + 0.5
    tt = ref_blk.permute(0, 2, 3, 1).contiguous().to(G.DEV)
    xin = xd
    if precision == "bf16":
        tt, xin = tt.bfloat16(), xd.bfloat16()
    got = plan.tail(tt, xin, precision).float().cpu()
    if precision == "fp32":
        assert G.maxabs(got, ref_tail) <= 2e-5
    else:
        assert G.psnr(got, ref_tail) >= 50.0


# ---------------------------------------------------------------- golden fixtures -----------------------
@pytest.mark.parametrize("name", ["basic_x4_nb2", "basic_x2_nb3", "basic_x4_nb16"])
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_basic_model_golden(G, name, precision):
    meta, arrs, sd, x = golden_case(name)
    m = G.load_np_state(G.sr.BASIC_MODEL(G.params(meta["scale"], meta["nb"])), sd)
    y = G.run(m, torch.from_numpy(x), precision)
    ref = torch.from_numpy(arrs["y"])
    if precision == "fp32":
        assert G.maxabs(y, ref) <= FP32_TOL
    else:
        assert G.psnr(y, ref) >= BF16_PSNR


@pytest.mark.parametrize("name", ["pruned_x4_P1", "pruned_x2_P2", "pruned_x2_ragged"])
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_pruned_model_golden(G, name, precision):
    meta, arrs, sd, x = golden_case(name)
    fn = G.block_index_file(meta["widths"])
    m = G.load_np_state(G.sr.Model(meta["scale"], fn), sd)
    os.unlink(fn)
    y = G.run(m, torch.from_numpy(x), precision)
    ref = torch.from_numpy(arrs["y"])
    if precision == "fp32":
        assert G.maxabs(y, ref) <= FP32_TOL
    else:
        assert G.psnr(y, ref) >= BF16_PSNR


def test_kat1_seeded_init_64x64(G, kat):
    """SURVEY.md App. D KAT1 = BASELINE cfg1: x4 16/24, reference's own seeded init, one 64x64 patch, fp32."""
    meta, arrs = load_golden("kat1_basic_x4_seed0")
    torch.manual_seed(0)
    m = G.sr.BASIC_MODEL(G.params(4, 16)).eval()
    wsum = float(sum(v.double().sum() for v in m.state_dict().values()))
    assert abs(wsum - kat["KAT1"]["weights_sum"]) < 1e-6, "constructor no longer consumes the RNG like the reference"
    x = torch.rand(1, 3, 64, 64, generator=torch.Generator().manual_seed(1234))
    y = G.run(m, x, "fp32")
    s = meta["stride"]
    assert G.maxabs(y[:, :, ::s, ::s], torch.from_numpy(arrs["y_strided"])) <= FP32_TOL
    assert G.maxabs(y[:, :, :12, :12], torch.from_numpy(arrs["y_corner"])) <= FP32_TOL
    assert abs(float(y.double().sum()) - kat["KAT1"]["sum"]) < 2.0          # 196,608 outputs, each within 1e-5
    yb = G.run(m, x, "bf16")
    ref_like = y                                                              # fp32 CUDA result verified just above
    assert G.psnr(yb, ref_like) >= BF16_PSNR


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_kat2_pretrained_x2(G, precision):
    """Shipped x2 weights (trained, non-zero biases, output range ~[-0.6,1.6]): bf16 gate uses peak = 1."""
    meta, arrs = load_golden("kat2_pretrained_x2")
    z = np.load(os.path.join(GOLDEN, "wdsr_b_x2_16_24_pretrained.npz"))
    m = G.load_np_state(G.sr.BASIC_MODEL(G.params(2, 16)), {k: z[k] for k in z.files})
    x = torch.rand(1, 3, 64, 64, generator=torch.Generator().manual_seed(1234))
    y = G.run(m, x, precision)
    ref = torch.from_numpy(arrs["y"])
    if precision == "fp32":
        assert G.maxabs(y, ref) <= FP32_TOL
    else:
        assert G.psnr(y, ref, peak=1.0) >= BF16_PSNR


# ---------------------------------------------------------------- oracle at odd / larger sizes ----------
@pytest.mark.parametrize("shape,scale,nb", [((3, 3, 33, 47), 4, 3), ((1, 3, 1, 1), 2, 2), ((1, 3, 5, 130), 3, 2),
                                            ((2, 3, 70, 9), 4, 1), ((1, 3, 17, 16), 2, 0)])
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_ragged_shapes_vs_oracle(G, shape, scale, nb, precision):
    from oracle import port
    m = G.sr.BASIC_MODEL(G.params(scale, nb)).eval()
    sd = G.synth_load(m, 60 + nb)
    x = torch.from_numpy(G.synth.synth_input(shape, 61))
    ref = port.basic_model_forward(sd, x, scale)
    y = G.run(m, x, precision)
    assert y.shape == ref.shape
    if precision == "fp32":
        assert G.maxabs(y, ref) <= FP32_TOL
    else:
        assert G.psnr(y, ref) >= BF16_PSNR


def test_empty_batch(G):
    m = G.sr.BASIC_MODEL(G.params(4, 1)).to(G.DEV).eval()
    y = m(torch.empty(0, 3, 8, 8, device=G.DEV))
    assert tuple(y.shape) == (0, 3, 32, 32)


def test_masked_supernet_equals_sliced(G):
    """P3: NAS supernet with width masks and a depth gate -> pruned plan == the reference's masked forward."""
    from oracle import port
    m = G.sr.NAS_MODEL_classic(G.params(2, 4, width_search=True)).eval()
    sd = G.synth_load(m, 71)
    sd["body.1.alpha1"], sd["body.1.alpha2"] = torch.tensor([0.9]), torch.tensor([0.1])      # block 1 skipped
    sd["body.2.alpha1"], sd["body.2.alpha2"] = torch.tensor([0.1]), torch.tensor([0.9])
    m.load_state_dict(sd)
    x = torch.from_numpy(G.synth.synth_input((1, 3, 19, 23), 72))
    ref = port.supernet_classic_forward(sd, x, 2)
    out, speed = m.to(G.DEV).set_precision("fp32")(x.to(G.DEV))
    assert G.maxabs(out.cpu(), ref) <= FP32_TOL
    assert speed.numel() == 1
    widths = m.get_width_from_block_idx(m.get_block_status())
    assert 1 not in m.get_block_status() and all(8 <= w[0] <= 24 for w in widths)


def test_weight_mutation_invalidates_cache(G):
    """The reference re-folds weight-norm every forward: mutating weight_g must change the output here too."""
    from oracle import port
    m = G.sr.BASIC_MODEL(G.params(2, 1)).eval()
    G.synth_load(m, 81)
    x = torch.from_numpy(G.synth.synth_input((1, 3, 12, 12), 82))
    y0 = G.run(m, x, "fp32")
    with torch.no_grad():
        m.tail.weight_g.mul_(1.5)
    y1 = G.run(m, x, "fp32")
    ref = port.basic_model_forward({k: v.cpu() for k, v in m.state_dict().items()}, x, 2)
    assert G.maxabs(y0, y1) > 1e-3 and G.maxabs(y1, ref) <= FP32_TOL


def test_standalone_block_module(G):
    from oracle import port
    b = G.sr.Block(num_residual_units=24, kernel_size=3, res_scale=0.25, width_search=True).eval()
    meta, arrs, sd, x = golden_case("block_masked")
    G.load_np_state(b, sd)
    y = b.to(G.DEV)(torch.from_numpy(x).to(G.DEV)).cpu()
    assert G.maxabs(y, torch.from_numpy(arrs["y"])) <= 2e-5


# ---------------------------------------------------------------- BASELINE sizes ------------------------
def test_cfg2_batch64_patches_bf16(G):
    """cfg2: x4 16/24, batch 64 of 96x96, bf16.  Oracle on 3 patches + batch-independence on the whole batch."""
    from oracle import port
    torch.manual_seed(0)
    m = G.sr.BASIC_MODEL(G.params(4, 16)).eval()
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    x = torch.rand(64, 3, 96, 96, generator=torch.Generator().manual_seed(1234))
    y = G.run(m, x, "bf16")
    assert tuple(y.shape) == (64, 3, 384, 384)
    for i in (0, 31, 63):
        ref = port.basic_model_forward(sd, x[i:i + 1], 4)
        assert G.psnr(y[i:i + 1], ref) >= BF16_PSNR
    y1 = G.run(m, x[17:18], "bf16")
    assert torch.equal(y1, y[17:18]), "a patch's output must not depend on its batch neighbours"
    yf = G.run(m, x[:2], "fp32")
    assert G.maxabs(yf, port.basic_model_forward(sd, x[:2], 4)) <= FP32_TOL


@pytest.mark.parametrize("which", ["dense", "P1"])
def test_cfg3_360p_to_1440p(G, which):
    """cfg3: 640x360 -> 2560x1440 frames; dense 16/24 (north-star target) and the searched widths P1."""
    from oracle import port
    x = torch.rand(1, 3, 360, 640, generator=torch.Generator().manual_seed(1234))
    if which == "dense":
        torch.manual_seed(0)
        m = G.sr.BASIC_MODEL(G.params(4, 16)).eval()
        sd = {k: v.clone() for k, v in m.state_dict().items()}
        ref = port.basic_model_forward(sd, x, 4)
    else:
        P1 = [(9, 91, 14), (9, 94, 10), (9, 107, 12), (9, 110, 13), (9, 115, 12), (9, 94, 12), (9, 115, 17), (9, 116, 16)]
        fn = G.block_index_file(P1)
        m = G.sr.Model(4, fn).eval()
        os.unlink(fn)
        sd = G.synth_load(m, 91)
        ref = port.pruned_model_forward(sd, x, 4)
    assert G.maxabs(G.run(m, x, "fp32"), ref) <= FP32_TOL
    assert G.psnr(G.run(m, x, "bf16"), ref) >= BF16_PSNR


def test_cfg5_1080p_x2_crops(G):
    """cfg5: x2 1920x1080 -> 3840x2160.  The net's receptive field is 19 px, so the oracle on a crop with a 24 px
    margin reproduces the interior of the full-frame result (and crops touching the frame border check padding)."""
    from oracle import port
    z = np.load(os.path.join(GOLDEN, "wdsr_b_x2_16_24_pretrained.npz"))
    sd = {k: torch.from_numpy(z[k]) for k in z.files}
    m = G.load_np_state(G.sr.BASIC_MODEL(G.params(2, 16)), {k: z[k] for k in z.files})
    x = torch.rand(2, 3, 1080, 1920, generator=torch.Generator().manual_seed(5))
    yf = G.run(m, x, "fp32")
    yb = G.run(m, x, "bf16")
    M = 24
    for (n, y0, x0, hh, ww) in [(0, 0, 0, 96, 96), (1, 1080 - 96, 1920 - 96, 96, 96), (0, 500, 900, 96, 128)]:
        ya, xa = max(y0 - M, 0), max(x0 - M, 0)
        yb_, xb_ = min(y0 + hh + M, 1080), min(x0 + ww + M, 1920)
        ref = port.basic_model_forward(sd, x[n:n + 1, :, ya:yb_, xa:xb_], 2)
        ref = ref[:, :, 2 * (y0 - ya):2 * (y0 - ya + hh), 2 * (x0 - xa):2 * (x0 - xa + ww)]
        assert G.maxabs(yf[n:n + 1, :, 2 * y0:2 * (y0 + hh), 2 * x0:2 * (x0 + ww)], ref) <= FP32_TOL
        assert G.psnr(yb[n:n + 1, :, 2 * y0:2 * (y0 + hh), 2 * x0:2 * (x0 + ww)], ref, peak=1.0) >= BF16_PSNR


# ---------------------------------------------------------------- stream / graph behaviour ---------------
@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_cuda_graph_capture_and_determinism(G, precision):
    """The forward must be CUDA-graph capturable on a side stream (the reference's callers own the stream) and
    bit-reproducible: the block kernels are launched with programmatic stream serialization, so consecutive launches overlap
    their prologues -- replaying the graph must still give exactly the eager result, every time."""
    torch.manual_seed(3)
    m = G.sr.BASIC_MODEL(G.params(4, 4)).eval().to(G.DEV).set_precision(precision)
    x = torch.rand(3, 3, 70, 90, device=G.DEV)
    if precision == "bf16":
        x = x.bfloat16()
    with torch.no_grad():
        ref = m(x).clone()
        st = torch.cuda.Stream()
        st.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(st):
            for _ in range(2):
                y = m(x)
            st.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=st):
                y = m(x)
            outs = []
            for _ in range(3):
                y.zero_()
                g.replay()
                st.synchronize()
                outs.append(y.clone())
    for o in outs:
        assert torch.equal(o, ref)


def test_back_to_back_forwards_do_not_race(G):
    """Two different inputs through the same plan and workspace, queued back to back without a host sync: the second forward's
    first block kernel may start its prologue while the first forward's tail still runs, but must not touch the trunk early."""
    torch.manual_seed(4)
    m = G.sr.BASIC_MODEL(G.params(4, 3)).eval().to(G.DEV).set_precision("bf16")
    xa = torch.rand(8, 3, 96, 96, device=G.DEV).bfloat16()
    xb = torch.rand(8, 3, 96, 96, device=G.DEV).bfloat16()
    with torch.no_grad():
        ra, rb = m(xa).clone(), m(xb).clone()
        torch.cuda.synchronize()
        for _ in range(5):
            ya = m(xa)
            yb = m(xb)
            torch.cuda.synchronize()
            assert torch.equal(ya, ra) and torch.equal(yb, rb)


@pytest.mark.gpu
def test_forward_u8_matches_quantised_forward():
    """The 8-bit frame written by the tcgen05 tail epilogue == (sr * 255).round().clamp(0, 255) of the same forward's float output
    (common/metrics.py:12), and only the tcgen05 bf16 path offers it (fp32 precision raises: no silent conversion pass)."""
    import types
    import mobilesuperresolution_b200 as sr
    torch.manual_seed(11)
    for scale in (2, 4):
        p = types.SimpleNamespace(image_mean=0.5, num_channels=3, scale=scale, num_blocks=2, num_residual_units=24, width_search=False,
                                  pretrained=False)
        m = sr.BASIC_MODEL(p).cuda().eval().set_precision("bf16")
        x = torch.rand(2, 3, 40, 72, device="cuda")
        with torch.no_grad():
            plan = m.prepare()
            yf = plan.forward(x, "bf16", out_dtype=torch.float32)          # same arithmetic, float32 store
            yu = m.forward_u8(x)
        assert yu.dtype == torch.uint8 and tuple(yu.shape) == (2, 3, 40 * scale, 72 * scale)
        ref = (yf * 255).round().clamp(0, 255)
        d = (yu.float() - ref).abs()
        assert float(d.max()) <= 1.0 and float((d > 0).float().mean()) < 1e-3      # ties at x.5 may round differently after the fp32 mul
    with pytest.raises(RuntimeError):
        m.set_precision("fp32").forward_u8(x)


@pytest.mark.parametrize("shape", [(2, 3, 37, 45), (1, 3, 96, 96), (1, 3, 5, 300), (3, 3, 1, 1)])
@pytest.mark.parametrize("xdtype", ["fp32", "bf16"])
def test_head_mma_and_tcgen05_forms(G, shape, xdtype, monkeypatch):
    """The head of the tcgen05 path runs on mma.sync by default (csrc/wdsr_head_mma.cu); B200SR_HEAD_IMPL=tc5 keeps the tcgen05 kernel
    (csrc/wdsr_tc5_head.cuh).  Both round (x - mean) and the filters to bf16 at the same points: each within the bf16 gate of the fp32
    reference conv, and within a couple of bf16 ulps of each other, on ragged and tiny shapes, fp32 and bf16 input."""
    from oracle import port
    sr = G.sr
    x = torch.from_numpy(G.synth.synth_input(shape, 91))
    outs = {}
    for impl in ("mma", "tc5"):
        monkeypatch.setenv("B200SR_HEAD_IMPL", impl)
        m = sr.BASIC_MODEL(G.params(4, 1)).eval()
        sd = G.synth_load(m, 57)
        plan = m.to(G.DEV).set_precision("bf16").prepare()
        xd = x.to(G.DEV)
        if xdtype == "bf16":
            xd = xd.bfloat16()
        outs[impl] = plan.head(xd, "bf16").float().cpu().permute(0, 3, 1, 2)
        torch.cuda.synchronize()
    xin = x.bfloat16().float() if xdtype == "bf16" else x
    ref = F.conv2d(xin - 0.5, port.weight_norm_fold(sd["head.weight_g"], sd["head.weight_v"]), sd["head.bias"], padding=1)
    for impl in outs:
        assert G.maxabs(outs[impl], ref) <= 2e-2, impl
    assert G.maxabs(outs["mma"], outs["tc5"]) <= 2 ** -6 * max(1.0, float(ref.abs().max()))
