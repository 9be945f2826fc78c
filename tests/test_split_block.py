"""The fork's searchable block (Split_Block / MyAggregationLayer, models/wdsr_b.py:406-546): oracle vs the golden vectors generated
from the unmodified reference (oracle/make_golden_split.py), state_dict layout, and the fused CUDA kernel vs the oracle."""
import numpy as np
import pytest
import torch

from conftest import load_golden

CASES = ["split_block_a", "split_block_b"]


def _case(name):
    from oracle import synth
    meta, arrs = load_golden(name)
    sd = synth.synth_state_dict(meta["shapes"], meta["seed"])
    for k in ("alpha1", "alpha2"):
        if k in meta:
            sd[k][:] = meta[k]
    x = synth.synth_input(meta["shape"], meta["input_seed"], *meta["input_range"])
    return meta, arrs, {k: torch.from_numpy(v) for k, v in sd.items()}, torch.from_numpy(x)


@pytest.mark.parametrize("name", CASES)
def test_oracle_split_block_golden(name):
    from oracle import port
    meta, arrs, sd, x = _case(name)
    with torch.no_grad():
        y = port.split_block(sd, "", x)
    assert float((y - torch.from_numpy(arrs["y"])).abs().max()) <= 1e-6
    assert 0 < meta["kept_channels"] < 24          # the fixtures exercise both searched and passed-through channels


@pytest.mark.parametrize("name", ["my_agg_keep", "my_agg_skip"])
def test_oracle_my_aggregation_layer_golden(name):
    from oracle import port
    meta, arrs, sd, x = _case(name)
    with torch.no_grad():
        y = port.my_aggregation_layer(sd, "", x)
    assert float((y - torch.from_numpy(arrs["y"])).abs().max()) <= 1e-6
    if name == "my_agg_skip":
        assert torch.equal(y, x)


def test_split_block_state_dict_layout():
    """per block alpha(3), beta(3), split.weight(C,1,1,1), body.{3,5,7}.0.body.0 depthwise / .2 pointwise."""
    import mobilesuperresolution_b200 as sr
    m = sr.MyAggregationLayer(num_residual_units=24, kernel_size=3)
    sd = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    meta, _ = load_golden("my_agg_keep")
    assert sd == {k: tuple(v) for k, v in meta["shapes"].items()}
    assert sd["body.7.0.body.0.weight_v"] == (24, 1, 7, 7) and sd["body.5.0.body.2.weight_v"] == (24, 24, 1, 1)
    with pytest.raises(RuntimeError):
        m.eval()(torch.zeros(1, 24, 8, 8), torch.ones(1), torch.zeros(1))      # CPU tensors raise: no fallback
    m.alpha1.data.fill_(0.9), m.alpha2.data.fill_(0.1)
    y, speed = m.eval()(torch.zeros(1, 24, 8, 8), torch.tensor([2.0]), torch.tensor([1.0]))   # skipped block: identity, no kernel
    assert float(speed) == 1.0 + float(m.beta2) * 2.0


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_split_block_cuda_golden(name):
    import mobilesuperresolution_b200 as sr
    from oracle import port
    meta, arrs, sd, x = _case(name)
    m = sr.Split_Block(num_residual_units=24, kernel_size=3).eval()
    m.load_state_dict(sd, strict=True)
    m = m.cuda()
    ref = torch.from_numpy(arrs["y"])
    with torch.no_grad():
        y = m(x.cuda()).cpu()
        yb = m(x.cuda().bfloat16()).float().cpu()
    assert float((y - ref).abs().max()) <= 1e-4, float((y - ref).abs().max())      # fp32 gate
    assert port.psnr_db(yb, ref) >= 50.0
    kept = torch.from_numpy(np.asarray(sr.rounding(sd["split.weight"], 0))).view(-1) > 0
    assert torch.equal(y[:, ~kept], x[:, ~kept])          # channels the search did not select pass through bit-exactly


@pytest.mark.gpu
@pytest.mark.parametrize("c,n,h,w", [(8, 1, 5, 7), (16, 2, 33, 70), (32, 1, 64, 64), (24, 1, 360, 640)])
def test_split_block_cuda_vs_oracle_shapes(c, n, h, w):
    """Other channel counts / sizes (partial tiles; images smaller than the 7x7 halo; the north-star 360p frame) against the oracle."""
    import mobilesuperresolution_b200 as sr
    from oracle import port, synth
    m = sr.Split_Block(num_residual_units=c, kernel_size=3).eval()
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    sd = {k: torch.from_numpy(v) for k, v in synth.synth_state_dict(shapes, 31 + c).items()}
    m.load_state_dict(sd)
    x = torch.from_numpy(synth.synth_input((n, c, h, w), 77 + c, -1.0, 1.0))
    with torch.no_grad():
        ref = port.split_block(sd, "", x)
        y = m.cuda()(x.cuda()).cpu()
    assert float((y - ref).abs().max()) <= 1e-4


@pytest.mark.gpu
def test_my_aggregation_layer_cuda():
    import mobilesuperresolution_b200 as sr
    for name in ("my_agg_keep", "my_agg_skip"):
        meta, arrs, sd, x = _case(name)
        m = sr.MyAggregationLayer(num_residual_units=24, kernel_size=3).eval()
        m.load_state_dict(sd, strict=True)
        m = m.cuda()
        with torch.no_grad():
            y, speed = m(x.cuda(), torch.tensor([2.0], device="cuda"), torch.tensor([1.0], device="cuda"))
        assert float((y.cpu() - torch.from_numpy(arrs["y"])).abs().max()) <= 1e-4
        assert float((speed.cpu() - torch.from_numpy(arrs["speed"])).abs().max()) <= 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("c,n,h,w", [(8, 1, 5, 8), (16, 2, 33, 72), (32, 1, 64, 64), (24, 2, 45, 104), (24, 1, 360, 640)])
def test_split_block_bf16_tensor_core_arm(c, n, h, w, monkeypatch):
    """bf16 tensors whose rows are 16-byte multiples take split_block_tc.cu (depthwise on FFMA, the 1x1s on mma.sync): >= 50 dB against the
    fp32 oracle on the same bf16-rounded input, pass-through channels bit-exact, and within bf16 rounding of the FFMA arm -- both as
    B200SR_SPLIT_IMPL=ffma selects it at create time and as widths that are not a multiple of 8 still take it."""
    import mobilesuperresolution_b200 as sr
    from oracle import port, synth
    m = sr.Split_Block(num_residual_units=c, kernel_size=3).eval()
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    sd = {k: torch.from_numpy(v) for k, v in synth.synth_state_dict(shapes, 31 + c).items()}
    m.load_state_dict(sd)
    x = torch.from_numpy(synth.synth_input((n, c, h, w), 77 + c, -1.0, 1.0)).bfloat16()
    with torch.no_grad():
        ref = port.split_block(sd, "", x.float())
        m = m.cuda()
        y = m(x.cuda()).float().cpu()
        # the FFMA arm on the same input: one column more makes the rows 2 bytes off a 16-byte multiple
        xw = torch.nn.functional.pad(x, (0, 1))
        y_ffma = m(xw.cuda()).float().cpu()
        monkeypatch.setenv("B200SR_SPLIT_IMPL", "ffma")
        m2 = sr.Split_Block(num_residual_units=c, kernel_size=3).eval()
        m2.load_state_dict(sd)
        y_env = m2.cuda()(x.cuda()).float().cpu()
    assert torch.equal(y_env, y_ffma[..., :w])       # a zero column past the edge is what the padding supplies anyway
    assert port.psnr_db(y, ref) >= 50.0, port.psnr_db(y, ref)
    kept = torch.from_numpy(np.asarray(sr.rounding(sd["split.weight"], 0))).view(-1) > 0
    assert torch.equal(y[:, ~kept], x.float()[:, ~kept])
    assert port.psnr_db(y, y_env, peak=float(ref.max() - ref.min())) >= 50.0


@pytest.mark.gpu
@pytest.mark.parametrize("c,n,h,w", [(24, 2, 45, 104), (16, 1, 9, 8), (32, 1, 30, 40), (24, 1, 360, 640)])
def test_split_block_tc_red_zones_and_repeat_under_load(c, n, h, w):
    """compute-sanitizer is closed on this pool (profiles/r02_compute_sanitizer_closed.txt): the persistent cp.async kernel of the bf16 arm
    is checked directly -- x and y sit between canary zones that must survive (partial tiles in both directions, a 3-pixel halo read around
    every tile), and 10 runs give bit-identical output while a second stream keeps the SMs and L2 busy (the double-buffered staging tile,
    the ds / st aliasing and the four barriers per tile are the things a race would show in)."""
    import mobilesuperresolution_b200 as sr
    from mobilesuperresolution_b200 import _lib
    from oracle import synth
    m = sr.Split_Block(num_residual_units=c, kernel_size=3).eval()
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    m.load_state_dict({k: torch.from_numpy(v) for k, v in synth.synth_state_dict(shapes, 5 + c).items()})
    m = m.cuda()
    xs = torch.from_numpy(synth.synth_input((n, c, h, w), 3 + c, -1.0, 1.0)).bfloat16().cuda()
    with torch.no_grad():
        want = m(xs)
    plan = m._plan_obj
    nb = xs.numel() * 2
    RZ = 1 << 16
    arena = torch.full((RZ + nb + RZ + nb + RZ,), 0xA5, dtype=torch.uint8, device="cuda")
    xa, ya = arena[RZ:RZ + nb], arena[RZ + nb + RZ:RZ + nb + RZ + nb]
    xa.copy_(xs.view(torch.uint8).reshape(-1))
    side, noise = torch.cuda.Stream(), torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
    for it in range(10):
        with torch.cuda.stream(side):
            for _ in range(1 + it % 3):
                noise.fill_(it)
        _lib.check(_lib.lib().b200sr_split_forward(plan._h, xa.data_ptr(), ya.data_ptr(), n, h, w, _lib.BF16, _lib.current_stream_ptr(xs.device)))
        torch.cuda.synchronize()
        assert torch.equal(ya.view(torch.bfloat16).view(n, c, h, w), want), f"iteration {it}"
    for a, b in [(0, RZ), (RZ + nb, RZ + nb + RZ), (RZ + 2 * nb + RZ, RZ + 2 * nb + 2 * RZ)]:
        assert bool((arena[a:b] == 0xA5).all()), "the kernel wrote outside its output"
    assert torch.equal(xa.view(torch.bfloat16).view(n, c, h, w), xs)
