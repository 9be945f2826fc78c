"""CPU: the oracle (torch port + plain-C restatement) against the golden vectors that
oracle/make_golden.py produced from the unmodified reference.  No GPU, no /root/reference."""
import numpy as np
import pytest
import torch

from conftest import golden_case, load_golden
from oracle import c_oracle, port, synth

torch.set_grad_enabled(False)


def t(sd):
    return {k: torch.from_numpy(v) for k, v in sd.items()}


@pytest.mark.parametrize("name", ["basic_x4_nb2", "basic_x2_nb3", "basic_x4_nb16"])
def test_basic_model_golden(name):
    meta, arrs, sd, x = golden_case(name)
    y = port.basic_model_forward(t(sd), torch.from_numpy(x), meta["scale"]).numpy()
    assert np.abs(y - arrs["y"]).max() <= 1e-6          # same ops, same backend: bit-exact bar one oneDNN choice
    yc = c_oracle.basic_model_forward(sd, x, meta["scale"])
    assert np.abs(yc - arrs["y"]).max() <= 2e-5


@pytest.mark.parametrize("name", ["pruned_x4_P1", "pruned_x2_P2", "pruned_x2_ragged"])
def test_pruned_model_golden(name):
    meta, arrs, sd, x = golden_case(name)
    y = port.pruned_model_forward(t(sd), torch.from_numpy(x), meta["scale"]).numpy()
    assert np.abs(y - arrs["y"]).max() <= 1e-6
    yc = c_oracle.pruned_model_forward(sd, x, meta["scale"])
    assert np.abs(yc - arrs["y"]).max() <= 2e-5


def test_masked_block_and_gate_golden():
    meta, arrs, sd, x = golden_case("block_masked")
    y = port.wdsr_block_masked(t(sd), "", torch.from_numpy(x)).numpy()
    assert np.abs(y - arrs["y"]).max() <= 1e-6
    for nm in ("agg_layer_keep", "agg_layer_skip"):
        meta, arrs, sd, x = golden_case(nm)
        if meta["alpha1"] >= meta["alpha2"]:
            assert np.array_equal(arrs["y"], x)           # identity branch, models/wdsr_b.py:359-360
        else:
            assert np.abs(port.wdsr_block_masked(t(sd), "", torch.from_numpy(x)).numpy() - arrs["y"]).max() <= 1e-6


def test_rounding_golden(kat):
    for c in kat["rounding"]:
        w = torch.tensor(c["w"], dtype=torch.float32).view(-1, 1, 1, 1)
        assert port.rounding(w, c["least"]).view(-1).int().tolist() == c["keep"]


def test_flow_warp_golden():
    meta, arrs = load_golden("flow_warp_small")
    x = synth.synth_input(meta["xshape"], meta["xseed"], meta["xlo"], meta["xhi"])
    fl = synth.synth_input(meta["fshape"], meta["fseed"], meta["flo"], meta["fhi"])
    for mode in ("zeros", "border"):
        y = port.flow_warp(torch.from_numpy(x), torch.from_numpy(fl), padding_mode=mode).numpy()
        assert np.abs(y - arrs[mode]).max() <= 1e-6
        assert np.abs(c_oracle.flow_warp(x, fl, mode) - arrs[mode]).max() <= 2e-6


def test_spynet_golden():
    meta, arrs = load_golden("spynet_small")
    sd = synth.synth_state_dict(meta["shapes"], meta["wseed"])
    a, b = synth.synth_input(meta["shape"], meta["aseed"]), synth.synth_input(meta["shape"], meta["bseed"])
    f = port.spynet_forward(t(sd), torch.from_numpy(a), torch.from_numpy(b)).numpy()
    assert np.abs(f - arrs["flow"]).max() <= 1e-5
    assert np.abs(c_oracle.spynet_forward(sd, a, b) - arrs["flow"]).max() <= 1e-4


def test_basicvsr_origin_golden():
    meta, arrs, sd, x = golden_case("basicvsr_origin_small")
    h, w = meta["out_hw"]
    y = port.basicvsr_origin_forward(t(sd), torch.from_numpy(x), h, w).numpy()
    s = meta["stride"]
    assert np.abs(y[..., ::s, ::s] - arrs["y_strided"]).max() <= 1e-5
    assert np.abs(y[..., :16, :16] - arrs["y_corner"]).max() <= 1e-5


def test_kat2_pretrained(kat):
    """shipped x2 weights, seeded input."""
    meta, arrs = load_golden("kat2_pretrained_x2")
    z = np.load(__import__("os").path.join(__import__("conftest").GOLDEN, "wdsr_b_x2_16_24_pretrained.npz"))
    sd = {k: z[k] for k in z.files}
    x = torch.rand(1, 3, 64, 64, generator=torch.Generator().manual_seed(1234))
    assert abs(float(x.double().sum()) - kat["x64"]["sum"]) < 1e-6
    y = port.basic_model_forward(t(sd), x, 2)
    assert np.abs(y.numpy() - arrs["y"]).max() <= 1e-6
    assert abs(float(y.double().sum()) - kat["KAT2"]["sum"]) < 1e-2
    assert abs(float(y[0, 1, 64, 64]) - 0.467211) < 1e-5 and abs(float(y.abs().max()) - 1.530487) < 1e-5
    yc = c_oracle.basic_model_forward(sd, x.numpy(), 2)
    assert np.abs(yc - arrs["y"]).max() <= 1e-5


@pytest.mark.parametrize("name", ["mvvsr_nf64", "mvvsr_nf16"])
def test_mvvsr_golden(name):
    """MotionVectorVSR (models/mvvsr_arch.py:56-109): the port against the reference-generated fixture."""
    from oracle import synth
    meta, arrs = load_golden(name)
    sd = {k: torch.from_numpy(v) for k, v in synth.synth_state_dict(meta["shapes"], meta["seed"]).items()}
    x = torch.from_numpy(synth.synth_mv_clip(meta["shape"], meta["input_seed"]))
    with torch.no_grad():
        y = port.mvvsr_forward(sd, x, *meta["size"])
    assert float((y - torch.from_numpy(arrs["y"])).abs().max()) <= 1e-5
