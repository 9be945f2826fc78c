"""Generate tests/golden/* from the UNMODIFIED reference (run in the build container only).

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.

For every case it (1) builds the reference module from /root/reference, (2) loads
deterministic synthetic weights (``oracle.synth``; non-zero biases) or the reference's own
seeded init / shipped pretrained weights, (3) runs the reference forward on CPU fp32,
(4) asserts the torch port (``oracle.port``) reproduces it BIT-EXACTLY and the plain-C oracle
(``oracle.c_oracle``) to <= 1e-4, and (5) writes inputs-by-seed + outputs as small fixtures.
The GPU box has no /root/reference: tests there pin the oracle against these files.
"""
from __future__ import annotations

import json
import os
import sys
import tempfile
import types

import numpy as np
import torch

sys.dont_write_bytecode = True
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import c_oracle, port, ref_import, synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
torch.set_grad_enabled(False)


def params(scale, nb, nru=24):
    return types.SimpleNamespace(image_mean=0.5, num_channels=3, scale=scale, num_blocks=nb,
                                 num_residual_units=nru, width_search=False, pretrained=False)


def load_synth(module, seed):
    shapes = {k: tuple(v.shape) for k, v in module.state_dict().items()}
    sd_np = synth.synth_state_dict(shapes, seed)
    module.load_state_dict({k: torch.from_numpy(v) for k, v in sd_np.items()}, strict=True)
    return shapes, sd_np


def tsd(sd_np):
    return {k: torch.from_numpy(v) for k, v in sd_np.items()}


def check(name, ref, via_port, via_c=None, tol_c=1e-4):
    d = float((via_port - ref).abs().max())
    assert d == 0.0, f"{name}: torch port differs from reference by {d}"
    msg = f"{name}: port==ref exactly"
    if via_c is not None:
        dc = float(np.abs(via_c - ref.numpy()).max())
        assert dc <= tol_c, f"{name}: C oracle differs by {dc}"
        msg += f"; C oracle max-abs {dc:.2e}"
    print(msg, f"(range [{float(ref.min()):.3f},{float(ref.max()):.3f}])")


def save(name, meta, **arrays):
    np.savez_compressed(os.path.join(OUT, name + ".npz"), meta=np.frombuffer(json.dumps(meta).encode(), np.uint8),
                        **arrays)


def block_index_file(widths):
    """The search artefact format export_onnx.Model.file_reader consumes (export_onnx.py:81-88)."""
    f = tempfile.NamedTemporaryFile("w", suffix="_block_index.txt", delete=False)
    f.write(repr((list(range(len(widths))), [list(w) for w in widths])) + "\n")
    f.close()
    return f.name


P1 = [(9, 91, 14), (9, 94, 10), (9, 107, 12), (9, 110, 13), (9, 115, 12), (9, 94, 12), (9, 115, 17), (9, 116, 16)]
P2 = [(14, 123, 20), (14, 126, 20), (14, 137, 20), (14, 141, 20), (14, 144, 20)]


def main():
    os.makedirs(OUT, exist_ok=True)
    R = ref_import.modules()
    kat = {}

    # ---------------- Appendix D known answers (reference's own seeded init) ----------------
    x64 = torch.rand(1, 3, 64, 64, generator=torch.Generator().manual_seed(1234))
    kat["x64"] = {"sum": float(x64.double().sum()), "first3": [float(v) for v in x64[0, 0, 0, :3]]}
    torch.manual_seed(0)
    m = R.BASIC_MODEL(params(4, 16)).eval()
    y = m(x64)
    sd = m.state_dict()
    check("KAT1 basic x4 16/24 seed0", y, port.basic_model_forward(sd, x64, 4))
    kat["KAT1"] = {"sum": float(y.double().sum()), "absmax": float(y.abs().max()),
                   "y[0,0,0,0]": float(y[0, 0, 0, 0]), "y[0,1,100,200]": float(y[0, 1, 100, 200]),
                   "y[0,2,255,255]": float(y[0, 2, 255, 255]),
                   "weights_sum": float(sum(v.double().sum() for v in sd.values())),
                   "head.weight_v[0,0,0,:]": [float(v) for v in sd["head.weight_v"][0, 0, 0, :]]}
    save("kat1_basic_x4_seed0", {"scale": 4, "nb": 16, "nru": 24, "weights": "torch.manual_seed(0) reference init",
                                 "input": "torch.rand(1,3,64,64, Generator().manual_seed(1234))", "stride": 8},
         y_strided=y[:, :, ::8, ::8].numpy(), y_corner=y[:, :, :12, :12].numpy())

    pre = torch.load(os.path.join(ref_import.REF, "models/pretrained_weights/wdsr_b_x2_16_24.pt"), map_location="cpu")
    m2 = R.BASIC_MODEL(params(2, 16)).eval()
    m2.load_state_dict(pre, strict=True)
    y2 = m2(x64)
    yc = c_oracle.basic_model_forward({k: v.numpy() for k, v in pre.items()}, x64.numpy(), 2)
    check("KAT2 pretrained x2", y2, port.basic_model_forward(pre, x64, 2), yc)
    kat["KAT2"] = {"sum": float(y2.double().sum()), "absmax": float(y2.abs().max()), "y[0,0,0,0]": float(y2[0, 0, 0, 0]),
                   "y[0,1,64,64]": float(y2[0, 1, 64, 64])}
    # the shipped trained weights are a data fixture (real-range weights, trained non-zero biases)
    np.savez_compressed(os.path.join(OUT, "wdsr_b_x2_16_24_pretrained.npz"), **{k: v.numpy() for k, v in pre.items()})
    save("kat2_pretrained_x2", {"scale": 2, "nb": 16, "nru": 24, "weights": "wdsr_b_x2_16_24_pretrained.npz",
                                "input": "torch.rand(1,3,64,64, Generator().manual_seed(1234))"}, y=y2.numpy())

    torch.manual_seed(0)
    sp = R.spynet_arch.SpyNet().eval()
    g = torch.Generator().manual_seed(7)
    a = torch.rand(2, 3, 180, 320, generator=g)
    b = torch.rand(2, 3, 180, 320, generator=g)
    f = sp(a, b)
    check("KAT3 spynet seed0 180x320", f, port.spynet_forward(sp.state_dict(), a, b))
    kat["KAT3"] = {"sum": float(f.double().sum()), "absmax": float(f.abs().max()), "f[0,0,90,160]": float(f[0, 0, 90, 160]),
                   "f[1,1,0,0]": float(f[1, 1, 0, 0]),
                   "weights_sum": float(sum(v.double().sum() for v in sp.state_dict().values()))}
    save("kat3_spynet_seed0", {"weights": "torch.manual_seed(0); SpyNet()", "stride": 10,
                               "input": "g=Generator().manual_seed(7); a=rand(2,3,180,320,g); b=rand(2,3,180,320,g)"},
         f_strided=f[:, :, ::10, ::10].numpy())

    g = torch.Generator().manual_seed(9)
    feat = torch.rand(1, 8, 45, 80, generator=g)
    fl = (torch.rand(1, 45, 80, 2, generator=g) - 0.5) * 20
    wz = R.spynet_arch.flow_warp(feat, fl)
    wb = R.spynet_arch.flow_warp(feat, fl, padding_mode="border")
    check("KAT4 flow_warp zeros", wz, port.flow_warp(feat, fl), c_oracle.flow_warp(feat.numpy(), fl.numpy()), 2e-5)
    check("KAT4 flow_warp border", wb, port.flow_warp(feat, fl, padding_mode="border"),
          c_oracle.flow_warp(feat.numpy(), fl.numpy(), "border"), 2e-5)
    kat["KAT4"] = {"zeros_sum": float(wz.double().sum()), "w[0,0,0,0]": float(wz[0, 0, 0, 0]),
                   "w[0,3,22,40]": float(wz[0, 3, 22, 40]), "border_sum": float(wb.double().sum())}

    # ---------------- synthetic-weight cases (numpy streams; non-zero biases) ----------------
    cases = []

    def wdsr_case(name, scale, nb, shape, seed):
        mod = R.BASIC_MODEL(params(scale, nb)).eval()
        shapes, sd_np = load_synth(mod, seed)
        x_np = synth.synth_input(shape, seed + 100)
        yy = mod(torch.from_numpy(x_np))
        check(name, yy, port.basic_model_forward(tsd(sd_np), torch.from_numpy(x_np), scale),
              c_oracle.basic_model_forward(sd_np, x_np, scale))
        save(name, {"kind": "basic", "scale": scale, "nb": nb, "nru": 24, "shapes": shapes, "wseed": seed,
                    "xshape": list(shape), "xseed": seed + 100}, y=yy.numpy())
        cases.append(name)

    wdsr_case("basic_x4_nb2", 4, 2, (2, 3, 13, 11), 1)
    wdsr_case("basic_x2_nb3", 2, 3, (1, 3, 9, 37), 2)
    wdsr_case("basic_x4_nb16", 4, 16, (1, 3, 20, 24), 3)

    def pruned_case(name, scale, widths, shape, seed):
        fn = block_index_file(widths)
        mod = R.export_onnx.Model(scale, fn).eval()
        os.unlink(fn)
        shapes, sd_np = load_synth(mod, seed)
        x_np = synth.synth_input(shape, seed + 100)
        yy = mod(torch.from_numpy(x_np))
        check(name, yy, port.pruned_model_forward(tsd(sd_np), torch.from_numpy(x_np), scale),
              c_oracle.pruned_model_forward(sd_np, x_np, scale))
        save(name, {"kind": "pruned", "scale": scale, "widths": [list(w) for w in widths], "shapes": shapes,
                    "wseed": seed, "xshape": list(shape), "xseed": seed + 100}, y=yy.numpy())
        cases.append(name)

    pruned_case("pruned_x4_P1", 4, P1, (1, 3, 12, 18), 4)
    pruned_case("pruned_x2_P2", 2, P2, (1, 3, 15, 10), 5)
    pruned_case("pruned_x2_ragged", 2, [(11, 53, 8), (11, 37, 19), (11, 8, 9)], (2, 3, 7, 9), 6)

    # masked supernet block (models/wdsr_b.py Block(width_search=True)) and depth gate
    blk = R.wdsr_b.Block(num_residual_units=24, kernel_size=3, res_scale=0.25, width_search=True).eval()
    shapes, sd_np = load_synth(blk, 7)
    x_np = synth.synth_input((1, 24, 9, 11), 107, -1.0, 1.0)
    yy = blk(torch.from_numpy(x_np))
    check("block_masked", yy, port.wdsr_block_masked(tsd(sd_np), "", torch.from_numpy(x_np)))
    save("block_masked", {"kind": "block_masked", "shapes": shapes, "wseed": 7, "xshape": [1, 24, 9, 11], "xseed": 107,
                          "xlo": -1.0, "xhi": 1.0}, y=yy.numpy())
    for nm, seed in (("agg_layer_keep", 8), ("agg_layer_skip", 9)):
        agg = R.wdsr_b.AggregationLayer(num_residual_units=24, kernel_size=3, res_scale=0.25, width_search=True).eval()
        shapes, sd_np = load_synth(agg, seed)
        if nm.endswith("skip"):
            sd_np["alpha1"], sd_np["alpha2"] = np.asarray([0.9], np.float32), np.asarray([0.1], np.float32)
        else:
            sd_np["alpha1"], sd_np["alpha2"] = np.asarray([0.1], np.float32), np.asarray([0.9], np.float32)
        agg.load_state_dict(tsd(sd_np))
        yy, _ = agg(torch.from_numpy(x_np), torch.zeros(1), torch.zeros(1))
        exp = torch.from_numpy(x_np) if nm.endswith("skip") else port.wdsr_block_masked(tsd(sd_np), "", torch.from_numpy(x_np))
        check(nm, yy, exp)
        save(nm, {"kind": "agg", "shapes": shapes, "wseed": seed, "alpha1": float(sd_np["alpha1"][0]),
                  "alpha2": float(sd_np["alpha2"][0]), "xshape": [1, 24, 9, 11], "xseed": 107, "xlo": -1.0, "xhi": 1.0},
             y=yy.numpy())

    # rounding() incl. the top-k floor (models/ops.py:33-43)
    rs = np.random.RandomState(11)
    rounding_cases = []
    for c, lo, hi in ((24, 0.2, 1.0), (24, 0.0, 0.45), (144, 0.2, 1.0), (20, 0.0, 0.6), (8, 0.0, 0.4)):
        w = rs.uniform(lo, hi, (c, 1, 1, 1)).astype(np.float32)
        for least in (8, 0):
            keep = R.ops.rounding(torch.from_numpy(w), least).numpy()
            assert np.array_equal(keep, port.rounding(torch.from_numpy(w), least).numpy())
            rounding_cases.append({"w": w.reshape(-1).tolist(), "least": least, "keep": keep.reshape(-1).astype(int).tolist()})
    kat["rounding"] = rounding_cases

    # flow_warp, odd sizes, large flows (OOB paths)
    x_np = synth.synth_input((2, 5, 12, 17), 21, -1.0, 1.0)
    fl_np = synth.synth_input((2, 12, 17, 2), 22, -9.0, 9.0)
    wz = R.spynet_arch.flow_warp(torch.from_numpy(x_np), torch.from_numpy(fl_np))
    wb = R.spynet_arch.flow_warp(torch.from_numpy(x_np), torch.from_numpy(fl_np), padding_mode="border")
    check("flow_warp_small zeros", wz, port.flow_warp(torch.from_numpy(x_np), torch.from_numpy(fl_np)),
          c_oracle.flow_warp(x_np, fl_np), 2e-5)
    check("flow_warp_small border", wb, port.flow_warp(torch.from_numpy(x_np), torch.from_numpy(fl_np), padding_mode="border"),
          c_oracle.flow_warp(x_np, fl_np, "border"), 2e-5)
    save("flow_warp_small", {"kind": "flow_warp", "xshape": [2, 5, 12, 17], "xseed": 21, "xlo": -1.0, "xhi": 1.0,
                             "fshape": [2, 12, 17, 2], "fseed": 22, "flo": -9.0, "fhi": 9.0},
         zeros=wz.numpy(), border=wb.numpy())

    # SPyNet with synthetic weights, size not a multiple of 32
    sp = R.spynet_arch.SpyNet().eval()
    shapes, sd_np = load_synth(sp, 31)
    a_np, b_np = synth.synth_input((1, 3, 40, 72), 32), synth.synth_input((1, 3, 40, 72), 33)
    f = sp(torch.from_numpy(a_np), torch.from_numpy(b_np))
    check("spynet_small", f, port.spynet_forward(tsd(sd_np), torch.from_numpy(a_np), torch.from_numpy(b_np)),
          c_oracle.spynet_forward(sd_np, a_np, b_np), 2e-4)
    save("spynet_small", {"kind": "spynet", "shapes": shapes, "wseed": 31, "shape": [1, 3, 40, 72], "aseed": 32, "bseed": 33},
         flow=f.numpy())

    # BasicVSR_origin end to end (tiny), and the fork's get_flow/propagation + its shape error
    vs = R.basicvsr_origin.BasicVSR_origin(num_feat=16, num_block=2).eval()
    shapes, sd_np = load_synth(vs, 41)
    clip = synth.synth_input((1, 3, 3, 36, 68), 42)
    o = vs(torch.from_numpy(clip), 144, 272)
    check("basicvsr_origin_small", o, port.basicvsr_origin_forward(tsd(sd_np), torch.from_numpy(clip), 144, 272))
    save("basicvsr_origin_small", {"kind": "basicvsr_origin", "num_feat": 16, "num_block": 2, "shapes": shapes, "wseed": 41,
                                   "xshape": [1, 3, 3, 36, 68], "xseed": 42, "out_hw": [144, 272], "stride": 4},
         y_strided=o[..., ::4, ::4].numpy(), y_corner=o[..., :16, :16].numpy())
    fk = R.basicvsr_fork.BasicVSR(num_feat=8, num_block=1).eval()
    try:
        fk(torch.from_numpy(clip), 144, 272)
        kat["fork_basicvsr_error"] = None
    except RuntimeError as e:   # conv_last -> num_feat channels added to a 3-channel base
        kat["fork_basicvsr_error"] = type(e).__name__
    print("fork BasicVSR(num_feat=8) forward raises:", kat["fork_basicvsr_error"])

    kat["cases"] = cases
    kat["torch"] = torch.__version__
    with open(os.path.join(OUT, "kat.json"), "w") as fh:
        json.dump(kat, fh, indent=1)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
