"""CPU: host-side logic of the drop-in boundary -- no compute calls (there is no GPU here and no CPU fallback)."""
import ctypes
import os
import re
import subprocess
import sys
import types

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import ROOT, golden_case

torch.set_grad_enabled(False)


@pytest.fixture(scope="module")
def sr():
    from mobilesuperresolution_b200 import build
    build.build()                      # nvcc cross-compiles sm_100a without a GPU
    import mobilesuperresolution_b200 as m
    return m


def P(scale=4, nb=16, nru=24, ws=False):
    return types.SimpleNamespace(image_mean=0.5, num_channels=3, scale=scale, num_blocks=nb, num_residual_units=nru,
                                 width_search=ws, pretrained=False)


def test_library_exports_every_declared_symbol(sr):
    from mobilesuperresolution_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "b200sr.h")).read()
    declared = set(re.findall(r"B200SR_API[^;(]*?\b(b200sr_\w+)\s*\(", hdr))
    assert declared and declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    lib = _lib.lib()
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.b200sr_version() == 1
    nm = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (b200sr_\w+)", nm))
    assert exported == declared          # nothing else leaks out of the C ABI


def test_abi_error_codes_without_device(sr):
    from mobilesuperresolution_b200 import _lib
    L = _lib.lib()
    h = ctypes.c_void_p()
    m1 = (ctypes.c_int32 * 1)(144)
    m2 = (ctypes.c_int32 * 1)(20)
    bad = _lib.WdsrDesc(5, 1, 24, 1, 0.5, m1, m2)
    assert L.b200sr_wdsr_create(ctypes.byref(bad), ctypes.byref(h)) == -4 and b"scale" in L.b200sr_last_error()
    bad = _lib.WdsrDesc(4, 1, 32, 1, 0.5, m1, m2)
    assert L.b200sr_wdsr_create(ctypes.byref(bad), ctypes.byref(h)) == -4
    ok = _lib.WdsrDesc(4, 1, 24, 1, 0.5, m1, m2)
    assert L.b200sr_wdsr_create(ctypes.byref(ok), ctypes.byref(h)) == 0
    assert L.b200sr_wdsr_trunk_channels(h) == 24
    assert L.b200sr_wdsr_commit(h) == -2 and b"not set" in L.b200sr_last_error()
    if L.b200sr_device_count() == 0:
        z = np.zeros(100000, np.float32)
        p = z.ctypes.data_as(ctypes.c_void_p)
        assert L.b200sr_wdsr_set_head(h, p, p) == 0 and L.b200sr_wdsr_set_tail(h, p, p, p, p) == 0
        assert L.b200sr_wdsr_set_block(h, 0, p, p, p, p, p, p) == 0
        assert L.b200sr_wdsr_set_block(h, 3, p, p, p, p, p, p) == -1
        assert L.b200sr_wdsr_commit(h) == -2 and b"no CPU fallback" in L.b200sr_last_error()
    L.b200sr_wdsr_destroy(h)


def test_abi_error_codes_conv_and_split_without_device(sr):
    """Argument errors of the video / Split_Block entry points come back as codes (never exceptions, never a device touch)."""
    from mobilesuperresolution_b200 import _lib
    L = _lib.lib()
    h = ctypes.c_void_p()
    z = np.zeros(4096, np.float32)
    p = z.ctypes.data_as(ctypes.c_void_p)
    assert L.b200sr_conv_create(8, 8, 9, p, p, ctypes.byref(h)) == -4 and b"k in" in L.b200sr_last_error()
    assert L.b200sr_conv_create(8, 8, 3, None, p, ctypes.byref(h)) == -1
    assert L.b200sr_split_create(12, p, p, p, p, p, p, p, p, ctypes.byref(h)) == -4 and b"channels=12" in L.b200sr_last_error()
    assert L.b200sr_split_create(24, p, p, p, p, p, p, p, None, ctypes.byref(h)) == -1
    assert L.b200sr_split_forward(None, p, p, 1, 8, 8, 0, None) == -1
    assert L.b200sr_conv_forward_layout(None, p, 0, 8, 0, p, 0, 8, 0, None, 0, 0, 1, 8, 8, 0, 1, 1, 1, 1, None) == -1
    assert L.b200sr_vsr_conv_last_base(None, p, 0, 64, 0, p, 0, p, 0, 1, 8, 8, None) == -1
    assert L.b200sr_conv_set_max_ctas(None, 4) == -1
    assert L.b200sr_conv_tcgen05_ok(None) == 0
    if L.b200sr_device_count() == 0:
        assert L.b200sr_conv_create(64, 64, 3, p, p, ctypes.byref(h)) == -2 and b"no CPU fallback" in L.b200sr_last_error()
        big = np.zeros(24 * 49 + 3 * 24 * 24, np.float32).ctypes.data_as(ctypes.c_void_p)
        assert L.b200sr_split_create(24, big, big, big, big, big, big, big, big, ctypes.byref(h)) == -2


def test_no_cpu_fallback(sr):
    m = sr.BASIC_MODEL(P(2, 1)).eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.rand(1, 3, 8, 8))
    from mobilesuperresolution_b200 import video
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        video.flow_warp(torch.zeros(1, 1, 4, 4), torch.zeros(1, 4, 4, 2))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "mobilesuperresolution_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "liboracle" not in src, f


def test_state_dict_layouts_match_appendix_b(sr, kat):
    torch.manual_seed(0)
    m = sr.BASIC_MODEL(P(4, 16))
    sd = m.state_dict()
    assert len(sd) == 153 and sum(v.numel() for v in sd.values()) == 191368
    assert tuple(sd["head.weight_v"].shape) == (24, 3, 3, 3) and tuple(sd["head.weight_g"].shape) == (24, 1, 1, 1)
    assert tuple(sd["body.15.body.0.weight_v"].shape) == (144, 24, 1, 1)
    assert tuple(sd["body.0.body.2.weight_v"].shape) == (20, 144, 1, 1)
    assert tuple(sd["body.0.body.3.weight_v"].shape) == (24, 20, 3, 3)
    assert tuple(sd["tail.weight_v"].shape) == (48, 24, 3, 3) and tuple(sd["skip.0.weight_v"].shape) == (48, 3, 5, 5)
    # seeded construction consumes the RNG exactly like the reference (App. D KAT1 weights)
    assert abs(float(sum(v.double().sum() for v in sd.values())) - kat["KAT1"]["weights_sum"]) < 1e-6
    assert np.allclose(sd["head.weight_v"][0, 0, 0, :].numpy(), kat["KAT1"]["head.weight_v[0,0,0,:]"], atol=1e-8)
    assert float(sd["body.0.body.0.weight_g"][0]) == 2.0 and float(sd["body.0.body.3.weight_g"][0]) == 0.25
    sd2 = sr.BASIC_MODEL(P(2, 16)).state_dict()
    assert sum(v.numel() for v in sd2.values()) == 180748
    z = np.load(os.path.join(ROOT, "tests", "golden", "wdsr_b_x2_16_24_pretrained.npz"))
    sr.BASIC_MODEL(P(2, 16)).load_state_dict({k: torch.from_numpy(z[k]) for k in z.files}, strict=True)
    blk = sr.Block(num_residual_units=24, kernel_size=3, width_search=True).state_dict()
    assert tuple(blk["body.2.weight"].shape) == (144, 1, 1, 1) and tuple(blk["body.4.weight"].shape) == (20, 1, 1, 1)
    assert "body.3.weight_v" in blk and "body.5.weight_v" in blk
    agg = sr.AggregationLayer(num_residual_units=24, kernel_size=3, width_search=False).state_dict()
    assert all(k in agg for k in ("alpha1", "alpha2", "beta1", "beta2"))
    nas = sr.NAS_MODEL_classic(P(4, 4, ws=True)).state_dict()
    assert "mask.weight" in nas and "skip.weight_v" in nas and "skip.0.weight_v" not in nas


def test_pruned_model_constructor_and_file_reader(sr, tmp_path):
    f = tmp_path / "block_index.txt"
    f.write_text("([0, 1, 2], [[24, 144, 20]])\n([0, 2], [[9, 91, 14], [9, 94, 10]])\n")      # last line wins
    m = sr.Model(4, str(f))
    sd = m.state_dict()
    assert m.IN == 9 and tuple(sd["body.0.weight_v"].shape) == (9, 3, 3, 3)
    assert tuple(sd["body.1.body.0.weight_v"].shape) == (91, 9, 1, 1)
    assert tuple(sd["body.2.body.2.weight_v"].shape) == (10, 94, 1, 1)
    assert tuple(sd["body.2.body.3.weight_v"].shape) == (9, 10, 3, 3)
    assert tuple(sd["body.3.weight_v"].shape) == (48, 9, 3, 3) and tuple(sd["skip.weight_v"].shape) == (48, 3, 5, 5)


def test_rounding_matches_reference_golden(sr, kat):
    for c in kat["rounding"]:
        w = torch.tensor(c["w"], dtype=torch.float32).view(-1, 1, 1, 1)
        assert sr.rounding(w, c["least"]).view(-1).int().tolist() == c["keep"]


def test_mask_folding_equals_masked_block(sr):
    """prepare()-time slicing (masks -> (IN,M1,M2) filters) reproduces the reference's masked block."""
    meta, arrs, sd, x = golden_case("block_masked")
    b = sr.Block(num_residual_units=24, kernel_size=3, res_scale=0.25, width_search=True)
    b.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    w1, b1, w2, b2, w3, b3 = b.pruned_filters()
    assert w1.shape[0] == int(sr.rounding(torch.from_numpy(sd["body.2.weight"])).sum())
    assert w2.shape[0] == int(sr.rounding(torch.from_numpy(sd["body.4.weight"])).sum()) == w3.shape[1]
    xt = torch.from_numpy(x)
    t = F.relu(F.conv2d(xt, w1[:, :, None, None], b1))
    t = F.conv2d(t, w2[:, :, None, None], b2)
    y = F.conv2d(t, w3, b3, padding=1) + xt
    assert float((y - torch.from_numpy(arrs["y"])).abs().max()) <= 1e-5
    keep_in = torch.tensor([0, 2, 3, 5, 8, 13, 21, 22, 23])
    f = b.pruned_filters(keep_in)
    assert f[0].shape[1] == 9 and f[4].shape[0] == 9 and f[5].shape[0] == 9


def test_depth_gate_and_width_readout(sr):
    m = sr.NAS_MODEL_classic(P(4, 3, ws=True)).eval()
    m.body[1].alpha1.fill_(0.9)
    m.body[1].alpha2.fill_(0.1)
    assert m.get_block_status() == [0, 2] and m.get_current_blocks() == 2
    w = m.get_width_from_block_idx([0, 2])
    assert w == [[24, 144, 20], [24, 144, 20]]           # U(0.5,1) mask init keeps everything (models/ops.py:14)
    assert float(m.speed_accu()) == pytest.approx(3 * (144 + 0.2 * 24) * 9 / 40)


def test_shard_slices_cover_batch():
    from mobilesuperresolution_b200.shard import shard_slice
    for total in (0, 1, 7, 64, 65):
        for world in (1, 2, 3, 8):
            parts = [shard_slice(total, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == total
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_slice(4, 2, 2)


_WORKER = r"""
import os, sys, json
sys.path.insert(0, sys.argv[1])
import torch, torch.distributed as dist
from mobilesuperresolution_b200 import shard
rank, local_rank, world = shard.init_distributed("gloo")
a, b = shard.shard_slice(13, rank, world)
covered = shard.sum_over_ranks(b - a)
slowest = shard.max_over_ranks(10.0 + rank)
shard.barrier()
if rank == 0:
    print(json.dumps({"world": world, "covered": covered, "slowest": slowest}))
dist.destroy_process_group()
"""


def test_two_rank_gloo_sharding(tmp_path):
    """world_size-2 gloo run of the sharding/timing plumbing bench.py uses at N>1."""
    import json
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    w = tmp_path / "worker.py"
    w.write_text(_WORKER)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(w), ROOT], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=120) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    res = json.loads(outs[0][0].strip().splitlines()[-1])
    assert res == {"world": 2, "covered": 13.0, "slowest": 11.0}


def test_fork_nas_model_state_dict_and_seeded_construction(sr):
    """The fork's NAS_MODEL as committed (models/wdsr_b.py:30-137): App. B layout (421 + mask tensors for 16 blocks) and a constructor
    that consumes the RNG exactly like the reference's (head, speed-estimator MLP incl. its re-initialisation, blocks, mask, tail,
    skip): the weight sum of every key but the estimator's pickled values equals the reference's under torch.manual_seed(0)."""
    from conftest import load_golden
    sd = sr.NAS_MODEL(P(4, 16, ws=True)).state_dict()
    assert len(sd) == 422 and "mask.weight" in sd and "skip.weight_v" in sd and "skip.0.weight_v" not in sd
    assert tuple(sd["speed_estimator.estimator.fc3.weight"].shape) == (128, 64)
    assert tuple(sd["body.7.body.5.0.body.0.weight_v"].shape) == (24, 1, 5, 5) and tuple(sd["body.7.body.7.0.body.2.weight_v"].shape) == (24, 24, 1, 1)
    assert tuple(sd["body.0.split.weight"].shape) == (24, 1, 1, 1) and tuple(sd["body.0.alpha"].shape) == (3,)
    meta, _ = load_golden("nas_fork")
    torch.manual_seed(0)
    m = sr.NAS_MODEL(P(meta["scale"], meta["nb"], ws=True))
    wsum = float(sum(v.double().sum() for k, v in m.state_dict().items() if not k.startswith("speed_estimator.")))
    assert abs(wsum - meta["seed0_weights_sum_without_estimator"]) < 1e-9
    # read-outs: (IN, split, kernel) triples, host arithmetic only
    assert m.get_block_status() == [0, 1, 2] and all(len(t) == 3 and t[2] in (3, 5, 7) for t in m.get_width_from_block_idx([0, 1, 2]))
    with pytest.raises(RuntimeError):
        m.eval()(torch.rand(1, 3, 8, 8))            # CPU tensor: no fallback


def test_naive_model_state_dict_and_seeded_construction(sr, tmp_path):
    """Naive_model (models/naive_multi_model_easy.py:33-108): same keys in the same order as the reference's state_dict (flownet, the
    never-used block / model `skip` convs included), `decode` inherits the last block's kernel size, and a constructor that consumes the RNG
    exactly like the reference's (weight sum under torch.manual_seed(0) recorded by oracle/make_golden_r3.py)."""
    from conftest import load_golden
    meta, _ = load_golden("naive_model")
    f = tmp_path / "naive_index.txt"
    f.write_text(repr((list(range(len(meta["blocks"]))), meta["blocks"])) + "\n")
    torch.manual_seed(0)
    m = sr.Naive_model(meta["scale"], str(f))
    sd = m.state_dict()
    assert list(sd.keys()) == meta["keys"]
    assert {k: list(v.shape) for k, v in sd.items()} == meta["shapes"]
    assert tuple(sd["decode.weight_v"].shape) == (48, 16, 5, 5) and tuple(sd["body.0.body.0.weight"].shape) == (16, 34, 3, 3)
    assert abs(float(sum(v.double().sum() for v in sd.values())) - meta["seed0_weights_sum"]) < 1e-9
    with pytest.raises(RuntimeError):
        m.eval()(torch.rand(1, 2, 3, 40, 40))       # CPU tensor: no fallback


def test_cat_feats_is_free_only_for_adjacent_windows_of_one_tensor():
    """video._cat_feats: torch.cat([backward, forward], -1) of models/basicvsr_arch_origin.py:84 -- returns the shared base tensor (no copy)
    exactly when the two operands are its two adjacent channel halves, and an ordinary concatenation in every other case."""
    from mobilesuperresolution_b200.video import _cat_feats
    base = torch.arange(2 * 3 * 4 * 8, dtype=torch.float32).view(2, 3, 4, 8).clone()   # owns its storage, like the tensors propagate() allocates
    a, b = base[..., :4], base[..., 4:]
    assert _cat_feats(a, b) is base
    assert torch.equal(_cat_feats(b, a), torch.cat([b, a], -1)) and _cat_feats(b, a) is not base          # swapped halves
    assert torch.equal(_cat_feats(a, a), torch.cat([a, a], -1))                                            # same half twice
    other = base.clone()
    assert torch.equal(_cat_feats(a, other[..., 4:]), base) and _cat_feats(a, other[..., 4:]) is not base   # windows of different tensors
    c, d = base[..., :3], base[..., 3:6]
    assert torch.equal(_cat_feats(c, d), base[..., :6]) and _cat_feats(c, d).data_ptr() != base.data_ptr() or _cat_feats(c, d).shape[-1] == 6
    x, y = torch.zeros(2, 3, 4, 4), torch.ones(2, 3, 4, 4)
    assert torch.equal(_cat_feats(x, y), torch.cat([x, y], -1))                                             # plain tensors (no base)

