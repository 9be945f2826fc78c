"""torch.nn.functional restatement of the reference forward path (CPU).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.

Every function takes a plain ``state_dict`` (name -> tensor) in the
reference's own layout and replays the op sequence the
reference module executes, citing the reference file:line it follows
(paths relative to /root/reference).  No nn.Module, no hooks: the weight-norm
hook becomes an explicit fold per conv per call, exactly as the reference
recomputes it on every forward.

Parity: PINNED.  ``oracle/make_golden.py`` asserts max-abs 0.0 between these
functions and the imported reference modules, and writes the golden vectors
the CPU tests re-check on boxes without /root/reference.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


# ----------------------------------------------------------------------------
# weight norm + masks
# ----------------------------------------------------------------------------
def weight_norm_fold(g: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """W[o] = g[o] * v[o] / ||v[o]||_2 over (in, kh, kw).

    Reference: ``torch.nn.utils.weight_norm`` (dim=0) applied to every conv,
    models/basic_wdsr_b.py:23,32,55,68,108,119,129; the hook calls
    ``torch._weight_norm(v, g, 0)`` before each forward.
    """
    return torch._weight_norm(v, g, 0)


def rounding(weight: torch.Tensor, least_channel: int = 8) -> torch.Tensor:
    """Binary keep-mask of a BinaryConv2d weight.  models/ops.py:33-43."""
    keep = (weight >= 0.5).float()
    if least_channel > 0:
        top, _ = torch.topk(weight, least_channel, dim=0)
        keep_topk = (weight >= top[-1]).float()
        return keep if torch.sum(keep) >= least_channel else keep_topk
    return keep


def binary_mask_apply(x: torch.Tensor, mask_weight: torch.Tensor, least_channel: int = 8) -> torch.Tensor:
    """BinaryConv2d.forward value: depthwise 1x1 with weight w-(w-m).  models/ops.py:18-26."""
    w = mask_weight.detach()
    m = rounding(w, least_channel)
    eff = mask_weight - (w - m)
    return F.conv2d(x, eff, None, 1, 0, 1, x.shape[1])


def _wn_conv(sd: SD, prefix: str, x: torch.Tensor, pad: int) -> torch.Tensor:
    w = weight_norm_fold(sd[prefix + "weight_g"], sd[prefix + "weight_v"])
    return F.conv2d(x, w, sd[prefix + "bias"], padding=pad)


# ----------------------------------------------------------------------------
# WDSR-B image path
# ----------------------------------------------------------------------------
def wdsr_block(sd: SD, prefix: str, x: torch.Tensor, idx: Sequence[int] = (0, 2, 3)) -> torch.Tensor:
    """x + conv3x3(conv1x1(relu(conv1x1(x)))).

    Reference: Block, models/basic_wdsr_b.py:96-144 (forward :142-144);
    identical to models/wdsr_b.py:253-319 with width_search=False and to
    export_onnx.py:91-114.  ``idx`` are the Sequential indices of the three
    convs (0,2,3 classic; 0,3,5 with width_search masks).
    """
    e, r, c = idx
    t = F.relu(_wn_conv(sd, f"{prefix}body.{e}.", x, 0))
    t = _wn_conv(sd, f"{prefix}body.{r}.", t, 0)
    t = _wn_conv(sd, f"{prefix}body.{c}.", t, 1)
    return t + x


def wdsr_block_masked(sd: SD, prefix: str, x: torch.Tensor) -> torch.Tensor:
    """Block(width_search=True): masks after ReLU and after the reduce conv.

    Reference: models/wdsr_b.py:281-303 (Sequential order: conv, ReLU,
    BinaryConv2d(144), conv, BinaryConv2d(20), conv) and forward :316-319.
    """
    t = F.relu(_wn_conv(sd, f"{prefix}body.0.", x, 0))
    t = binary_mask_apply(t, sd[f"{prefix}body.2.weight"])
    t = _wn_conv(sd, f"{prefix}body.3.", t, 0)
    t = binary_mask_apply(t, sd[f"{prefix}body.4.weight"])
    t = _wn_conv(sd, f"{prefix}body.5.", t, 1)
    return t + x


def split_block(sd: SD, prefix: str, x: torch.Tensor) -> torch.Tensor:
    """The fork's searchable block body, ``Split_Block.forward_body`` (models/wdsr_b.py:482-496):

        x1 = split(x); x2 = x - x1; x3 = x2 + sum_k softmax(alpha)_k * relu(PW_k(relu(DW_k(x1)))) + x1; out = x2 + split(x3)

    ``split`` = BinaryConv2d(least_channel=0) (models/wdsr_b.py:424, models/ops.py:18-26); DW_k = weight-normed depthwise
    k x k (k = 3, 5, 7), PW_k = weight-normed 1x1 (``Conv_sep`` models/wdsr_b.py:375-402); keys ``body.{3,5,7}.0.body.{0,2}.*``.
    """
    c = x.shape[1]
    mw = sd[prefix + "split.weight"]
    x1 = binary_mask_apply(x, mw, 0)
    x2 = x - x1
    x3 = torch.clone(x2)
    pro = F.softmax(sd[prefix + "alpha"], dim=0)
    for i, k in enumerate((3, 5, 7)):
        q = f"{prefix}body.{k}.0.body."
        wd = weight_norm_fold(sd[q + "0.weight_g"], sd[q + "0.weight_v"])
        t = F.relu(F.conv2d(x1, wd, sd[q + "0.bias"], padding=k // 2, groups=c))
        wp = weight_norm_fold(sd[q + "2.weight_g"], sd[q + "2.weight_v"])
        t = F.relu(F.conv2d(t, wp, sd[q + "2.bias"]))
        x3 = x3 + t * pro[i]
    x3 = x3 + x1
    return x2 + binary_mask_apply(x3, mw, 0)


def my_aggregation_layer(sd: SD, prefix: str, x: torch.Tensor) -> torch.Tensor:
    """``MyAggregationLayer.forward`` eval branch (models/wdsr_b.py:539-546): identity when alpha1 >= alpha2."""
    if float(sd[prefix + "alpha1"]) >= float(sd[prefix + "alpha2"]):
        return x
    return split_block(sd, prefix, x)


def count_blocks(sd: SD, prefix: str = "body.") -> int:
    ids = {int(k[len(prefix):].split(".")[0]) for k in sd if k.startswith(prefix) and k[len(prefix)].isdigit()}
    return len(ids)


def basic_model_forward(sd: SD, x: torch.Tensor, scale: int, image_mean: float = 0.5) -> torch.Tensor:
    """BASIC_MODEL.forward, models/basic_wdsr_b.py:85-93."""
    nb = count_blocks(sd, "body.")
    x = x - image_mean                                      # :86
    y = _wn_conv(sd, "head.", x, 1)                         # :87
    for b in range(nb):                                     # :88-89
        y = wdsr_block(sd, f"body.{b}.", y)
    skip_prefix = "skip.0." if "skip.0.weight_v" in sd else "skip."
    t = _wn_conv(sd, "tail.", y, 1)
    if skip_prefix + "weight_v" in sd:                      # skip only if num_inputs != num_outputs (:66)
        t = t + _wn_conv(sd, skip_prefix, x, 2)             # :90
    else:
        t = t + x
    if scale > 1:
        t = F.pixel_shuffle(t, scale)                       # :91
    return t + image_mean                                   # :92


def pruned_model_forward(sd: SD, x: torch.Tensor, scale: int, image_mean: float = 0.5) -> torch.Tensor:
    """export_onnx.Model.forward, export_onnx.py:59-79.

    Layout: body.0 head, body.{1..nb} blocks, body.{nb+1} tail, skip.* ; the
    forward does NOT add image_mean back (:62-63,79).
    """
    ids = sorted({int(k.split(".")[1]) for k in sd if k.startswith("body.")})
    head_i, tail_i = ids[0], ids[-1]
    x = x - image_mean
    y = _wn_conv(sd, f"body.{head_i}.", x, 1)
    for b in ids[1:-1]:
        y = wdsr_block(sd, f"body.{b}.", y)
    t = _wn_conv(sd, f"body.{tail_i}.", y, 1) + _wn_conv(sd, "skip.", x, 2)
    if scale > 1:
        t = F.pixel_shuffle(t, scale)
    return t


def supernet_classic_forward(sd: SD, x: torch.Tensor, scale: int, image_mean: float = 0.5,
                             width_search: bool = True) -> torch.Tensor:
    """NAS_MODEL.forward image path with the classic AggregationLayer body (eval).

    Reference: models/wdsr_b.py:107-137 (global mask before every block and
    before the tail, :116,:119) + AggregationLayer eval branch :358-365
    (block is identity iff alpha1 >= alpha2).  speed_accu is a scalar side
    output and is not restated here.
    """
    nb = count_blocks(sd, "body.")
    x = x - image_mean
    y = _wn_conv(sd, "head.", x, 1)
    for b in range(nb):
        p = f"body.{b}."
        if width_search:
            y = binary_mask_apply(y, sd["mask.weight"])
        a1, a2 = sd.get(p + "alpha1"), sd.get(p + "alpha2")
        if a1 is not None and bool(a1 >= a2):
            continue
        y = wdsr_block_masked(sd, p, y) if width_search else wdsr_block(sd, p, y)
    if width_search:
        y = binary_mask_apply(y, sd["mask.weight"])
    t = _wn_conv(sd, "tail.", y, 1) + _wn_conv(sd, "skip.", x, 2)
    if scale > 1:
        t = F.pixel_shuffle(t, scale)
    return t + image_mean


# ----------------------------------------------------------------------------
# video path: flow_warp, SPyNet, BasicVSR
# ----------------------------------------------------------------------------
def flow_warp(x: torch.Tensor, flow: torch.Tensor, interp_mode: str = "bilinear",
              padding_mode: str = "zeros", align_corners: bool = True) -> torch.Tensor:
    """models/spynet_arch.py:98-129."""
    assert x.size()[-2:] == flow.size()[1:3]
    _, _, h, w = x.size()
    gy, gx = torch.meshgrid(torch.arange(0, h).type_as(x), torch.arange(0, w).type_as(x), indexing="ij")
    grid = torch.stack((gx, gy), 2).float()
    v = grid + flow
    vx = 2.0 * v[:, :, :, 0] / max(w - 1, 1) - 1.0
    vy = 2.0 * v[:, :, :, 1] / max(h - 1, 1) - 1.0
    return F.grid_sample(x, torch.stack((vx, vy), dim=3), mode=interp_mode,
                         padding_mode=padding_mode, align_corners=align_corners)


_SPY_IDX = (0, 2, 4, 6, 8)


def spynet_basic_module(sd: SD, level: int, t: torch.Tensor, prefix: str = "") -> torch.Tensor:
    """BasicModule.forward, models/spynet_arch.py:10-25: 7x7 convs 8-32-64-32-16-2, ReLU between."""
    for j, i in enumerate(_SPY_IDX):
        p = f"{prefix}basic_module.{level}.basic_module.{i}."
        t = F.conv2d(t, sd[p + "weight"], sd[p + "bias"], padding=3)
        if j < 4:
            t = F.relu(t)
    return t


def spynet_process(sd: SD, ref: torch.Tensor, supp: torch.Tensor, prefix: str = "") -> torch.Tensor:
    """SpyNet.process, models/spynet_arch.py:49-79."""
    mean, std = sd[prefix + "mean"], sd[prefix + "std"]
    refs = [(ref - mean) / std]
    supps = [(supp - mean) / std]
    for _ in range(5):
        refs.insert(0, F.avg_pool2d(refs[0], 2, 2, count_include_pad=False))
        supps.insert(0, F.avg_pool2d(supps[0], 2, 2, count_include_pad=False))
    flow = refs[0].new_zeros([refs[0].size(0), 2, int(math.floor(refs[0].size(2) / 2.0)),
                              int(math.floor(refs[0].size(3) / 2.0))])
    for level in range(len(refs)):
        up = F.interpolate(flow, scale_factor=2, mode="bilinear", align_corners=True) * 2.0
        if up.size(2) != refs[level].size(2):
            up = F.pad(up, [0, 0, 0, 1], mode="replicate")
        if up.size(3) != refs[level].size(3):
            up = F.pad(up, [0, 1, 0, 0], mode="replicate")
        warped = flow_warp(supps[level], up.permute(0, 2, 3, 1), padding_mode="border")
        flow = spynet_basic_module(sd, level, torch.cat([refs[level], warped, up], 1), prefix) + up
    return flow


def spynet_forward(sd: SD, ref: torch.Tensor, supp: torch.Tensor, prefix: str = "") -> torch.Tensor:
    """SpyNet.forward, models/spynet_arch.py:81-96."""
    assert ref.size() == supp.size()
    h, w = ref.size(2), ref.size(3)
    w_up = math.floor(math.ceil(w / 32.0) * 32.0)
    h_up = math.floor(math.ceil(h / 32.0) * 32.0)
    ref = F.interpolate(ref, size=(h_up, w_up), mode="bilinear", align_corners=False)
    supp = F.interpolate(supp, size=(h_up, w_up), mode="bilinear", align_corners=False)
    flow = F.interpolate(spynet_process(sd, ref, supp, prefix), size=(h, w), mode="bilinear", align_corners=False)
    flow[:, 0] *= float(w) / float(w_up)
    flow[:, 1] *= float(h) / float(h_up)
    return flow


def vsr_get_flow(sd: SD, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """BasicVSR.get_flow, models/basicvsr_arch_origin.py:42-51 (== basicvsr_arch.py:45-54)."""
    b, n, c, h, w = x.size()
    x1 = x[:, :-1].reshape(-1, c, h, w)
    x2 = x[:, 1:].reshape(-1, c, h, w)
    fb = spynet_forward(sd, x1, x2, "spynet.").view(b, n - 1, 2, h, w)
    ff = spynet_forward(sd, x2, x1, "spynet.").view(b, n - 1, 2, h, w)
    return ff, fb


def vsr_trunk(sd: SD, prefix: str, t: torch.Tensor, num_block: int) -> torch.Tensor:
    """ConvResidualBlocks, models/basicvsr_arch_origin.py:98-137."""
    t = F.leaky_relu(F.conv2d(t, sd[prefix + "main.0.weight"], sd[prefix + "main.0.bias"], padding=1), 0.1)
    for k in range(num_block):
        p = f"{prefix}main.2.{k}."
        o = F.relu(F.conv2d(t, sd[p + "conv1.weight"], sd[p + "conv1.bias"], padding=1))
        o = F.conv2d(o, sd[p + "conv2.weight"], sd[p + "conv2.bias"], padding=1)
        t = t + o
    return t


def _count_trunk_blocks(sd: SD, prefix: str) -> int:
    return len({int(k.split(".")[3]) for k in sd if k.startswith(prefix + "main.2.")})


def vsr_propagate(sd: SD, x: torch.Tensor, ff: torch.Tensor, fb: torch.Tensor, num_feat: int
                  ) -> Tuple[List[torch.Tensor], List[torch.Tensor]]:
    """The two recurrent loops, models/basicvsr_arch_origin.py:61-82, returning per-frame features."""
    b, n, _, h, w = x.size()
    nb = _count_trunk_blocks(sd, "backward_trunk.")
    back: List[torch.Tensor] = []
    feat = x.new_zeros(b, num_feat, h, w)
    for i in range(n - 1, -1, -1):
        if i < n - 1:
            feat = flow_warp(feat, fb[:, i].permute(0, 2, 3, 1))
        feat = vsr_trunk(sd, "backward_trunk.", torch.cat([x[:, i], feat], 1), nb)
        back.insert(0, feat)
    fwd: List[torch.Tensor] = []
    feat = torch.zeros_like(feat)
    for i in range(n):
        if i > 0:
            feat = flow_warp(feat, ff[:, i - 1].permute(0, 2, 3, 1))
        feat = vsr_trunk(sd, "forward_trunk.", torch.cat([x[:, i], feat], 1), nb)
        fwd.append(feat)
    return back, fwd


def basicvsr_origin_forward(sd: SD, x: torch.Tensor, height: int, weight: int) -> torch.Tensor:
    """BasicVSR_origin.forward, models/basicvsr_arch_origin.py:53-96."""
    num_feat = sd["fusion.weight"].shape[0]
    ff, fb = vsr_get_flow(sd, x)
    back, fwd = vsr_propagate(sd, x, ff, fb, num_feat)
    outs = []
    lr = lambda t: F.leaky_relu(t, 0.1)
    for i in range(x.size(1)):
        o = torch.cat([back[i], fwd[i]], 1)
        o = lr(F.conv2d(o, sd["fusion.weight"], sd["fusion.bias"]))
        o = lr(F.pixel_shuffle(F.conv2d(o, sd["upconv1.weight"], sd["upconv1.bias"], padding=1), 2))
        o = lr(F.pixel_shuffle(F.conv2d(o, sd["upconv2.weight"], sd["upconv2.bias"], padding=1), 2))
        o = lr(F.conv2d(o, sd["conv_hr.weight"], sd["conv_hr.bias"], padding=1))
        o = F.conv2d(o, sd["conv_last.weight"], sd["conv_last.bias"], padding=1)
        o = o + F.interpolate(x[:, i], scale_factor=4, mode="bilinear", align_corners=False)
        o = F.interpolate(o, size=(height, weight), mode="bilinear")
        outs.append(o)
    return torch.stack(outs, dim=1)


def basicvsr_fork_forward(sd: SD, x: torch.Tensor, height: int, weight: int) -> torch.Tensor:
    """The fork's BasicVSR.forward, models/basicvsr_arch.py:56-105 (runs as committed only for num_feat == 3): SPyNet flows, the two
    propagation loops, then per frame lrelu(fusion 1x1 (2nf -> 2nf)) -> ConvTranspose2d(2nf, nf, 5, stride 4) -> bilinear resize to
    (height, weight); ``out += base`` with the bilinear base of the 3-channel frame."""
    num_feat = sd["backward_trunk.main.0.weight"].shape[0]
    ff, fb = vsr_get_flow(sd, x)
    back, fwd = vsr_propagate(sd, x, ff, fb, num_feat)
    outs = []
    for i in range(x.size(1)):
        o = F.leaky_relu(F.conv2d(torch.cat([back[i], fwd[i]], 1), sd["fusion.weight"], sd["fusion.bias"]), 0.1)
        o = F.conv_transpose2d(o, sd["conv_last.weight"], sd["conv_last.bias"], stride=4)
        o = F.interpolate(o, size=(height, weight), mode="bilinear")
        o = o + F.interpolate(x[:, i], size=(height, weight), mode="bilinear", align_corners=False)
        outs.append(o)
    return torch.stack(outs, dim=1)


def nas_fork_speed_curr(sd: SD, prefix: str) -> torch.Tensor:
    """BlockBSpeedEstimator.estimateByMyMask, speed_models/speed_estimator.py:57-84 (the MLP call is commented out upstream; the
    analytic proxy uses ``rounding`` with its DEFAULT least_channel = 8 for both masks)."""
    channels = torch.cat([rounding(sd["mask.weight"]).sum().unsqueeze(0), rounding(sd[prefix + "split.weight"]).sum().unsqueeze(0)])
    output = 0
    kernels = torch.Tensor([3, 5, 7])
    for i in range(3):
        output = output + ((channels[1] + 0.2 * channels[0]) * ((kernels[i] * kernels[i]).unsqueeze(0)) * sd[prefix + "alpha"][i]) / 40
    return output


def nas_fork_forward(sd: SD, x: torch.Tensor, scale: int, image_mean: float = 0.5) -> Tuple[torch.Tensor, torch.Tensor]:
    """The fork's NAS_MODEL.forward as committed (width_search=True), models/wdsr_b.py:105-137, eval:
    y = head(x - mean); per block: y = mask(y); y = MyAggregationLayer(y) (identity when alpha1 >= alpha2, :539-546) and
    speed_accu += beta2 * speed_curr; y = mask(y); out = shuffle(tail(y) + skip(x - mean)) + mean.  Returns (out, speed_accu)."""
    x0 = x - image_mean
    y = _wn_conv(sd, "head.", x0, 1)
    speed = x0.new_zeros(1)
    for i in range(count_blocks(sd)):
        p = f"body.{i}."
        cur = nas_fork_speed_curr(sd, p)
        y = binary_mask_apply(y, sd["mask.weight"])
        y = my_aggregation_layer(sd, p, y)
        speed = speed + sd[p + "beta2"] * cur
    y = binary_mask_apply(y, sd["mask.weight"])
    y = _wn_conv(sd, "tail.", y, 1) + _wn_conv(sd, "skip.", x0, 2)
    return F.pixel_shuffle(y, scale) + image_mean, speed


# ----------------------------------------------------------------------------
# parity metrics
# ----------------------------------------------------------------------------
def mvvsr_forward(sd: SD, x_: torch.Tensor, height: int = 1080, weight: int = 1920) -> torch.Tensor:
    """MotionVectorVSR.forward, models/mvvsr_arch.py:56-109: flows come with the input (channels 3:5, codec motion vectors;
    forward = mv[:, 1:], backward = -forward), BasicVSR's two propagation loops, then per frame
    lrelu(fusion 1x1 (2nf -> 2nf)) -> ConvTranspose2d(2nf, 3, 5, stride 4) -> bilinear resize to (height, weight) + bilinear base."""
    x, mv = x_[:, :, :3], x_[:, :, 3:]
    ff = mv[:, 1:]
    fb = ff * (-1)
    num_feat = sd["backward_trunk.main.0.weight"].shape[0]
    back, fwd = vsr_propagate(sd, x, ff, fb, num_feat)
    outs = []
    for i in range(x.size(1)):
        o = F.leaky_relu(F.conv2d(torch.cat([back[i], fwd[i]], 1), sd["fusion.weight"], sd["fusion.bias"]), 0.1)
        o = F.conv_transpose2d(o, sd["conv_last.weight"], sd["conv_last.bias"], stride=4)
        o = F.interpolate(o, size=(height, weight), mode="bilinear")
        o = o + F.interpolate(x[:, i], size=(height, weight), mode="bilinear", align_corners=False)
        outs.append(o)
    return torch.stack(outs, dim=1)


def psnr_db(y: torch.Tensor, ref: torch.Tensor, peak: float | None = None) -> float:
    """10*log10(peak^2/MSE); peak defaults to the reference's dynamic range."""
    y, ref = y.double(), ref.double()
    if peak is None:
        peak = float(ref.max() - ref.min())
    mse = float(((y - ref) ** 2).mean())
    return float("inf") if mse == 0 else 10.0 * math.log10(peak * peak / mse)


def naive_model_forward(sd: SD, x: torch.Tensor, num_blocks: int) -> torch.Tensor:
    """Naive_model.forward, models/naive_multi_model_easy.py:110-147 (scale 4): flows between consecutive frames (:114-116), per frame
    encode (weight-normed 3x3), block 0 on cat(flow, warp(previous encode output), current) -- frame 0: zero flow and its own features
    (:123-127) -- then conv-ReLU-conv residual blocks (:137, Block :157-183 without weight norm), decode (weight-normed), PixelShuffle(4)
    plus the x4 bilinear base of the frame (:140-144).  The model's 5x5 `skip` conv and the blocks' 1x1 `skip` are never applied."""
    B, N, C, H, W = x.shape
    if N > 1:
        lqs_1 = x[:, :-1].reshape(-1, C, H, W)
        lqs_2 = x[:, 1:].reshape(-1, C, H, W)
        flows_forward = spynet_forward(sd, lqs_2, lqs_1, "flownet.").view(B, N - 1, 2, H, W)
    kenc = sd["encode.weight_v"].shape[-1]
    kdec = sd["decode.weight_v"].shape[-1]
    outs, pre = [], None
    for i in range(N):
        xi = x[:, i]
        x_ = _wn_conv(sd, "encode.", xi, kenc // 2)
        for b in range(num_blocks):
            if b == 0:
                if i == 0:
                    x_warp, flow, pre = x_, torch.zeros(B, 2, H, W, dtype=x.dtype), x_
                else:
                    x_pre, pre = pre, x_
                    flow = flows_forward[:, i - 1]
                    x_warp = flow_warp(x_pre, flow.permute(0, 2, 3, 1))
                x_c = torch.cat((flow, x_warp, x_), dim=1)
            else:
                x_c = x_
            w0, w2 = sd[f"body.{b}.body.0.weight"], sd[f"body.{b}.body.2.weight"]
            t = F.relu(F.conv2d(x_c, w0, sd[f"body.{b}.body.0.bias"], padding=w0.shape[-1] // 2))
            x_ = F.conv2d(t, w2, sd[f"body.{b}.body.2.bias"], padding=w2.shape[-1] // 2) + x_
        base = F.interpolate(xi, scale_factor=4, mode="bilinear", align_corners=False)
        outs.append(F.pixel_shuffle(_wn_conv(sd, "decode.", x_, kdec // 2), 4) + base)
    return torch.stack(outs, dim=1)
