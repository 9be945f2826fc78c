"""The fork's searchable supernet as committed: ``NAS_MODEL`` of models/wdsr_b.py:30-137.

Same constructor (``params``: image_mean, num_channels, scale, num_blocks, num_residual_units, width_search, pretrained), the same
``state_dict`` (``head.*``, ``tail.*``, ``skip.*``, ``mask.weight``, ``speed_estimator.estimator.fc{1,2,3,6,7,8}.*``,
per block ``body.N.{alpha,beta,alpha1,beta1,alpha2,beta2,split.weight}`` and ``body.N.body.{3,5,7}.0.body.{0,2}.*``) and the same
``forward(x) -> (sr, speed_accu)`` (:105-137):

    y = head(x - mean)
    for block in body:  y = mask(y);  y = block(y) unless its depth gate says skip (alpha1 >= alpha2, :539-546)
    y = mask(y);  sr = shuffle(tail(y) + skip(x - mean)) + mean

runs as ONE C-ABI call, ``b200sr_nas_forward``: tcgen05 / FFMA head -> fused ``Split_Block`` kernels of the KEPT blocks (the global
mask is each block's pre-mask, fused into its load) -> fused tail (the last mask is folded into the tail filter's input channels).
``speed_accu`` is the reference's analytic latency proxy ``estimateByMyMask`` (speed_models/speed_estimator.py:57-84) replayed on the
host in the same float32 operation order; the pickled phone-latency MLP is a parameter container only (its ``estimator(...)`` call is
commented out upstream, :71-75).  Like the reference, the model only runs with ``width_search=True`` (``self.mask`` is used
unconditionally at :116: without it the reference raises ``AttributeError``, and so does this class).

The upstream classic-body supernet (``AggregationLayer`` blocks) stays available as ``wdsr.NAS_MODEL_classic``.
"""
from __future__ import annotations

import ctypes
import math
from typing import List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F
import torch.nn.init as init

from . import _lib
from .masks import BinaryConv2d, rounding
from .split import MyAggregationLayer, _SplitPlan
from .wdsr import WdsrPlan, _fold, _PlanCacheMixin, _ptr, _weight_norm

__all__ = ["NAS_MODEL", "BlockBSpeedEstimator", "ConvBlockModel"]


class ConvBlockModel(nn.Module):
    """Parameter container of speed_models/SpeedModel.py:9-62 (the phone-latency MLP, 3 -> 32 -> 64 -> 128 -> 64 -> 32 -> 1).  Built and
    initialised like the reference so that seeded construction consumes the RNG identically; the reference then overwrites the values
    from a pickled ``.pt`` (absent here: load them through ``load_state_dict``).  Never evaluated on the forward path."""

    def __init__(self, num_feat=3):
        super().__init__()
        self.fc1 = nn.Linear(num_feat, 32)
        self.fc2 = nn.Linear(32, 64)
        self.fc3 = nn.Linear(64, 128)
        self.fc6 = nn.Linear(128, 64)
        self.fc7 = nn.Linear(64, 32)
        self.fc8 = nn.Linear(32, 1)
        for m in self.modules():                      # _initialize_weights, speed_models/SpeedModel.py:41-52
            if isinstance(m, nn.Linear):
                nn.init.kaiming_normal_(m.weight)
                m.weight.data *= 0.1
                m.bias.data.fill_(0)
        for p in self.parameters():                   # frozen_layer, :58-60
            p.requires_grad = False


class BlockBSpeedEstimator(nn.Module):
    """speed_models/speed_estimator.py:8-84: holder of the MLP + the analytic proxies.  ``estimateByMyMask`` in the reference's own
    float32 operation order (the MLP call is commented out upstream)."""

    def __init__(self, type):
        super().__init__()
        self.estimator = ConvBlockModel(3).eval()
        self.type = type

    @staticmethod
    def get_unmask_number(m: BinaryConv2d) -> torch.Tensor:
        return rounding(m.weight.detach().float().cpu()).sum().unsqueeze(0)           # default least_channel = 8, as the reference

    @torch.no_grad()
    def estimateByMyMask(self, module: nn.Module, block_mask: nn.Module) -> torch.Tensor:
        channels = torch.cat([self.get_unmask_number(block_mask), self.get_unmask_number(module.split)])
        output = 0
        kernels = torch.Tensor([3, 5, 7])
        alpha = module.alpha.detach().float().cpu()
        for i in range(3):
            output = output + ((channels[1] + 0.2 * channels[0]) * ((kernels[i] * kernels[i]).unsqueeze(0)) * alpha[i]) / 40
        return output


class _NasPlan:
    """Folded head / tail + the kept blocks' Split_Block plans, resident on one device; one C-ABI call per forward."""

    def __init__(self, model: "NAS_MODEL", device: torch.device):
        wm = model.mask.weight.detach().float().cpu()
        gm = (wm - (wm - rounding(wm, model.mask.least_channel))).reshape(-1).contiguous()      # BinaryConv2d forward weight, models/ops.py:19-23
        hw, hb = _fold(model.head)
        tw, tb = _fold(model.tail)
        tw = (tw * gm.view(1, -1, 1, 1)).contiguous()                                            # tail(mask(y)): the mask scales the tail's input channels
        self.wdsr = WdsrPlan(model.scale, int(hw.shape[0]), [], (hw, hb), (tw, tb), _fold(model.skip), True, model.image_mean, device)
        self.blocks: List[_SplitPlan] = []
        for m in model.body:
            if m.is_skipped():
                continue
            sp = _SplitPlan(m, device)
            _lib.check(_lib.lib().b200sr_split_set_premask(sp._h, _ptr(gm)))
            self.blocks.append(sp)
        self._handles = (ctypes.c_void_p * max(1, len(self.blocks)))(*[b._h for b in self.blocks])
        self.device, self.scale = device, model.scale
        self._ws: Optional[torch.Tensor] = None
        # speed_accu (models/wdsr_b.py:110-119, 544-545): scalar host arithmetic in the reference's order
        acc = torch.zeros(1)
        for m in model.body:
            acc = acc + m.beta2.detach().float().cpu() * model.speed_estimator.estimateByMyMask(m, model.mask)
        self.speed_accu = acc.to(device)

    def launches_per_forward(self) -> int:
        return self.wdsr.launches_per_forward()

    def forward(self, x: torch.Tensor, precision: str, out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
        _lib.require_cuda_tensor(x, "input")
        if x.dim() != 4 or x.shape[1] != 3:
            raise RuntimeError(f"expected input of shape (N,3,H,W), got {tuple(x.shape)}")
        x = x.contiguous()
        n, _, h, w = x.shape
        s = self.scale
        out = torch.empty((n, 3, s * h, s * w), dtype=out_dtype or x.dtype, device=x.device)
        if n == 0 or h == 0 or w == 0:
            return out
        L = _lib.lib()
        prec = _lib.precision_code(precision)
        need = L.b200sr_nas_workspace_bytes(self.wdsr.handle, n, h, w, prec)
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(L.b200sr_nas_forward(self.wdsr.handle, self._handles, len(self.blocks), _ptr(x), _lib.dtype_code(x.dtype), _ptr(out),
                                            _lib.dtype_code(out.dtype), n, h, w, prec, _ptr(self._ws), self._ws.numel(),
                                            _lib.current_stream_ptr(self.device)))
        return out


class NAS_MODEL(nn.Module, _PlanCacheMixin):
    """models/wdsr_b.py:30-137 (the fork).  ``forward(x)`` -> ``(sr, speed_accu)``; inference only, CUDA tensors only."""

    def __init__(self, params):
        super().__init__()
        self.image_mean = params.image_mean
        if params.num_channels != 3:
            raise NotImplementedError("b200sr kernels are specialised for 3-channel (RGB) images")
        scale = params.scale
        self.scale = scale
        self.num_blocks = params.num_blocks
        self.num_residual_units = nru = params.num_residual_units
        self.remain_blocks = params.num_blocks
        self.width_search = params.width_search
        self.idx_kernel = [3, 5, 7]
        num_outputs = scale * scale * params.num_channels

        conv = _weight_norm(nn.Conv2d(params.num_channels, nru, 3, padding=1))
        init.ones_(conv.weight_g)
        init.zeros_(conv.bias)
        self.head = conv
        self.speed_estimator = BlockBSpeedEstimator("mask" if params.width_search else "channel").eval()
        self.body = nn.ModuleList(
            [MyAggregationLayer(num_residual_units=nru, kernel_size=3, res_scale=1 / math.sqrt(params.num_blocks),
                                width_search=params.width_search) for _ in range(params.num_blocks)])
        if self.width_search:
            self.mask = BinaryConv2d(in_channels=nru, out_channels=nru, groups=nru)
        conv = _weight_norm(nn.Conv2d(nru, num_outputs, 3, padding=1))
        init.ones_(conv.weight_g)
        init.zeros_(conv.bias)
        self.tail = conv
        conv = _weight_norm(nn.Conv2d(params.num_channels, num_outputs, 5, padding=2))
        init.ones_(conv.weight_g)
        init.zeros_(conv.bias)
        self.skip = conv
        self.shuf = nn.Sequential(*([nn.PixelShuffle(scale)] if scale > 1 else []))
        if getattr(params, "pretrained", False):
            raise NotImplementedError("load_pretrained(): load a state_dict explicitly")

    # ---- search read-outs (host side) ------------------------------------------------------------------
    @torch.no_grad()
    def get_current_blocks(self) -> int:
        return int(sum(1 for m in self.body if not m.is_skipped()))

    @torch.no_grad()
    def get_block_status(self) -> List[int]:
        """:149-158 (softmax of the two gate logits is monotone: same comparison)."""
        out = []
        for idx, m in enumerate(self.body):
            a1, a2 = F.softmax(torch.stack([m.alpha1.detach().float().cpu(), m.alpha2.detach().float().cpu()], dim=0), dim=0)
            if a1 < a2:
                out.append(idx)
        return out

    @torch.no_grad()
    def get_width_from_block_idx(self, remain_block_idx) -> List[List[int]]:
        """:160-183: ``(IN, split, kernel)`` per kept block -- the fork's triple (consumed by its ``result_net``), NOT export_onnx's
        ``(IN, M1, M2)``."""
        mw = self.mask.weight.detach().float().cpu()
        all_width = []
        for idx, m in enumerate(self.body):
            if idx in remain_block_idx:
                sw = m.split.weight.detach().float().cpu()
                best = self.idx_kernel[int(torch.max(m.alpha.detach().float().cpu(), 0)[1])]
                all_width.append([int(rounding(mw).sum()), int((rounding(mw) * rounding(sw)).sum()), best])
        return all_width

    @torch.no_grad()
    def get_mask_weight(self):
        return self.mask.weight.data

    # ---- forward -------------------------------------------------------------------------------------------
    def _build_plan(self, device) -> _NasPlan:
        return _NasPlan(self, device)

    def launches_per_forward(self) -> int:
        return self._plan.launches_per_forward()

    def forward(self, x: torch.Tensor):
        if self.training:
            raise NotImplementedError("b200sr is inference-only: call .eval() (training branch models/wdsr_b.py:519-538)")
        _lib.require_cuda_tensor(x, "input")
        self.mask                                   # width_search=False: AttributeError, as models/wdsr_b.py:116 raises it
        plan = self._get_plan(x.device)
        return plan.forward(x, self.precision), plan.speed_accu.clone()
