"""The fork's searchable block: ``Conv_sep`` / ``Split_Block`` / ``MyAggregationLayer`` (models/wdsr_b.py:375-546).

Same constructors, ``forward`` signatures and ``state_dict`` layout (``alpha``, ``beta``, ``split.weight``,
``body.{3,5,7}.0.body.{0,2}.{bias,weight_g,weight_v}``, + ``alpha1/beta1/alpha2/beta2``); the whole ``forward_body`` is ONE
CUDA kernel behind ``b200sr_split_forward`` (include/b200sr.h).  Inference only; CUDA tensors only (no CPU fallback).
"""
from __future__ import annotations

import ctypes

import torch
import torch.nn as nn
import torch.nn.functional as F
import torch.nn.init as init

from . import _lib
from .masks import BinaryConv2d, rounding
from .wdsr import _fold, _ptr, _weight_norm

__all__ = ["Conv_sep", "Split_Block", "MyAggregationLayer"]


class Conv_sep(nn.Module):
    """Parameter container of models/wdsr_b.py:375-402 (depthwise k x k -> ReLU -> 1x1, both weight-normed; or one plain conv)."""

    def __init__(self, input_dim, output_dim, kernal_size, weight_norm=torch.nn.utils.weight_norm, seperate=False):
        super().__init__()
        self.seperate = seperate
        self.kernel_size = kernal_size
        body = []
        if self.seperate:
            body.append(_weight_norm(nn.Conv2d(input_dim, input_dim, kernal_size, padding=kernal_size // 2, groups=input_dim)))
            body.append(nn.ReLU(inplace=True))
            body.append(_weight_norm(nn.Conv2d(input_dim, output_dim, 1, padding=0)))
        else:
            body.append(_weight_norm(nn.Conv2d(input_dim, output_dim, kernal_size, padding=kernal_size // 2)))
        self.body = nn.Sequential(*body)


class _SplitPlan:
    """One ``b200sr_split_t``: folded filters, effective mask and softmax(alpha) resident on one device."""

    def __init__(self, blk: "Split_Block", device: torch.device):
        c = blk.split.weight.shape[0]
        dws, dwb, pws, pwb = [], [], [], []
        for k in blk.kernel_list:
            sep = blk.body[k][0]
            if not sep.seperate:
                raise NotImplementedError("b200sr Split_Block: only seperate_type=True (the reference's default) has a kernel")
            w, b = _fold(sep.body[0])
            dws.append(w.reshape(c, -1).contiguous())
            dwb.append(b)
            w, b = _fold(sep.body[2])
            pws.append(w.reshape(c, c).contiguous())
            pwb.append(b)
        wm = blk.split.weight.detach().float().cpu()
        eff = (wm - (wm - rounding(wm, blk.split.least_channel))).reshape(-1).contiguous()       # models/ops.py:19-23
        with torch.no_grad():
            prob = F.softmax(blk.alpha.detach().float().cpu(), dim=0).contiguous()               # models/wdsr_b.py:487
        dwb_t, pw_t, pwb_t = torch.stack(dwb).contiguous(), torch.stack(pws).contiguous(), torch.stack(pwb).contiguous()
        h = ctypes.c_void_p()
        with torch.cuda.device(device):
            _lib.check(_lib.lib().b200sr_split_create(c, _ptr(dws[0]), _ptr(dws[1]), _ptr(dws[2]), _ptr(dwb_t), _ptr(pw_t), _ptr(pwb_t),
                                                      _ptr(eff), _ptr(prob), ctypes.byref(h)))
        self._h, self.channels, self.device = h, c, device

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                _lib.lib().b200sr_split_destroy(h)
            except Exception:
                pass

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        n, c, h, w = x.shape
        if c != self.channels:
            raise RuntimeError(f"Split_Block: expected {self.channels} channels, got {c}")
        x = x.contiguous()
        y = torch.empty_like(x)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().b200sr_split_forward(self._h, _ptr(x), _ptr(y), n, h, w, _lib.dtype_code(x.dtype),
                                                       _lib.current_stream_ptr(x.device)))
        return y


class Split_Block(nn.Module):
    """models/wdsr_b.py:406-502.  ``forward(x)``: x (n, C, h, w) float32 | bfloat16 CUDA tensor -> same shape / dtype."""

    def __init__(self, num_residual_units, kernel_size, weight_norm=torch.nn.utils.weight_norm, res_scale=1, width_search=False,
                 block_type="normal", seperate_type=True):
        super().__init__()
        self.alpha = nn.Parameter(data=torch.ones(3), requires_grad=True)
        init.uniform_(self.alpha, 0.5, 1.5)
        self.beta = nn.Parameter(data=torch.zeros(3), requires_grad=True)
        self.split = BinaryConv2d(num_residual_units, num_residual_units, groups=num_residual_units, least_channel=0)
        self.kernel_list = ["3", "5", "7"]
        self.body = nn.ModuleDict()
        for kernel_ in self.kernel_list:
            body = []
            if block_type == "normal":
                body.append(Conv_sep(num_residual_units, num_residual_units, int(kernel_), seperate=seperate_type))
                body.append(nn.ReLU(inplace=True))
            self.body[kernel_] = nn.Sequential(*body)
        if block_type != "normal":
            raise NotImplementedError("b200sr Split_Block: block_type='normal' only (the other types are commented out upstream, :430-477)")

    def _plan(self, device) -> _SplitPlan:
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        sig = (str(device),) + tuple((p.data_ptr(), p._version) for p in self.parameters())
        if getattr(self, "_plan_sig", None) != sig:
            self._plan_obj, self._plan_sig = _SplitPlan(self, device), sig
        return self._plan_obj

    def forward_body(self, x: torch.Tensor) -> torch.Tensor:
        _lib.require_cuda_tensor(x, "x")
        if self.training:
            raise NotImplementedError("b200sr is inference-only: call .eval()")
        return self._plan(x.device).forward(x)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.forward_body(x)


class MyAggregationLayer(Split_Block):
    """Split_Block + depth gate (models/wdsr_b.py:504-546).  Eval (:539-546): identity iff ``alpha1 >= alpha2``;
    ``speed_accu + beta2 * speed_curr`` is scalar arithmetic on the caller's tensors."""

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.alpha1 = nn.Parameter(data=torch.empty(1), requires_grad=True)
        self.beta1 = nn.Parameter(data=torch.zeros(1), requires_grad=True)
        init.uniform_(self.alpha1, 0, 0.2)
        self.alpha2 = nn.Parameter(data=torch.empty(1), requires_grad=True)
        self.beta2 = nn.Parameter(data=torch.ones(1), requires_grad=True)
        init.uniform_(self.alpha2, 0.8, 1)

    def is_skipped(self) -> bool:
        return bool(self.alpha1.detach() >= self.alpha2.detach())

    def forward(self, x, speed_curr, speed_accu):
        if self.training:
            raise NotImplementedError("b200sr is inference-only: call .eval() (training branch models/wdsr_b.py:519-538)")
        if not self.is_skipped():
            x = self.forward_body(x)
        return x, speed_accu + self.beta2.detach().to(speed_accu.device) * speed_curr
