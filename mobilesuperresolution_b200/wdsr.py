"""WDSR-B image super-resolution: drop-in mirrors of the reference nn.Modules over the B200 C ABI.

Same constructors, ``forward`` signatures and ``state_dict`` layout as the reference
(SURVEY.md App. B); the forward runs hand-written sm_100a kernels through ``include/b200sr.h``:

* ``BASIC_MODEL(params)``            models/basic_wdsr_b.py:16-93
* ``Block(...)``                     models/basic_wdsr_b.py:96-144 / models/wdsr_b.py:253-319 (optional masks)
* ``AggregationLayer(...)``          models/wdsr_b.py:322-373 (depth gate, eval branch :358-365)
* ``NAS_MODEL_classic(params)``      the UPSTREAM supernet design: models/wdsr_b.py:30-137 with the classic AggregationLayer body
                                      (the fork's own ``NAS_MODEL`` -- Split_Block body + speed estimator -- lives in ``nas.py``)
* ``Model(scale, filename)``         export_onnx.py:6-88 (pruned net from a search ``block_index.txt``)

Weight-norm (recomputed by a hook on every reference forward) is folded once in fp32 when the
weights change; masks and depth gates are resolved into pruned filter slices at the same time.
The parameter containers are real ``weight_norm(nn.Conv2d)`` modules so seeded construction consumes
the RNG exactly like the reference and checkpoints load ``strict=True`` -- but their forward (cuDNN)
is never called.  Inference only (``torch.no_grad`` semantics); there is no CPU fallback.
"""
from __future__ import annotations

import ctypes
import os
import math
import warnings
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.init as init

from . import _lib
from .masks import BinaryConv2d, rounding

__all__ = ["BASIC_MODEL", "NAS_MODEL_classic", "Model", "Block", "AggregationLayer", "WdsrPlan"]


def _weight_norm(conv: nn.Conv2d) -> nn.Conv2d:
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return torch.nn.utils.weight_norm(conv)


def _fold(conv: nn.Module) -> Tuple[torch.Tensor, torch.Tensor]:
    """W = g * v / ||v|| in fp32 on the host -- the same ``torch._weight_norm`` the reference's hook calls."""
    with torch.no_grad():
        w = torch._weight_norm(conv.weight_v.detach().float().cpu(), conv.weight_g.detach().float().cpu(), 0)
        return w.contiguous(), conv.bias.detach().float().cpu().contiguous()


def _ptr(t: torch.Tensor) -> ctypes.c_void_p:
    return ctypes.c_void_p(t.data_ptr())


class WdsrPlan:
    """Owner of one ``b200sr_wdsr_t`` handle: folded weights resident on one device + reusable workspace."""

    def __init__(self, scale: int, c_trunk: int, blocks: Sequence[Sequence[torch.Tensor]], head, tail, skip,
                 add_mean: bool, image_mean: float, device: torch.device):
        L = _lib.lib()
        self.scale, self.c_trunk, self.device = int(scale), int(c_trunk), device
        nb = len(blocks)
        m1 = (ctypes.c_int32 * max(nb, 1))(*[int(b[0].shape[0]) for b in blocks])
        m2 = (ctypes.c_int32 * max(nb, 1))(*[int(b[2].shape[0]) for b in blocks])
        desc = _lib.WdsrDesc(int(scale), nb, int(c_trunk), 1 if add_mean else 0, float(image_mean), m1, m2)
        h = ctypes.c_void_p()
        _lib.check(L.b200sr_wdsr_create(ctypes.byref(desc), ctypes.byref(h)))
        self._h = h
        try:
            hw, hb = [t.contiguous().float() for t in head]
            _lib.check(L.b200sr_wdsr_set_head(h, _ptr(hw), _ptr(hb)))
            for i, blk in enumerate(blocks):
                ts = [t.contiguous().float() for t in blk]
                _lib.check(L.b200sr_wdsr_set_block(h, i, *[_ptr(t) for t in ts]))
            tw, tb = [t.contiguous().float() for t in tail]
            sw, sb = [t.contiguous().float() for t in skip]
            _lib.check(L.b200sr_wdsr_set_tail(h, _ptr(tw), _ptr(tb), _ptr(sw), _ptr(sb)))
            with torch.cuda.device(device):
                _lib.check(L.b200sr_wdsr_commit(h))
        except Exception:
            L.b200sr_wdsr_destroy(h)
            self._h = None
            raise
        self._ws: Optional[torch.Tensor] = None

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                _lib.lib().b200sr_wdsr_destroy(h)
            except Exception:
                pass

    # -- helpers ------------------------------------------------------------------------------
    @property
    def handle(self):
        return self._h

    @property
    def trunk_channels(self) -> int:
        return _lib.lib().b200sr_wdsr_trunk_channels(self._h)

    def workspace(self, n: int, h: int, w: int, prec: int) -> torch.Tensor:
        need = _lib.lib().b200sr_wdsr_workspace_bytes(self._h, n, h, w, prec)
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws

    def launches_per_forward(self) -> int:
        return _lib.lib().b200sr_wdsr_launches_per_forward(self._h)

    # -- compute ------------------------------------------------------------------------------
    def forward(self, x: torch.Tensor, precision: str, out_dtype: Optional[torch.dtype] = None,
                out: Optional[torch.Tensor] = None) -> torch.Tensor:
        _lib.require_cuda_tensor(x, "input")
        if x.dim() != 4 or x.shape[1] != 3:
            raise RuntimeError(f"expected input of shape (N,3,H,W), got {tuple(x.shape)}")
        if x.device != self.device:
            raise RuntimeError(f"input on {x.device}, plan on {self.device}")
        x = x.contiguous()
        n, _, h, w = x.shape
        prec = _lib.precision_code(precision)
        out_dtype = out_dtype or x.dtype
        s = self.scale
        if out is None:
            out = torch.empty((n, 3, s * h, s * w), dtype=out_dtype, device=x.device)
        elif tuple(out.shape) != (n, 3, s * h, s * w) or out.device != x.device or not out.is_contiguous():
            raise RuntimeError(f"out must be a contiguous {(n, 3, s * h, s * w)} tensor on {x.device}, got "
                               f"{tuple(out.shape)} on {out.device} (contiguous={out.is_contiguous()})")
        if n == 0 or h == 0 or w == 0:
            return out
        ws = self.workspace(n, h, w, prec)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().b200sr_wdsr_forward(self._h, _ptr(x), _lib.dtype_code(x.dtype), _ptr(out),
                                                      _lib.dtype_code(out.dtype), n, h, w, prec, _ptr(ws), ws.numel(),
                                                      _lib.current_stream_ptr(self.device)))
        return out

    def forward_host(self, x_host: torch.Tensor, y_host: torch.Tensor, precision: str, x_stage: torch.Tensor,
                     y_stage: torch.Tensor) -> None:
        """Host-buffer entry (b200sr_wdsr_forward_host): H2D, forward, D2H enqueued on the current stream."""
        n, _, h, w = x_host.shape
        prec = _lib.precision_code(precision)
        ws = self.workspace(n, h, w, prec)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().b200sr_wdsr_forward_host(
                self._h, _ptr(x_host), _lib.dtype_code(x_host.dtype), _ptr(y_host), _lib.dtype_code(y_host.dtype), n, h, w,
                prec, _ptr(x_stage), _ptr(y_stage), _ptr(ws), ws.numel(), _lib.current_stream_ptr(self.device)))

    # stage-level (tests)
    # ---- stage-level calls (parity tests, bench).  The public tensors are NHWC; on the bf16 tcgen05 path the kernels keep the
    #      trunk planar-8 ([n][c/8][h][w][8]), so these wrappers convert on the way in and out.  ``*_internal`` skip the conversion.
    def _planar(self, prec: int) -> bool:
        return _lib.lib().b200sr_wdsr_trunk_layout(self._h, prec) == 1

    def to_internal(self, trunk_nhwc: torch.Tensor, precision: str) -> torch.Tensor:
        if not self._planar(_lib.precision_code(precision)):
            return trunk_nhwc.contiguous()
        n, h, w, c = trunk_nhwc.shape
        return trunk_nhwc.view(n, h, w, c // 8, 8).permute(0, 3, 1, 2, 4).contiguous()

    def from_internal(self, trunk: torch.Tensor, precision: str) -> torch.Tensor:
        if not self._planar(_lib.precision_code(precision)):
            return trunk
        n, cb, h, w, _ = trunk.shape
        return trunk.permute(0, 2, 3, 1, 4).reshape(n, h, w, cb * 8).contiguous()

    def head_internal(self, x: torch.Tensor, precision: str) -> torch.Tensor:
        n, _, h, w = x.shape
        prec = _lib.precision_code(precision)
        c = self.trunk_channels
        shape = (n, c // 8, h, w, 8) if self._planar(prec) else (n, h, w, c)
        t = torch.empty(shape, dtype=torch.float32 if prec == _lib.F32 else torch.bfloat16, device=x.device)
        x = x.contiguous()
        _lib.check(_lib.lib().b200sr_wdsr_head(self._h, _ptr(x), _lib.dtype_code(x.dtype), _ptr(t), n, h, w, prec,
                                               _lib.current_stream_ptr(self.device)))
        return t

    def block_internal(self, i: int, trunk: torch.Tensor, precision: str, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        prec = _lib.precision_code(precision)
        n, h, w = (trunk.shape[0], trunk.shape[2], trunk.shape[3]) if self._planar(prec) else trunk.shape[:3]
        assert trunk.is_contiguous()
        out = torch.empty_like(trunk) if out is None else out
        _lib.check(_lib.lib().b200sr_wdsr_block(self._h, i, _ptr(trunk), _ptr(out), n, h, w, prec, _lib.current_stream_ptr(self.device)))
        return out

    def head(self, x: torch.Tensor, precision: str) -> torch.Tensor:
        return self.from_internal(self.head_internal(x, precision), precision)

    def block(self, i: int, trunk: torch.Tensor, precision: str) -> torch.Tensor:
        n, h, w, c = trunk.shape
        assert c == self.trunk_channels
        return self.from_internal(self.block_internal(i, self.to_internal(trunk, precision), precision), precision)

    def tail(self, trunk: torch.Tensor, x: torch.Tensor, precision: str, out_dtype=None) -> torch.Tensor:
        n, h, w, _ = trunk.shape
        x = x.contiguous()
        trunk = self.to_internal(trunk, precision)
        out = torch.empty((n, 3, self.scale * h, self.scale * w), dtype=out_dtype or x.dtype, device=x.device)
        _lib.check(_lib.lib().b200sr_wdsr_tail(self._h, _ptr(trunk), _ptr(x), _lib.dtype_code(x.dtype), _ptr(out),
                                               _lib.dtype_code(out.dtype), n, h, w, _lib.precision_code(precision),
                                               _lib.current_stream_ptr(self.device)))
        return out


# ------------------------------------------------------------------------------------------------------
# modules
# ------------------------------------------------------------------------------------------------------
class _PlanCacheMixin:
    """Folded-weight cache: rebuilt when any parameter is modified in place (``_version``), re-assigned
    (``data_ptr``), the module moves device, or a plain attribute folded into the plan (``image_mean``, ``scale``)
    changes -- the reference re-folds on every forward, so mutating ``weight_g`` after construction must change
    the output here too (hard part 7).

    Two things the signature cannot see, by contract:
      * writes through ``.data`` (``p.data.copy_()``, ``p.data.clamp_()``, EMA swaps) do NOT bump ``_version``:
        call ``invalidate()`` after them;
      * ``freeze()`` skips the per-forward signature walk (153 parameters for the 16-block net, ~30 us of Python
        that a 100 us forward notices) until ``invalidate()`` / ``unfreeze()``: the serving form.
    ``B200SR_DEBUG_REFOLD=1`` re-folds on every forward like the reference does (slow; for chasing stale-plan bugs).
    """

    precision: str = "fp32"
    _frozen: bool = False

    def set_precision(self, precision: str):
        _lib.precision_code(precision)
        self.precision = precision
        return self

    def _signature(self, device):
        return (str(device), getattr(self, "image_mean", None), getattr(self, "scale", None)) + \
            tuple((p.data_ptr(), p._version) for p in self.parameters())

    def invalidate(self):
        """Drop the folded plan: the next forward (or ``prepare()``) folds and uploads again."""
        self._plan_sig = None
        self._frozen = False
        return self

    def freeze(self, device=None):
        """Prepare now and stop checking the parameters on every forward (until ``invalidate()`` / ``unfreeze()``)."""
        self.prepare(device)
        self._frozen = True
        return self

    def unfreeze(self):
        self._frozen = False
        return self

    def _get_plan(self, device) -> WdsrPlan:
        if self._frozen and getattr(self, "_plan", None) is not None and self._plan.device == device:
            return self._plan
        sig = self._signature(device)
        if getattr(self, "_plan_sig", None) != sig or os.environ.get("B200SR_DEBUG_REFOLD") == "1":
            self._plan = self._build_plan(device)
            self._plan_sig = sig
        return self._plan

    def prepare(self, device=None) -> WdsrPlan:
        """Fold + upload now (otherwise done lazily by the first forward)."""
        device = torch.device(device) if device is not None else next(self.parameters()).device
        if device.type != "cuda":
            raise RuntimeError("b200sr: prepare() needs a CUDA device -- there is no CPU fallback")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        return self._get_plan(device)


def _block_filters(body: nn.Sequential, idx: Sequence[int]):
    e, r, c = (body[i] for i in idx)
    w1, b1 = _fold(e)
    w2, b2 = _fold(r)
    w3, b3 = _fold(c)
    return [w1.flatten(1), b1, w2.flatten(1), b2, w3, b3]


def _slice_block(f, keep_in=None, keep1=None, keep2=None):
    """Masked supernet block -> pruned (IN,M1,M2) block: keep rows/cols where the mask is 1."""
    w1, b1, w2, b2, w3, b3 = f
    if keep1 is not None:
        w1, b1, w2 = w1[keep1], b1[keep1], w2[:, keep1]
    if keep2 is not None:
        w2, b2, w3 = w2[keep2], b2[keep2], w3[:, keep2]
    if keep_in is not None:
        w1, w3, b3 = w1[:, keep_in], w3[keep_in], b3[keep_in]
    return [t.contiguous() for t in (w1, b1, w2, b2, w3, b3)]


class Block(nn.Module, _PlanCacheMixin):
    """Residual block ``x + conv3x3(conv1x1(relu(conv1x1(x))))``.

    Constructor of models/wdsr_b.py:255-260 (a superset of models/basic_wdsr_b.py:98-103): expand=6,
    linear=0.84, weight_g init 2.0 / 2.0 / res_scale, zero biases; ``width_search=True`` inserts the two
    ``BinaryConv2d`` masks so the Sequential indices match the reference (0,3,5 instead of 0,2,3).
    ``forward(x)`` takes/returns NCHW like the reference; inside a model the block runs NHWC-fused.
    """

    def __init__(self, num_residual_units, kernel_size, weight_norm=torch.nn.utils.weight_norm, res_scale=1,
                 width_search=False, **kwargs):
        super().__init__()
        expand, linear = 6, 0.84
        m1, m2 = int(num_residual_units * expand), int(num_residual_units * linear)
        body: List[nn.Module] = []
        conv = _weight_norm(nn.Conv2d(num_residual_units, m1, 1, padding=0))
        init.constant_(conv.weight_g, 2.0)
        init.zeros_(conv.bias)
        body += [conv, nn.ReLU(inplace=True)]
        if width_search:
            body.append(BinaryConv2d(in_channels=m1, out_channels=m1, groups=m1))
        conv = _weight_norm(nn.Conv2d(num_residual_units * expand, m2, 1, padding=0))
        init.constant_(conv.weight_g, 2.0)
        init.zeros_(conv.bias)
        body.append(conv)
        if width_search:
            body.append(BinaryConv2d(in_channels=m2, out_channels=m2, groups=m2))
        conv = _weight_norm(nn.Conv2d(m2, num_residual_units, kernel_size, padding=kernel_size // 2))
        init.constant_(conv.weight_g, res_scale)
        init.zeros_(conv.bias)
        body.append(conv)
        self.body = nn.Sequential(*body)
        self.width_search = bool(width_search)
        self.num_residual_units = num_residual_units
        if kernel_size != 3:
            raise NotImplementedError("b200sr Block: kernel_size must be 3")

    # -- prepare-time views ----------------------------------------------------------------------
    def conv_indices(self) -> Tuple[int, int, int]:
        return (0, 3, 5) if self.width_search else (0, 2, 3)

    def pruned_filters(self, keep_in: Optional[torch.Tensor] = None):
        f = _block_filters(self.body, self.conv_indices())
        keep1 = self.body[2].keep_indices() if self.width_search else None
        keep2 = self.body[4].keep_indices() if self.width_search else None
        return _slice_block(f, keep_in, keep1, keep2)

    def is_skipped(self) -> bool:
        return False

    # -- standalone forward (NCHW in/out, like the reference module) ----------------------------
    def _build_plan(self, device) -> WdsrPlan:
        c = self.num_residual_units
        z = lambda *s: torch.zeros(*s)
        return WdsrPlan(2, c, [self.pruned_filters()], (z(c, 3, 3, 3), z(c)), (z(12, c, 3, 3), z(12)),
                        (z(12, 3, 5, 5), z(12)), True, 0.0, device)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        _lib.require_cuda_tensor(x, "input")
        plan = self._get_plan(x.device)
        cp = plan.trunk_channels
        dt = torch.float32 if self.precision == "fp32" else torch.bfloat16
        t = x.permute(0, 2, 3, 1).to(dt)
        if cp != t.shape[-1]:
            t = torch.nn.functional.pad(t, (0, cp - t.shape[-1]))
        y = plan.block(0, t.contiguous(), self.precision)
        return y[..., : x.shape[1]].permute(0, 3, 1, 2).to(x.dtype).contiguous()


class AggregationLayer(Block):
    """Block + depth gate (models/wdsr_b.py:322-373).  Eval semantics (:358-365): identity iff
    ``alpha1 >= alpha2``; ``speed_accu += beta2 * speed_curr`` is scalar host arithmetic."""

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.alpha1 = nn.Parameter(torch.empty(1), requires_grad=True)
        self.beta1 = nn.Parameter(torch.zeros(1), requires_grad=True)
        init.uniform_(self.alpha1, 0, 0.2)
        self.alpha2 = nn.Parameter(torch.empty(1), requires_grad=True)
        self.beta2 = nn.Parameter(torch.ones(1), requires_grad=True)
        init.uniform_(self.alpha2, 0.8, 1)

    def is_skipped(self) -> bool:
        return bool(self.alpha1.detach() >= self.alpha2.detach())

    def forward(self, x, speed_curr, speed_accu):
        if self.training:
            raise NotImplementedError("b200sr is inference-only: call .eval() (training branch models/wdsr_b.py:343-357)")
        if not self.is_skipped():
            x = Block.forward(self, x)
        return x, speed_accu + self.beta2.detach().to(speed_accu.device) * speed_curr

    def get_num_channels(self):
        # models/wdsr_b.py:366-372 appends in_channels of EVERY nn.Conv2d child, and BinaryConv2d is one: with
        # width_search=True the reference returns [nru, m1, m1, m2, m2, nru] -- reproduced
        ch = [m.in_channels for m in self.body.children() if isinstance(m, nn.Conv2d)]
        return ch + [ch[0]]


class _WdsrNet(nn.Module, _PlanCacheMixin):
    scale: int
    image_mean: float

    def _check_channels(self, num_channels):
        if num_channels != 3:
            raise NotImplementedError("b200sr kernels are specialised for 3-channel (RGB) images")

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        _lib.require_cuda_tensor(x, "input")
        return self._get_plan(x.device).forward(x, self.precision)

    def forward_u8(self, x: torch.Tensor) -> torch.Tensor:
        """The forward with the 8-bit frame ``(sr * 255).round().clamp(0, 255)`` (common/metrics.py:12, what the reference's
        evaluation turns every output into before PSNR / PNG) written straight from the tail kernel's epilogue: uint8 (N,3,sH,sW),
        a quarter of the float32 bytes to bring back over PCIe.  bf16 precision on the tcgen05 path only (raises otherwise)."""
        _lib.require_cuda_tensor(x, "input")
        return self._get_plan(x.device).forward(x, self.precision, out_dtype=torch.uint8)

    def launches_per_forward(self) -> int:
        return self._plan.launches_per_forward()


class BASIC_MODEL(_WdsrNet):
    """models/basic_wdsr_b.py:16-93.  ``params``: image_mean, num_channels, scale, num_blocks, num_residual_units."""

    def __init__(self, params):
        super().__init__()
        self.image_mean = params.image_mean
        self._check_channels(params.num_channels)
        num_inputs, scale = params.num_channels, params.scale
        self.scale = scale
        self.remain_blocks = params.num_blocks
        self.kwargs = {}
        num_outputs = scale * scale * params.num_channels
        nru = params.num_residual_units

        conv = _weight_norm(nn.Conv2d(num_inputs, nru, 3, padding=1))
        init.ones_(conv.weight_g)
        init.zeros_(conv.bias)
        self.head = conv
        self.body = nn.ModuleList(
            [Block(num_residual_units=nru, kernel_size=3, res_scale=1 / math.sqrt(params.num_blocks))
             for _ in range(params.num_blocks)])
        conv = _weight_norm(nn.Conv2d(nru, num_outputs, 3, padding=1))
        init.ones_(conv.weight_g)
        init.zeros_(conv.bias)
        self.tail = conv
        skip = []
        if num_inputs != num_outputs:
            conv = _weight_norm(nn.Conv2d(num_inputs, num_outputs, 5, padding=2))
            init.ones_(conv.weight_g)
            init.zeros_(conv.bias)
            skip.append(conv)
        else:
            raise NotImplementedError("scale 1 (identity skip) is not on the accelerated path")
        self.skip = nn.Sequential(*skip)
        self.shuf = nn.Sequential(*([nn.PixelShuffle(scale)] if scale > 1 else []))

    def _build_plan(self, device) -> WdsrPlan:
        blocks = [b.pruned_filters() for b in self.body]
        return WdsrPlan(self.scale, self.head.out_channels, blocks, _fold(self.head), _fold(self.tail), _fold(self.skip[0]),
                        True, self.image_mean, device)


class NAS_MODEL_classic(_WdsrNet):
    """Supernet of models/wdsr_b.py:30-137 with the classic ``AggregationLayer`` body (the upstream design the
    north star describes: 1x1 expand / reduce / 3x3 blocks with width masks).  The fork's ``NAS_MODEL`` as committed
    (``MyAggregationLayer`` / ``Split_Block`` body, ``speed_estimator``) is ``nas.NAS_MODEL``.

    ``forward`` returns ``(sr, speed_accu)`` like the reference (:137).  Width masks (``mask``, per-block
    ``body.2``/``body.4``) and depth gates are resolved at prepare time into a pruned plan; nothing is
    multiplied by a mask at run time.  ``speed_accu`` uses the reference's analytic latency proxy
    ``(c1 + 0.2*c0) * k^2 / 40`` (speed_models/speed_estimator.py:41,75); the phone-latency MLP side-car is
    out of scope.
    """

    def __init__(self, params):
        super().__init__()
        self.image_mean = params.image_mean
        self._check_channels(params.num_channels)
        scale = params.scale
        self.scale = scale
        self.num_blocks = params.num_blocks
        self.num_residual_units = nru = params.num_residual_units
        self.remain_blocks = params.num_blocks
        self.width_search = bool(params.width_search)
        num_outputs = scale * scale * params.num_channels

        conv = _weight_norm(nn.Conv2d(params.num_channels, nru, 3, padding=1))
        init.ones_(conv.weight_g)
        init.zeros_(conv.bias)
        self.head = conv
        self.body = nn.ModuleList(
            [AggregationLayer(num_residual_units=nru, kernel_size=3, res_scale=1 / math.sqrt(params.num_blocks),
                              width_search=self.width_search) for _ in range(params.num_blocks)])
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

    # search read-outs (models/wdsr_b.py:139-183), host-side only
    @torch.no_grad()
    def get_block_status(self) -> List[int]:
        return [i for i, m in enumerate(self.body) if not m.is_skipped()]

    @torch.no_grad()
    def get_current_blocks(self) -> int:
        return len(self.get_block_status())

    @torch.no_grad()
    def get_width_from_block_idx(self, remain_block_idx) -> List[List[int]]:
        """(IN, M1, M2) per kept block -- the ``block_index.txt`` triple ``export_onnx.Model`` consumes."""
        n_in = int(rounding(self.mask.weight.detach().cpu()).sum()) if self.width_search else self.num_residual_units
        out = []
        for i, m in enumerate(self.body):
            if i in remain_block_idx:
                f = m.pruned_filters()
                out.append([n_in, int(f[0].shape[0]), int(f[2].shape[0])])
        return out

    def _build_plan(self, device) -> WdsrPlan:
        keep_in = self.mask.keep_indices() if self.width_search else None
        blocks = [m.pruned_filters(keep_in) for m in self.body if not m.is_skipped()]
        hw, hb = _fold(self.head)
        tw, tb = _fold(self.tail)
        if keep_in is not None:
            hw, hb, tw = hw[keep_in], hb[keep_in], tw[:, keep_in]
        c = int(hw.shape[0])
        return WdsrPlan(self.scale, c, blocks, (hw, hb), (tw, tb), _fold(self.skip), True, self.image_mean, device)

    @torch.no_grad()
    def speed_accu(self) -> torch.Tensor:
        c0 = float(rounding(self.mask.weight.detach().cpu()).sum()) if self.width_search else float(self.num_residual_units)
        total = torch.zeros(1)
        for m in self.body:
            f = m.pruned_filters()
            total = total + m.beta2.detach().cpu() * ((float(f[0].shape[0]) + 0.2 * c0) * 9.0 / 40.0)
        return total

    def forward(self, x: torch.Tensor):
        out = super().forward(x)
        return out, self.speed_accu().to(x.device)


class Model(_WdsrNet):
    """Pruned WDSR-B rebuilt from a search artefact: export_onnx.py:6-88.

    ``filename``'s last line is a Python tuple whose element [1] lists ``(IN, M1, M2)`` per kept block
    (written by search.py:125-126).  Everything incl. head and tail sits in one ``nn.Sequential`` ``body``
    and the forward does NOT add ``image_mean`` back (export_onnx.py:59-79) -- both reproduced.
    """

    def __init__(self, scale, filename):
        super().__init__()
        self.image_mean = 0.5
        self.scale = scale
        num_inputs, num_outputs = 3, scale * scale * 3
        status = self.file_reader(filename)
        body: List[nn.Module] = [_weight_norm(nn.Conv2d(num_inputs, self.IN, 3, padding=1))]
        for (IN, M1, M2) in status:
            if IN != self.IN:
                raise ValueError("all blocks of a pruned model share the trunk width IN")
            body.append(_PrunedBlock(IN, M1, M2))
        body.append(_weight_norm(nn.Conv2d(self.IN, num_outputs, 3, padding=1)))
        self.body = nn.Sequential(*body)
        self.skip = _weight_norm(nn.Conv2d(num_inputs, num_outputs, 5, padding=2))
        self.shuf = nn.Sequential(*([nn.PixelShuffle(scale)] if scale > 1 else []))

    def file_reader(self, filename):
        import ast
        with open(filename, "r") as f:
            status = ast.literal_eval(f.readlines()[-1].strip())[1]
        self.IN = status[0][0]
        return [tuple(int(v) for v in s) for s in status]

    def _build_plan(self, device) -> WdsrPlan:
        mods = list(self.body)
        blocks = [_block_filters(m.body, (0, 2, 3)) for m in mods[1:-1]]
        return WdsrPlan(self.scale, self.IN, blocks, _fold(mods[0]), _fold(mods[-1]), _fold(self.skip), False,
                        self.image_mean, device)


class _PrunedBlock(nn.Module):
    """export_onnx.py:91-114 (parameter container; runs fused inside ``Model``)."""

    def __init__(self, IN, M1, M2):
        super().__init__()
        self.body = nn.Sequential(_weight_norm(nn.Conv2d(IN, M1, 1)), nn.ReLU(inplace=True),
                                  _weight_norm(nn.Conv2d(M1, M2, 1)), _weight_norm(nn.Conv2d(M2, IN, 3, padding=1)))
