"""``Naive_model`` of models/naive_multi_model_easy.py:33-206 over the C ABI: SPyNet flow between consecutive frames,
the previous frame's ENCODED features warped to the current one, a stack of conv-ReLU-conv residual blocks and a conv + PixelShuffle(4) tail
on a bilinear base.

Same constructor (``scale, filename, spynet_pretrained``), forward signature and ``state_dict`` layout as the reference:
``flownet.*`` (SPyNet), ``encode`` / ``decode`` / ``skip`` weight-normed (``weight_g`` / ``weight_v``), ``body.<i>.body.{0,2}`` plain convs and the
never-used ``body.<i>.skip`` 1x1 (a parameter container, like upstream).  Upstream quirks mirrored on purpose:

* ``decode`` takes the kernel size of the LAST block (the constructor's loop re-binds ``kernel_size``, :72/:82/:91-96);
* the 5x5 ``skip`` conv is constructed but never applied -- the base is ``F.interpolate(x, scale_factor=4, 'bilinear')`` (:140), so the
  forward only composes for ``scale == 4`` (the reference fails with a shape error otherwise; so does this class);
* frame 0 is concatenated with its own features and a zero flow (:123-127); frame i >= 1 with ``flow_warp(previous encode output, flow)``.

Layout: activations NHWC, the first block's input kept as ``[warped | current | flow | zero pad]`` (its filters' input channels re-ordered to
match) so that the warp kernel writes its window in place.  No CPU fallback: the convolutions, the warp and the tail are b200sr kernels.
"""
from __future__ import annotations

import ast
import ctypes
from typing import Dict, List

import torch
import torch.nn as nn

from . import _lib
from .video import ACT_NONE, ACT_RELU, SpyNet, _ConvHandle, _VideoPlanMixin, _ptr, flow_warp, flow_warp_nhwc
from .wdsr import _fold, _weight_norm

__all__ = ["Naive_model", "NaiveBlock"]


class NaiveBlock(nn.Module):
    """``Block`` of models/naive_multi_model_easy.py:157-183 (parameter container; the forward lives in Naive_model)."""

    def __init__(self, IN, OUT, split, kernel_size, weight_norm=None):
        super().__init__()
        self.split = IN - split
        self.IN, self.OUT, self.conv_channel = IN, OUT, split
        # creation order (RNG) and registration order (state_dict) of the reference: body convs are created first, `skip` is created and
        # registered next, the Sequential is registered last (:166-174)
        c0 = nn.Conv2d(IN, OUT, kernel_size, padding=kernel_size // 2)
        c2 = nn.Conv2d(OUT, OUT, kernel_size, padding=kernel_size // 2)
        self.skip = nn.Conv2d(2 * self.IN, self.IN, 1, padding=0)          # never used in forward (upstream too)
        self.body = nn.Sequential(c0, nn.ReLU(inplace=True), c2)


class Naive_model(nn.Module, _VideoPlanMixin):
    def __init__(self, scale, filename, spynet_pretrained=None):
        super().__init__()
        self.image_mean = 0.5
        kernel_size, skip_kernel_size, num_inputs = 3, 5, 3
        self.scale = scale
        self.idx = self.file_reader(filename)
        self.IN = self.idx[0][0]
        # parameter order of the reference constructor: flownet, body (empty dict), encode, body entries, decode, skip -- seeded
        # construction consumes the RNG in the same order
        self.flownet = SpyNet(spynet_pretrained)
        for m in self.flownet.parameters():
            m.requires_grad = False
        num_outputs = scale * scale * num_inputs
        self.body = nn.ModuleDict()
        self.encode = _weight_norm(nn.Conv2d(num_inputs, self.IN, kernel_size, padding=kernel_size // 2))
        for i, block in enumerate(self.idx):
            cin = block[0] * 2 + 2 if i == 0 else block[0]
            kernel_size = block[2]                                # (re-bound: decode below uses the last block's)
            self.body[str(i)] = NaiveBlock(cin, block[0], block[1], kernel_size)
        self.decode = _weight_norm(nn.Conv2d(self.IN, num_outputs, kernel_size, padding=kernel_size // 2))
        self.skip = _weight_norm(nn.Conv2d(num_inputs, num_outputs, skip_kernel_size, padding=skip_kernel_size // 2))
        self.shuf = nn.Sequential(*([nn.PixelShuffle(scale)] if scale > 1 else []))

    def file_reader(self, filename):
        with open(filename, "r") as f:
            status = ast.literal_eval(f.readlines()[-1].replace("\n", ""))[1]     # the reference eval()s the line
        self.IN = status[0][0]
        return status

    # ------------------------------------------------------------------------------------------------------------------
    def _handles(self, device) -> Dict[str, _ConvHandle]:
        own = [self.encode, self.decode] + [m for b in self.body.values() for m in (b.body[0], b.body[2])]
        sig = (str(device),) + tuple((p.data_ptr(), p._version) for m in own for p in m.parameters())
        if getattr(self, "_naive_sig", None) != sig:
            IN = self.IN
            cs0 = (2 * IN + 2 + 7) // 8 * 8

            def plain(w, b, k):
                c = nn.Conv2d(w.shape[1], w.shape[0], k, padding=k // 2)
                with torch.no_grad():
                    c.weight.copy_(w), c.bias.copy_(b)
                return c

            hs: Dict[str, _ConvHandle] = {}
            w, b = _fold(self.encode)
            hs["encode"] = _ConvHandle(plain(w, b, w.shape[-1]), device)
            w, b = _fold(self.decode)
            hs["decode"] = _ConvHandle(plain(w, b, w.shape[-1]), device)
            for i, blk in self.body.items():
                c0, c2 = blk.body[0], blk.body[2]
                w0 = c0.weight.detach().float().cpu()
                if i == "0":
                    # reference order of the concatenation: [flow(2) | warped(IN) | current(IN)]; ours: [warped | current | flow | pad]
                    wr = torch.zeros(w0.shape[0], cs0, w0.shape[2], w0.shape[3])
                    wr[:, :IN], wr[:, IN:2 * IN], wr[:, 2 * IN:2 * IN + 2] = w0[:, 2:2 + IN], w0[:, 2 + IN:], w0[:, :2]
                    w0 = wr
                hs[f"{i}a"] = _ConvHandle(plain(w0, c0.bias.detach().float().cpu(), w0.shape[-1]), device)
                hs[f"{i}b"] = _ConvHandle(plain(c2.weight.detach().float().cpu(), c2.bias.detach().float().cpu(), c2.weight.shape[-1]), device)
            self._naive_handles, self._naive_sig, self._cs0 = hs, sig, cs0
        return self._naive_handles

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x (b,n,3,h,w) -> (b,n,3,4h,4w) float32."""
        _lib.require_cuda_tensor(x, "x")
        B, N, C, H, W = x.shape
        if self.scale != 4:
            raise RuntimeError(f"The size of tensor a ({self.scale * W}) must match the size of tensor b ({4 * W}) at non-singleton dimension 3 "
                               "(Naive_model adds a x4 bilinear base whatever its scale, models/naive_multi_model_easy.py:140-144)")
        dev, p = x.device, self.precision
        x = x.contiguous()
        if x.dtype != torch.float32:
            x = x.float()
        IN, act = self.IN, self._act_dtype()
        hs = self._handles(dev)
        cs0 = self._cs0
        L, st = _lib.lib(), _lib.current_stream_ptr(dev)
        self.flownet.set_precision(p)
        if N > 1:
            lqs_1 = x[:, :-1].reshape(-1, C, H, W)
            lqs_2 = x[:, 1:].reshape(-1, C, H, W)
            flows_forward = self.flownet(lqs_2, lqs_1).view(B, N - 1, 2, H, W)
        # the warp kernel's channel-window form needs IN * esize / 16 in {1,2,3,4,6,8,16}; other widths take the NCHW kernel
        esz = 4 if act == torch.float32 else 2
        windowed = (IN * esz) % 16 == 0 and (IN * esz) // 16 in (1, 2, 3, 4, 6, 8, 16)
        out = torch.empty((B, N, 3, 4 * H, 4 * W), dtype=torch.float32, device=dev)
        buf = torch.zeros((B, H, W, cs0), dtype=act, device=dev)          # [warped | current | flow | zero pad]
        pre = None
        for i in range(N):
            xi = x[:, i]
            xin = xi.permute(0, 2, 3, 1).contiguous().to(act)
            feat = hs["encode"](xin, p, ACT_NONE)                          # (B,H,W,IN)
            buf[..., IN:2 * IN].copy_(feat)
            if i == 0:
                buf[..., :IN].copy_(feat)
                buf[..., 2 * IN:2 * IN + 2].zero_()
            else:
                flow = flows_forward[:, i - 1].contiguous()
                if windowed:
                    flow_warp_nhwc(pre, flow, "zeros", out=buf, out_coff=0)
                else:
                    wr = flow_warp(pre.permute(0, 3, 1, 2).float().contiguous(), flow.permute(0, 2, 3, 1))
                    buf[..., :IN].copy_(wr.permute(0, 2, 3, 1))
                buf[..., 2 * IN:2 * IN + 2].copy_(flow.permute(0, 2, 3, 1))
            pre = feat
            t = hs["0a"](buf, p, ACT_RELU)
            y = hs["0b"](t, p, ACT_NONE, residual=feat)
            for k in range(1, len(self.idx)):
                t = hs[f"{k}a"](y, p, ACT_RELU)
                y = hs[f"{k}b"](t, p, ACT_NONE, residual=y)
            d = hs["decode"](y, p, ACT_NONE)                               # (B,H,W,48)
            with torch.cuda.device(dev):
                _lib.check(L.b200sr_vsr_shuffle4_base_add(_ptr(d), _lib.dtype_code(d.dtype), d.shape[-1], _ptr(xi), _lib.F32, x.stride(0),
                                                          _ptr(out[:, i]), out.stride(0), B, H, W, st))
        return out

    def load_state_dict(self, state_dict, strict=True, **kw):
        """Accepts mmedit-style SPyNet keys under ``flownet.`` (ConvModule wrappers) as well as the in-repo layout."""
        fl = {k[len("flownet."):]: v for k, v in state_dict.items() if k.startswith("flownet.")}
        rest = {k: v for k, v in state_dict.items() if not k.startswith("flownet.")}
        rest.update({"flownet." + k: v for k, v in SpyNet.remap_mmedit_state_dict(fl).items()})
        return super().load_state_dict(rest, strict=strict, **kw)
