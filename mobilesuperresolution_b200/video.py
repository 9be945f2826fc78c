"""Video path: ``flow_warp`` (and, below, SPyNet / BasicVSR) over the B200 C ABI.

Mirrors models/spynet_arch.py (the in-repo twin of the un-vendored ``mmedit`` functions the BasicVSR
files import).
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib


def _ptr(t: torch.Tensor) -> ctypes.c_void_p:
    return ctypes.c_void_p(t.data_ptr())


def flow_warp(x: torch.Tensor, flow: torch.Tensor, interp_mode: str = "bilinear", padding_mode: str = "zeros",
              align_corners: bool = True) -> torch.Tensor:
    """Warp ``x`` (n,c,h,w) with ``flow`` (n,h,w,2).  Signature and assert of models/spynet_arch.py:98-129.

    ``flow`` may be any strided view (the callers pass ``flow.permute(0,2,3,1)``): it is consumed in place.
    Only the configuration the reference's callers use is accelerated (bilinear, zeros|border, align_corners=True).
    """
    assert x.size()[-2:] == flow.size()[1:3]
    if interp_mode != "bilinear" or not align_corners or padding_mode not in ("zeros", "border"):
        raise NotImplementedError("b200sr.flow_warp: bilinear, padding_mode zeros|border, align_corners=True only")
    _lib.require_cuda_tensor(x, "x")
    _lib.require_cuda_tensor(flow, "flow")
    if x.dtype != torch.float32 or flow.dtype != torch.float32:
        raise TypeError("b200sr.flow_warp: float32 tensors (use flow_warp_nhwc for the bf16 internal layout)")
    x = x.contiguous()
    n, c, h, w = x.shape
    y = torch.empty_like(x)
    if x.numel() == 0:
        return y
    sn, sh, sw, sc = flow.stride()
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().b200sr_flow_warp_nchw(
            _ptr(x), _ptr(flow), sn, sh, sw, sc, _ptr(y), n, c, h, w,
            _lib.PAD_BORDER if padding_mode == "border" else _lib.PAD_ZEROS, _lib.current_stream_ptr(x.device)))
    return y


def flow_warp_nhwc(x: torch.Tensor, flow_nchw: torch.Tensor, padding_mode: str = "zeros") -> torch.Tensor:
    """Internal-layout warp: ``x`` (n,h,w,c) float32|bfloat16 contiguous, ``flow_nchw`` (n,2,h,w) float32."""
    _lib.require_cuda_tensor(x, "x")
    assert x.is_contiguous() and flow_nchw.is_contiguous() and flow_nchw.dtype == torch.float32
    n, h, w, c = x.shape
    assert tuple(flow_nchw.shape) == (n, 2, h, w)
    y = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().b200sr_flow_warp_nhwc(
            _ptr(x), _ptr(flow_nchw), _ptr(y), n, c, h, w,
            _lib.PAD_BORDER if padding_mode == "border" else _lib.PAD_ZEROS, _lib.dtype_code(x.dtype),
            _lib.current_stream_ptr(x.device)))
    return y
