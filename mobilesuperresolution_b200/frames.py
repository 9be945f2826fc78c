"""8-bit frame glue around the forward: what the reference's evaluation does on the host with every frame --
``to_tensor`` on the way in (datasets/_isr.py:74-75), ``(sr * 255).round().clamp(0, 255)`` and PSNR on the way out
(utils/estimate.py:23-133, common/metrics.py:10-19) -- kept on the device so that only 8-bit frames cross PCIe.

    y8 = forward_u8_frames(model, x8)          # uint8 (N,3,H,W) -> uint8 (N,3,sH,sW): quarter of the float32 bytes each way
    p  = psnr_u8(y8, hr8, shave=4)             # == common.metrics.psnr(sr, hr, shave) for hr = to_tensor(hr8), summed over the batch
"""
from __future__ import annotations

import math

import torch

from . import _lib
from .wdsr import _ptr

__all__ = ["u8_to_unit", "forward_u8_frames", "ssd_u8", "psnr_u8"]


def u8_to_unit(x: torch.Tensor, dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """``x / 255`` (torchvision ``to_tensor`` of an 8-bit frame) on the device; float32 or bfloat16 result."""
    _lib.require_cuda_tensor(x, "x")
    if x.dtype != torch.uint8:
        raise TypeError(f"expected a uint8 frame, got {x.dtype}")
    x = x.contiguous()
    y = torch.empty(x.shape, dtype=dtype, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().b200sr_u8_to_unit(_ptr(x), _ptr(y), _lib.dtype_code(dtype), x.numel(), _lib.current_stream_ptr(x.device)))
    return y


def forward_u8_frames(model, x8: torch.Tensor) -> torch.Tensor:
    """8-bit frames in, 8-bit frames out through a WDSR model on the bf16 tcgen05 path (``forward_u8``, the tail writes the frame)."""
    return model.forward_u8(u8_to_unit(x8, torch.bfloat16))


def ssd_u8(a: torch.Tensor, b: torch.Tensor, shave: int = 4) -> torch.Tensor:
    """Per-image sum of squared differences of two uint8 (N,C,H,W) frames over the shaved window; exact (uint64 -> int64 tensor)."""
    _lib.require_cuda_tensor(a, "a")
    if a.dtype != torch.uint8 or b.dtype != torch.uint8 or a.shape != b.shape or a.dim() != 4 or a.device != b.device:
        raise RuntimeError("ssd_u8: two uint8 (N,C,H,W) tensors of the same shape on the same device")
    a, b = a.contiguous(), b.contiguous()
    n, c, h, w = a.shape
    out = torch.empty(n, dtype=torch.int64, device=a.device)
    with torch.cuda.device(a.device):
        _lib.check(_lib.lib().b200sr_ssd_u8(_ptr(a), _ptr(b), _ptr(out), n, c, h, w, int(shave), _lib.current_stream_ptr(a.device)))
    return out


def psnr_u8(sr8: torch.Tensor, hr8: torch.Tensor, shave: int = 4) -> torch.Tensor:
    """``common.metrics.psnr(sr, hr, shave)`` (common/metrics.py:10-19) for ``sr8 = (sr * 255).round().clamp(0, 255)`` and
    ``hr = hr8 / 255``: ``-10 log10(mean((sr8 - hr8)^2) / 255^2)`` per image, summed over the batch like the reference."""
    n, c, h, w = sr8.shape
    count = c * (h - 2 * shave) * (w - 2 * shave)
    mse = ssd_u8(sr8, hr8, shave).double() / (255.0 ** 2 * count)
    return (-10.0 * torch.log10(mse)).sum().float()
